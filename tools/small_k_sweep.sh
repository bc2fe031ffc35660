#!/bin/bash
# Small-circuit latency against the MSM tunables (segment length B200ZK_MSM_SEG_MIN x fixed-base window B200ZK_MSM_PRE_C):
# one tools/small_k_latency.py run per setting, first 60 characters of each line (circuit, k, ms).  Run on a GPU box.
mkdir -p gpurun_out
for seg in 4 8 16 32; do for c in 0 10 11 13; do
  if [ $c = 0 ]; then env="B200ZK_MSM_SEG_MIN=$seg"; else env="B200ZK_MSM_SEG_MIN=$seg B200ZK_MSM_PRE_C=$c"; fi
  echo "== $env"; env $env timeout 120 python tools/small_k_latency.py 2>&1 | cut -c1-60
done; done > gpurun_out/smallk_sweep.txt 2>&1
cat gpurun_out/smallk_sweep.txt
