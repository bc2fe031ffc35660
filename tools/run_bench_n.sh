#!/bin/bash
# usage: run_n.sh N extra-args... ; writes gpurun_out/r02_final_bench_n$N$TAG.json
N=$1; shift
TAG=${TAG:-}
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/r02_final_bench_n1$TAG.json 2> gpurun_out/r02_final_bench_n1$TAG.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/r02_final_bench_n$N$TAG.json 2> gpurun_out/r02_final_bench_n$N$TAG.err
fi
tail -2 gpurun_out/r02_final_bench_n$N$TAG.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_final_bench_n$N$TAG.json"))
print("N=$N", "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "verified", d.get("verified"), "phase", {k: round(v, 1) for k, v in d["phase_ms"].items()})
print("timeline", d.get("timeline_ms"))
print("sharded", d.get("sharded"))
print("cpu", d.get("cpu_baseline"))
print("roofline", d["roofline"]["frac"], d["roofline"]["single"]["frac"], d.get("roofline_quotient", {}).get("frac"), d["roofline_msm"]["frac"])
PY
