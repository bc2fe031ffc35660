#!/usr/bin/env python
"""NTT launch-shape experiment: single and batched transforms per B200ZK_NTT_WARP_CFG (run once per value).
usage: B200ZK_NTT_WARP_CFG=c python tools/ntt_cfg_bench.py"""
import ctypes, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
from bench import random_scalars, ntt_work_mul32, IMAD_WIDE_PEAK
zk = load_package()
be = zk.Backend(0)
out = {"cfg": os.environ.get("B200ZK_NTT_WARP_CFG", "default")}
def timed(fn, reps=20):
    for _ in range(3): fn()
    be.sync(); be.event_record(0)
    for _ in range(reps): fn()
    be.event_record(1); be.sync()
    return be.event_elapsed_ms(0, 1) / reps
for k in (16, 18, 19, 20, 21):
    n = 1 << k
    dom = zk.EvaluationDomain(be, 2, k)
    d = be.to_device(random_scalars(n, 5))
    ms = timed(lambda: be.best_fft_dev(d, dom.omega, k))
    out[f"single_{k}"] = {"ms": round(ms, 4), "frac": round(ntt_work_mul32(k) / (ms / 1e3) / IMAD_WIDE_PEAK, 3)}
    if k >= 19:
        q = 5
        dr = be.to_device(random_scalars(q * n, 7))
        call = lambda: be._check(zk.lib().b200zk_fft_rows_dev(be._ctx, dr.ptr, ctypes.c_uint32(q), zk._p(zk._fr(dom.omega, 1)), ctypes.c_uint32(k)))
        ms = timed(call)
        out[f"batched5_{k}"] = {"ms": round(ms, 4), "frac": round(q * ntt_work_mul32(k) / (ms / 1e3) / IMAD_WIDE_PEAK, 3)}
        dr.free()
    d.free(); dom.close()
print(json.dumps(out), flush=True)
