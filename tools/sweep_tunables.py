"""Profiling aid: time the NTT / MSM kernels under different tuning overrides (env vars read by
csrc/ntt.cu and csrc/msm.cu).  Run on the GPU box:  python tools/sweep_tunables.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402
from bench import random_scalars  # noqa: E402

zk = load_package()


def time_fn(be, fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    be.sync()
    be.event_record(0)
    for _ in range(reps):
        fn()
    be.event_record(1)
    be.sync()
    return be.event_elapsed_ms(0, 1) / reps


def ntt_times(env, sizes=(20, 23)):
    os.environ.update(env)
    be = zk.Backend(0)
    out = {}
    for L in sizes:
        dom = zk.EvaluationDomain(be, 2, L)
        d = be.to_device(random_scalars(1 << L, L))
        out[L] = time_fn(be, lambda: be.best_fft_dev(d, dom.omega, L))
        d.free(); dom.close()
    be.close()
    for k in env:
        os.environ.pop(k, None)
    return out


def msm_times(env, L=20):
    os.environ.update(env)
    be = zk.Backend(0)
    params = zk.ParamsKZG.setup(be, L, random_scalars(1, 4242)[0])
    n = 1 << L
    dense = be.to_device(random_scalars(n, 1))
    rng = np.random.Generator(np.random.PCG64(3))
    small = np.zeros((n, 4), dtype=np.uint64)
    # Montgomery form of bytes is not small; use raw small limbs as "some field elements with skewed digits":
    # take the dense column and keep only 256 distinct values
    pool = random_scalars(256, 9)
    small[:] = pool[rng.integers(0, 256, size=n)]
    skew = be.to_device(small)
    out = {"dense": time_fn(be, lambda: params.commit_dev(dense, n, lagrange=False)),
           "256-values": time_fn(be, lambda: params.commit_dev(skew, n, lagrange=False))}
    be.close()
    for k in env:
        os.environ.pop(k, None)
    return out


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "ntt"):
        for env in ({}, {"B200ZK_NTT_TILE_CAP": "11"}, {"B200ZK_NTT_TILE_CAP": "10"}, {"B200ZK_NTT_TILE_CAP": "11", "B200ZK_NTT_MAX_M": "8"},
                    {"B200ZK_NTT_TILE_CAP": "11", "B200ZK_NTT_MAX_M": "9"}, {"B200ZK_NTT_TILE_CAP": "12", "B200ZK_NTT_MAX_M": "8", "B200ZK_NTT_MAX_TW": "4"},
                    {"B200ZK_NTT_TILE_CAP": "10", "B200ZK_NTT_MAX_M": "7"}, {"B200ZK_NTT_TILE_CAP": "11", "B200ZK_NTT_MAX_M": "7", "B200ZK_NTT_MAX_TW": "4"}):
            print(json.dumps({"ntt_ms": ntt_times(env), "env": env}), flush=True)
    if which in ("all", "msm"):
        for env in ({}, {"B200ZK_MSM_REDUCE_M": "4"}, {"B200ZK_MSM_REDUCE_M": "3"}, {"B200ZK_MSM_FAST_MAX": "0"},
                    {"B200ZK_MSM_FAST_MAX": "0", "B200ZK_MSM_SEG_MIN": "16"}, {"B200ZK_MSM_FAST_MAX": "0", "B200ZK_MSM_SEG_MIN": "8"},
                    {"B200ZK_MSM_FAST_MAX": "100000"}, {"B200ZK_MSM_FAST_MAX": "0", "B200ZK_MSM_SEG_MIN": "16", "B200ZK_MSM_REDUCE_M": "4"}):
            print(json.dumps({"msm_ms": msm_times(env), "env": env}), flush=True)
