"""BASELINE.json configs 3 and 4 on one GPU: MSM sweep 2^16..2^26 (dense uniform scalars and a
witness-like column), NTT / coeff_to_extended sweep k = 16..26.  Device-resident, CUDA-event timed.
  python tools/sweep_sizes.py [max_log_msm] [max_log_ntt] > out.jsonl"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402
from bench import IMAD_WIDE_PEAK, msm_work_mul32, random_scalars  # noqa: E402

zk = load_package()


def time_fn(be, fn, reps, warm=2):
    for _ in range(warm):
        fn()
    be.sync()
    be.event_record(0)
    for _ in range(reps):
        fn()
    be.event_record(1)
    be.sync()
    return be.event_elapsed_ms(0, 1) / reps


def main():
    max_msm = int(sys.argv[1]) if len(sys.argv) > 1 else 26
    max_ntt = int(sys.argv[2]) if len(sys.argv) > 2 else 26
    be = zk.Backend(0)
    for L in range(16, max_msm + 1, 2):
        n = 1 << L
        params = zk.ParamsKZG.setup(be, L, random_scalars(1, 4242)[0])
        dense = be.to_device(random_scalars(n, L))
        rng = np.random.Generator(np.random.PCG64(L))
        # witness-like: 99.9 % zeros, the rest uniform
        sparse = np.zeros((n, 4), dtype=np.uint64)
        idx = rng.choice(n, size=max(n // 1000, 1), replace=False)
        sparse[idx] = random_scalars(len(idx), L + 100)
        d_sparse = be.to_device(sparse)
        reps = 5 if L <= 22 else 2
        ms_dense = time_fn(be, lambda: params.commit_dev(dense, n, lagrange=False), reps)
        ms_sparse = time_fn(be, lambda: params.commit_dev(d_sparse, n, lagrange=False), reps)
        print(json.dumps({"msm_log_n": L, "dense_ms": ms_dense, "dense_mpts_per_s": n / 1e3 / ms_dense,
                          "imad_frac": msm_work_mul32(n) / (ms_dense / 1e3) / IMAD_WIDE_PEAK,
                          "witness_like_ms": ms_sparse, "witness_like_mpts_per_s": n / 1e3 / ms_sparse}), flush=True)
        dense.free(); d_sparse.free(); params.close()
    for L in range(16, max_ntt + 1, 2):
        n = 1 << L
        dom = zk.EvaluationDomain(be, 2, L)
        d = be.to_device(random_scalars(n, L))
        reps = 10 if L <= 22 else 3
        ms = time_fn(be, lambda: be.best_fft_dev(d, dom.omega, L), reps)
        rec = {"ntt_log_n": L, "ms": ms, "gb_per_s_64N": 64.0 * n / 1e6 / ms, "g_butterfly_mul_per_s": (n / 2 * L) / 1e6 / ms}
        d.free(); dom.close()
        if L <= 23:
            dom6 = zk.EvaluationDomain(be, 6, L)                       # ext = 8n
            ext_n = dom6.extended_len()
            if ext_n * 32 * 3 < 60e9:
                d_in, d_out = be.to_device(random_scalars(n, L + 1)), be.alloc(ext_n * 32)
                ms2 = time_fn(be, lambda: dom6.coeff_to_extended_dev(d_in, d_out), reps)
                rec.update({"coeff_to_extended_ms": ms2, "coeff_to_extended_gb_per_s": (32.0 * n + 32.0 * ext_n) / 1e6 / ms2})
                d_in.free(); d_out.free()
            dom6.close()
        print(json.dumps(rec), flush=True)
    be.close()


if __name__ == "__main__":
    main()
