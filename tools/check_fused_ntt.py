#!/usr/bin/env python
"""Multi-GPU check of the fused column-step + NVLink-scatter NTT against the NCCL all-to-all path
and (small sizes) a single-GPU best_fft:  torchrun --nproc-per-node G tools/check_fused_ntt.py [log_n]"""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402
from bench import random_scalars  # noqa: E402


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    zk = load_package()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    be = zk.Backend(local)
    sharded = importlib.import_module(zk.__name__ + ".sharded")
    log_r = min(10, L // 2)
    dev = torch.device("cuda", local)
    omega = zk.EvaluationDomain(be, 2, L).omega
    omega_c = zk.EvaluationDomain(be, 2, L - log_r).omega
    full = random_scalars(1 << L, 77)                               # same on every rank
    R, C = 1 << log_r, 1 << (L - log_r)
    cg, rg = C // world, R // world
    mine = np.ascontiguousarray(full.reshape(R, C, 4)[:, rank * cg:(rank + 1) * cg])
    block = torch.from_numpy(mine.view(np.int64)).to(dev)
    ref_fs = sharded.FourStepNTTDevice(zk, be, L, log_r, rank, world, dev)
    want = ref_fs.forward(block.clone(), omega, omega_c).cpu().numpy().view(np.uint64)
    fused = sharded.FourStepNTTFused(zk, be, L, log_r, rank, world, dev)
    got = fused.forward(block, omega, omega_c).download((rg, C, 4))
    ok = bool(np.array_equal(got, want))
    if L <= 22:                                                       # rows k_r of X[k_r + R k_c] vs one-GPU transform
        one = be.best_fft(full, omega, L).reshape(C, R, 4)            # [k_c][k_r]
        ok = ok and bool(np.array_equal(got, np.ascontiguousarray(np.transpose(one, (1, 0, 2)))[rank * rg:(rank + 1) * rg]))
    for _ in range(3):
        fused.forward(block, omega, omega_c); ref_fs.forward(block.clone(), omega, omega_c)
    t0 = time.perf_counter()
    for _ in range(10):
        fused.forward(block, omega, omega_c)
    t1 = time.perf_counter()
    for _ in range(10):
        ref_fs.forward(block, omega, omega_c)
    t2 = time.perf_counter()
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        print({"log_n": L, "world": world, "identical_on_all_ranks": all(flags), "fused_ms": round((t1 - t0) * 100, 3), "nccl_all_to_all_ms": round((t2 - t1) * 100, 3)}, flush=True)
    fused.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
