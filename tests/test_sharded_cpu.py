"""CPU, world size 2 over gloo: the host-side logic of the multi-GPU paths (point-range sharded
commit with a 96-byte all-gather; four-step NTT with an all-to-all), with the oracle as the local
arithmetic engine.  The reference has no distributed path; these tests pin the
sharded results to the unsharded oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleCommitEngine:
    def __init__(self, orc, bases):
        self.orc, self.bases = orc, bases

    def msm(self, scalars):
        out = self.orc.best_multiexp(scalars, self.bases)
        aff = self.orc.g1_batch_normalize(out)[0]
        return self.orc.g1_from_affine(aff)[0]

    def g1_sum(self, points):
        acc = np.array(points[0])
        for p in points[1:]:
            acc = self.orc.g1_add(acc, p)
        return self.orc.g1_from_affine(self.orc.g1_batch_normalize(acc)[0])[0]


class OracleNttEngine:
    def __init__(self, orc):
        self.orc = orc

    def col_step(self, block, omega_n, log_r, log_c, col0):
        from oracle import pyref
        R, cg = block.shape[0], block.shape[1]
        wn = self.orc.mont_to_ints(omega_n)[0]
        w_r = self.orc.ints_to_mont([pow(wn, 1 << log_c, pyref.R_MOD)])[0]
        out = np.zeros_like(block)
        for c in range(cg):
            col = self.orc.best_fft(np.ascontiguousarray(block[:, c]), w_r, log_r)
            tw = self.orc.ints_to_mont([pow(wn, (col0 + c) * k, pyref.R_MOD) for k in range(R)])
            out[:, c] = self.orc.binop("mul", col, tw)
        return out

    def row_step(self, rows, omega_c, log_c):
        return np.stack([self.orc.best_fft(np.ascontiguousarray(r), omega_c, log_c) for r in rows])


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from __graft_entry__ import load_package
    from oracle import binding as orc
    from oracle import pyref
    import importlib
    zk = load_package()
    sharded = importlib.import_module(zk.__name__ + ".sharded")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # ---- sharded commit vs unsharded oracle
        n = 203                                                  # ragged split
        g, _ = orc.params_setup(8, orc.random_fr(1, 5)[0])
        scalars = orc.random_fr(n, 6)
        lo, hi = sharded.shard_range(n, world, rank)
        sc = sharded.ShardedCommit(OracleCommitEngine(orc, g[lo:hi]))
        got = sc.commit(scalars[lo:hi])
        want = orc.g1_batch_normalize(orc.best_multiexp(scalars, g[:n]))[0]
        assert np.array_equal(got[:8], want), "sharded commit mismatch"
        # empty shard on one rank (n < world is the edge case of shard_range)
        lo1, hi1 = sharded.shard_range(1, world, rank)
        sc1 = sharded.ShardedCommit(OracleCommitEngine(orc, g[lo1:hi1]))
        got1 = sc1.commit(scalars[lo1:hi1])
        assert np.array_equal(got1[:8], orc.g1_batch_normalize(orc.best_multiexp(scalars[:1], g[:1]))[0])
        # ---- four-step NTT vs unsharded oracle
        for log_n, log_r in ((8, 4), (7, 2), (6, 5)):
            N = 1 << log_n
            a = orc.random_fr(N, 40 + log_n)
            w = pyref.omega_for_k(log_n)
            omega_n = orc.ints_to_mont([w])[0]
            omega_c = orc.ints_to_mont([pow(w, 1 << log_r, pyref.R_MOD)])[0]
            fs = sharded.FourStepNTT(OracleNttEngine(orc), log_n, log_r, rank, world)
            rows = fs.forward(fs.local_columns(a), omega_n, omega_c)
            assert rows.shape == ((1 << log_r) // world, 1 << (log_n - log_r), 4)
            nat = fs.gather_natural(rows)
            assert np.array_equal(nat, orc.best_fft(a, omega_n, log_n)), f"four-step NTT mismatch {log_n},{log_r}"
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_paths_world2_gloo(tmp_path, orc):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_shard_range_partitions():
    from __graft_entry__ import load_package
    import importlib
    sharded = importlib.import_module(load_package().__name__ + ".sharded")
    for n in (0, 1, 7, 8, 1000):
        for world in (1, 2, 3, 8):
            r = [sharded.shard_range(n, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
