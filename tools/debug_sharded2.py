#!/usr/bin/env python
"""Debug aid: reproduce the flaky world-8 mismatch and compare the ranks' buffers."""
import importlib, os, sys, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
from oracle import binding as orc
from oracle import prover as OP
zk = load_package()
orc.build(); orc.lib()
chips = importlib.import_module(zk.__name__ + ".chips")
synth = importlib.import_module(zk.__name__ + ".circuits_synth")
job = chips.merkle_sum_tree_job(11, levels=9, seed=5)
s = orc.random_fr(1, 4321)[0]
be = zk.Backend(0)

def single(job):
    params = zk.ParamsKZG.setup(be, job.k, s)
    pk = zk.ProvingKey(params, job.cs, job.k, job.fixed, job.map_col, job.map_row)
    wide = orc.XorShiftWide().draw(pk.rng_draws)
    inst = [orc.ints_to_mont([v % OP.R for v in c]) if len(c) else np.zeros((0, 4), dtype=np.uint64) for c in job.instances]
    tr = orc.ints_to_mont([job.transcript_repr])[0]
    p = pk.create_proof(job.advice, inst, wide, tr)
    pk.close(); params.close()
    return p, wide, inst, tr

def dbg(pk, name, count):
    out = np.zeros((count, 4), dtype=np.uint64)
    c = ctypes.c_size_t()
    pk.backend._check(zk.lib().b200zk_pk_debug_buffer(pk._h, name.encode(), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(count), ctypes.byref(c)))
    return out

def sharded(job, w, wide, inst, tr, want):
    g = zk.Group([0] * w)
    ps = [zk.ParamsKZG.setup(b, job.k, s) for b in g.backends]
    pks = [zk.ProvingKey(p, job.cs, job.k, job.fixed, job.map_col, job.map_row) for p in ps]
    got = g.create_proof(pks, job.advice, inst, wide, tr)
    ok = got == want
    if not ok:
        n = 1 << job.k
        diff = [i // 32 for i in range(0, len(want), 32) if got[i:i + 32] != want[i:i + 32]]
        print("DIFF items", diff[:10], "of", len(want) // 32, flush=True)
        A = job.cs.num_advice
        adv = [dbg(pk, "advice_values", A * n) for pk in pks]
        perm = [dbg(pk, "perm_polys", 4 * n) for pk in pks]
        for r in range(1, w):
            bad_cols = [c for c in range(A) if not np.array_equal(adv[r][c * n:(c + 1) * n], adv[0][c * n:(c + 1) * n])]
            bad_z = [z for z in range(4) if not np.array_equal(perm[r][z * n:(z + 1) * n], perm[0][z * n:(z + 1) * n])]
            print(f"rank {r}: advice cols differing from rank 0: {bad_cols}; perm polys differing: {bad_z}", flush=True)
            for c in bad_cols[:3]:
                rows = np.nonzero((adv[r][c * n:(c + 1) * n] != adv[0][c * n:(c + 1) * n]).any(axis=1))[0]
                print("   col", c, "rows", rows[:8], "...", len(rows), flush=True)
        for c in range(A):
            if not np.array_equal(adv[0][c * n:c * n + n - 6], np.asarray(job.advice[c]).reshape(n, 4)[: n - 6]):
                print("rank 0 advice col", c, "differs from the input", flush=True)
    for p in pks: p.close()
    for p in ps: p.close()
    g.close()
    return ok

others = [(getattr(synth, nm)(k), ws) for nm, k, ws in [("small", 6, (2, 3)), ("mst_shaped", 9, (2, 4, 8)), ("v3_shaped", 8, (2, 5)), ("generic_shapes", 8, (3, 8))]]
want, wide, inst, tr = single(job)
for it in range(12):
    for j, ws in others:
        w0, wd0, in0, tr0 = single(j)
        for w in ws:
            if not sharded(j, w, wd0, in0, tr0, w0): print("other circuit failed", j.k, w, flush=True)
    res = [sharded(job, w, wide, inst, tr, want) for w in (2, 4, 8)]
    print("iter", it, res, flush=True)
    if not all(res):
        break
