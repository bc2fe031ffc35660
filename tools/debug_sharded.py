#!/usr/bin/env python
"""Debug aid: sharded vs single-GPU proof of the MST circuit, first differing 32-byte item per world size."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
from oracle import binding as orc
from oracle import prover as OP
zk = load_package()
orc.build(); orc.lib()
chips = importlib.import_module(zk.__name__ + ".chips")
k = int(sys.argv[1]) if len(sys.argv) > 1 else 11
job = chips.merkle_sum_tree_job(k, levels=9, seed=5)
s = orc.random_fr(1, 4321)[0]
be = zk.Backend(0)
params = zk.ParamsKZG.setup(be, job.k, s)
pk = zk.ProvingKey(params, job.cs, job.k, job.fixed, job.map_col, job.map_row)
wide = orc.XorShiftWide().draw(pk.rng_draws)
inst = [orc.ints_to_mont([v % OP.R for v in c]) for c in job.instances]
tr = orc.ints_to_mont([job.transcript_repr])[0]
want = pk.create_proof(job.advice, inst, wide, tr)
worlds = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "8,8,5,6,7,3,8".split(","))]
for w in worlds:
    g = zk.Group([0] * w)
    ps = [zk.ParamsKZG.setup(b, job.k, s) for b in g.backends]
    pks = [zk.ProvingKey(p, job.cs, job.k, job.fixed, job.map_col, job.map_row) for p in ps]
    for rep in range(2):
        got = g.create_proof(pks, job.advice, inst, wide, tr)
        diff = [i // 32 for i in range(0, len(want), 32) if got[i:i + 32] != want[i:i + 32]]
        print(f"world {w} rep {rep}: {'OK' if got == want else 'DIFF first items ' + str(diff[:6]) + ' of ' + str(len(want) // 32)}", flush=True)
    for p in pks: p.close()
    for p in ps: p.close()
    g.close()
