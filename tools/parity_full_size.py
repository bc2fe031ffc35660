#!/usr/bin/env python
"""One-off parity check at BASELINE's headline size: the reference's Merkle Sum Tree circuit padded
to k (default 20) is proved by the GPU backend and by the CPU oracle (restatement of halo2's
prover, minutes at this size) from the same SRS, witness and RNG stream; the proof bytes are
compared.  Prints one JSON line (sha256 of both proofs, timings).  Test infrastructure, not the
product:  python tools/parity_full_size.py [k] [mst|mst_dense] [seed]"""
import hashlib
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402
from oracle import binding as orc  # noqa: E402
from oracle import prover as OP  # noqa: E402


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    circuit = sys.argv[2] if len(sys.argv) > 2 else "mst"          # mst | mst_dense (every row in use)
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    zk = load_package()
    chips = importlib.import_module(zk.__name__ + ".chips")
    synth = importlib.import_module(zk.__name__ + ".circuits_synth")
    orc.build()
    job = chips.merkle_sum_tree_job(k, seed=seed) if circuit == "mst" else synth.mst_shaped(k, seed=seed)
    be = zk.Backend(0)
    s = orc.random_fr(1, 20251018)[0]
    t0 = time.perf_counter()
    params = zk.ParamsKZG.setup(be, k, s)
    g, gl = params.read()
    pk = zk.ProvingKey(params, job.cs, k, job.fixed, job.map_col, job.map_row)
    wide = orc.XorShiftWide().draw(pk.rng_draws)
    inst = [orc.ints_to_mont([v % OP.R for v in c]) for c in job.instances]
    t1 = time.perf_counter()
    got = pk.create_proof(job.advice, inst, wide, orc.ints_to_mont([job.transcript_repr])[0])
    t2 = time.perf_counter()
    pk.close(); params.close(); be.close()
    opk = OP.keygen_pk(job.cs, k, job.fixed, job.map_col, job.map_row)
    t3 = time.perf_counter()
    want, _ = OP.create_proof(g, gl, opk, job.advice, job.instances, wide, job.transcript_repr)
    t4 = time.perf_counter()
    first = next((i // 32 for i in range(0, len(want), 32) if got[i:i + 32] != want[i:i + 32]), None)
    print(json.dumps({"circuit": "MerkleSumTreeCircuit, 16-level path" if circuit == "mst" else "MST-shaped synthetic circuit, dense witness", "k": k, "seed": seed, "proof_bytes": len(got),
                      "gpu_sha256": hashlib.sha256(got).hexdigest(), "oracle_sha256": hashlib.sha256(want).hexdigest(),
                      "identical": got == want, "first_differing_item": first,
                      "gpu_setup_keygen_s": round(t1 - t0, 2), "gpu_create_proof_e2e_s": round(t2 - t1, 3),
                      "oracle_keygen_s": round(t3 - t2, 1), "oracle_create_proof_s": round(t4 - t3, 1),
                      "oracle_threads": orc.get_threads()}), flush=True)


if __name__ == "__main__":
    main()
