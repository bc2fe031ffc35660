"""Generates tests/golden/full_size_sha256.json: sha256 of the ORACLE's proof (oracle/prover.cpp, the threaded C++
restatement of halo2's CPU prover; test infrastructure) for the benchmark's full-size configurations:

    python tests/golden/make_full_size_digests.py [name ...]        # default: all

    mst_k20        the reference's MerkleSumTreeCircuit (16-level path) padded to k = 20 — BASELINE's headline config
    mst_dense_k20  the MST-shaped circuit with every row in use, k = 20
    mst_k21        the same Merkle Sum Tree circuit padded to k = 21

Inputs are seeded: job seed 1, ParamsKZG::setup secret s = random_fr(seed 777), rng stream = XorShiftRng with halo2's
customary test seed (oracle.binding.XorShiftWide).  tests/test_gpu_full_size.py builds the same inputs, proves on the
GPU through the C ABI and compares the digest.  One k = 20 proof is ~1.5-3 minutes on 8 cores; k = 21 needs ~40 GB."""
import hashlib
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "full_size_sha256.json")

from __graft_entry__ import load_package  # noqa: E402
from oracle import binding as orc  # noqa: E402

SEED_S = 777
CONFIGS = {"mst_k20": ("chips", "merkle_sum_tree_job", 20), "mst_dense_k20": ("circuits_synth", "mst_shaped", 20),
           "mst_k21": ("chips", "merkle_sum_tree_job", 21)}


def build_job(zk, name):
    mod, fn, k = CONFIGS[name]
    return getattr(importlib.import_module(zk.__name__ + "." + mod), fn)(k, seed=1)


def main():
    zk = load_package()
    orc.build(); orc.lib()
    names = sys.argv[1:] or list(CONFIGS)
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in names:
        t0 = time.time()
        job = build_job(zk, name)
        g, gl = orc.params_setup(job.k, orc.random_fr(1, SEED_S)[0])
        pk = orc.CppProvingKey(job.cs, job.k, job.fixed, job.map_col, job.map_row)
        wide = orc.XorShiftWide().draw(pk.rng_draws)
        t1 = time.time()
        proof = pk.create_proof(g, gl, job.advice, job.instances, wide, job.transcript_repr)
        pk.close()
        res[name] = {"k": job.k, "seed_s": SEED_S, "job_seed": 1, "rng": "XorShiftWide (halo2 test seed)", "proof_bytes": len(proof),
                     "sha256": hashlib.sha256(proof).hexdigest(), "prover": "oracle/prover.cpp", "threads": orc.get_threads(),
                     "keygen_s": round(t1 - t0, 1), "prove_s": round(time.time() - t1, 1)}
        print(name, res[name], flush=True)
        json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
