"""Generates the golden vectors of tests/golden/ from the CPU oracle (test infrastructure).

    python tests/golden/make_golden.py

mst_k9.npz   — the reference's `test_full_prover` instance (/root/reference/src/circuits/merkle_sum_tree.rs:345-358:
               Merkle Sum Tree, k = 9, public inputs [10, 100, root, 500]) proved by oracle/prover.py with
               ParamsKZG::setup secret s = random_fr(seed 2024) and the XorShift rng stream of SURVEY.md 8(c).
small_k5.npz — the small synthetic circuit (gates + lookup + permutation) at k = 5, s = random_fr(seed 77).

Each file holds everything verify_proof needs (constraint-system blob, verifying-key commitments, verifier
params, instances, transcript_repr) and the oracle's proof bytes, so that (a) the product's host verifier
(b200zk_verify_proof) is checked on a CPU-only box against a proof it did not produce, and (b) the GPU prover's
bytes are compared with the committed ones.  The reference itself cannot run here (no Rust toolchain), so
these are oracle outputs, not reference outputs: parity stays "unpinned" in the sense of DESIGN.md §5.
"""
import hashlib
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import binding as orc  # noqa: E402
from oracle import pairing as PR  # noqa: E402
from oracle import prover as OP  # noqa: E402

FQ = 1


def g1_mont(pt):
    return np.zeros(8, dtype=np.uint64) if pt is None else np.array(orc.ints_to_mont([pt[0], pt[1]], which=FQ)).reshape(8)


def g2_mont(q):
    return np.array(orc.ints_to_mont([q[0][0], q[0][1], q[1][0], q[1][1]], which=FQ)).reshape(16)


def emit(name, job, seed_s):
    s = orc.random_fr(1, seed_s)[0]
    g, gl = orc.params_setup(job.k, s)
    pk = OP.keygen_pk(job.cs, job.k, job.fixed, job.map_col, job.map_row)
    wide = orc.XorShiftWide().draw(OP.rng_draws_needed(job.cs, job.k))
    proof, _ = OP.create_proof(g, gl, pk, job.advice, job.instances, wide, job.transcript_repr)
    s_int = orc.mont_to_ints(s)[0]
    assert OP.verify_full(s_int, g, pk, job.instances, proof, job.transcript_repr)
    fixed_c = np.array([g1_mont(OP.commit(g, pk.fixed_polys[c])) for c in range(job.cs.num_fixed)]).reshape(-1, 8)
    sigma_c = np.array([g1_mont(OP.commit(g, p)) for p in pk.perm_polys]).reshape(-1, 8)
    inst = [np.array(orc.ints_to_mont([v % PR.R for v in col])).reshape(-1, 4) for col in job.instances]
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        k=np.uint32(job.k), seed_s=np.uint32(seed_s), blob=np.asarray(job.cs.to_blob(job.k), dtype=np.uint32),
        fixed_commitments=fixed_c, sigma_commitments=sigma_c, g1=np.asarray(g[0]).reshape(8),
        g2=g2_mont(PR.G2_GEN), s_g2=g2_mont(PR.g2_mul(PR.G2_GEN, s_int)),
        instance_lens=np.array([c.shape[0] for c in inst], dtype=np.uint32),
        instances=np.concatenate(inst) if inst else np.zeros((0, 4), dtype=np.uint64),
        transcript_repr=np.array(orc.ints_to_mont([job.transcript_repr % PR.R])).reshape(4),
        proof=np.frombuffer(proof, dtype=np.uint8))
    print(name, "k", job.k, "proof bytes", len(proof), "sha256", hashlib.sha256(proof).hexdigest())


def mst_k9_job(zk):
    fe = importlib.import_module(zk.__name__ + ".frontend")
    chips = importlib.import_module(zk.__name__ + ".chips")
    leaf, elements, indices = (10, 100), [(1, 10), (5, 50), (6, 60), (9, 90), (9, 90)], [0] * 5
    root = chips.compute_merkle_sum_root(leaf, elements, indices)
    circuit = chips.MerkleSumTreeCircuit(leaf[0], leaf[1], [e[0] for e in elements], [e[1] for e in elements], indices, 500)
    return fe.synthesize_job(circuit, 9, [[leaf[0], leaf[1], root[0], 500]])


if __name__ == "__main__":
    from __graft_entry__ import load_package
    zk = load_package()
    orc.build(); orc.lib()
    synth = importlib.import_module(zk.__name__ + ".circuits_synth")
    emit("small_k5", synth.small(5), 77)
    emit("mst_k9", mst_k9_job(zk), 2024)
