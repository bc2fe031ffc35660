"""CPU: the oracle prover (restatement of halo2 create_proof) emits proofs that the independent
restatement of the verifier accepts, and tampering is rejected.  Mirrors the only hot-path test
of the reference, test_full_prover (/root/reference/src/circuits/merkle_sum_tree.rs:345-358,
assertion at /root/reference/src/circuits/utils.rs:56-63: verify_proof(..).is_ok())."""
import numpy as np
import pytest

from oracle import prover as OP
from oracle import pyref as P


def _job(zk, name, k):
    import importlib
    synth = importlib.import_module(zk.__name__ + ".circuits_synth")
    return getattr(synth, name)(k)


def _prove(orc, job, seed_s=77):
    s = orc.random_fr(1, seed_s)[0]
    g, gl = orc.params_setup(job.k, s)
    pk = OP.keygen_pk(job.cs, job.k, job.fixed, job.map_col, job.map_row)
    wide = orc.XorShiftWide().draw(OP.rng_draws_needed(job.cs, job.k))
    proof, trace = OP.create_proof(g, gl, pk, job.advice, job.instances, wide, job.transcript_repr)
    return s, g, pk, proof, trace


@pytest.mark.parametrize("name,k", [("small", 5), ("v3_shaped", 6), ("small", 7), ("generic_shapes", 5)])
def test_oracle_proof_verifies(zk, orc, name, k):
    job = _job(zk, name, k)
    cs = job.cs
    assert cs.blinding_factors() == 5 and cs.degree() == (17 if name == "generic_shapes" else 6)
    s, g, pk, proof, trace = _prove(orc, job)
    A, L, S, q = cs.num_advice, len(cs.lookups), cs.num_permutation_sets(), cs.degree() - 1
    n_evals = len(cs.advice_queries) + len(cs.fixed_queries) + 1 + len(cs.permutation) + (3 * S - 1) + 5 * L
    assert len(proof) == 32 * (A + 2 * L + S + L + 1 + q + n_evals + 2)
    assert OP.verify_full(orc.mont_to_ints(s)[0], g, pk, job.instances, proof, job.transcript_repr)
    # deterministic
    _, _, _, proof2, _ = _prove(orc, job)
    assert proof2 == proof
    # tampered proof / wrong public input are rejected
    bad = bytearray(proof)
    off = 32 * (A + 2 * L + S + L + 1 + q)          # first advice evaluation
    bad[off] ^= 1
    ok = True
    try:
        ok = OP.verify_full(orc.mont_to_ints(s)[0], g, pk, job.instances, bytes(bad), job.transcript_repr)
    except AssertionError:
        ok = False
    assert not ok
    wrong = [[v + 1 for v in job.instances[0]]]
    assert not OP.verify_full(orc.mont_to_ints(s)[0], g, pk, wrong, proof, job.transcript_repr)


def test_mst_shape_counts(zk):
    """The synthetic job has the MST shape of SURVEY.md §8: A=20, L=8, P=16, d=6, bf=5, S=4."""
    job = _job(zk, "mst_shaped", 6)
    cs = job.cs
    assert (cs.num_advice, len(cs.lookups), len(cs.permutation)) == (20, 8, 16)
    assert (cs.degree(), cs.blinding_factors(), cs.num_permutation_sets()) == (6, 5, 4)
    blob = cs.to_blob(6)
    assert blob[0] == 0x324B5A42 and blob[2] == 6 and blob[3] == 20


def test_unsatisfied_witness_is_rejected(zk, orc):
    job = _job(zk, "small", 5)
    adv = np.array(job.advice[2])
    adv[3] = orc.ints_to_mont([5])[0]                # breaks the bool gate
    job.advice[2] = adv
    s, g, pk, proof, _ = _prove(orc, job)
    ok = True
    try:
        ok = OP.verify_full(orc.mont_to_ints(s)[0], g, pk, job.instances, proof, job.transcript_repr)
    except AssertionError:
        ok = False
    assert not ok


def test_permute_expression_pair_cpp_matches_python(orc):
    import random
    rnd = random.Random(3)
    u = 300
    table = [rnd.randrange(40) for _ in range(u)]
    inp = [rnd.choice(table) for _ in range(u)]
    I_, T_ = orc.ints_to_mont(inp), orc.ints_to_mont(table)
    a1, s1 = orc.permute_expression_pair(I_, T_, u)
    a2, s2 = OP.permute_expression_pair_py(I_, T_, u)
    assert np.array_equal(a1, a2) and np.array_equal(s1, s2)
    assert sorted(orc.mont_to_ints(s1)) == sorted(table) and orc.mont_to_ints(a1) == sorted(inp)
    with pytest.raises(ValueError):
        orc.permute_expression_pair(orc.ints_to_mont([1, 2, 99]), orc.ints_to_mont([1, 2, 3]), 3)


def test_frontend_montgomery_conversion(zk, orc):
    """circuits_synth.mont_from_u64 (vectorised, no field multiplication) == x * 2^256 mod r."""
    import importlib
    synth = importlib.import_module(zk.__name__ + ".circuits_synth")
    rng = np.random.Generator(np.random.PCG64(1))
    vals = np.concatenate([rng.integers(0, 1 << 63, size=3000, dtype=np.int64), np.array([0, 1, 65535, 65536, (1 << 62) + 12345, (1 << 63) - 1])])
    got = synth.mont_from_u64(vals)
    assert np.array_equal(got, orc.ints_to_mont([int(v) for v in vals]))
    assert np.array_equal(synth.mont_from_small(vals), got)


@pytest.mark.parametrize("name,k", [("small", 5), ("v3_shaped", 6), ("mst_shaped", 7), ("generic_shapes", 6)])
def test_cpp_prover_matches_python_prover(zk, orc, name, k):
    """oracle/prover.cpp (the threaded restatement the CPU arm of the bench and the full-size parity checks run)
    emits exactly the bytes of oracle/prover.py (the readable restatement the verifier restatement is tested on)."""
    job = _job(zk, name, k)
    s, g, pk, want, _ = _prove(orc, job)
    _, gl = orc.params_setup(job.k, s)
    cpk = orc.CppProvingKey(job.cs, job.k, job.fixed, job.map_col, job.map_row)
    assert cpk.rng_draws == OP.rng_draws_needed(job.cs, job.k)
    wide = orc.XorShiftWide().draw(cpk.rng_draws)
    got = cpk.create_proof(g, gl, job.advice, job.instances, wide, job.transcript_repr)
    cpk.close()
    assert got == want


def test_cpp_prover_real_circuits(zk, orc):
    """The reference's circuits through both restatements: Merkle Sum Tree (test_full_prover's instance, k = 9),
    LessThan (dynamic lookup, k = 10), SafeAccumulator (16 quotient cosets, k = 8); and the failure path."""
    import importlib
    fe = importlib.import_module(zk.__name__ + ".frontend")
    chips = importlib.import_module(zk.__name__ + ".chips")
    leaf, elements, indices = (10, 100), [(1, 10), (5, 50), (6, 60), (9, 90), (9, 90)], [0] * 5
    root = chips.compute_merkle_sum_root(leaf, elements, indices)
    circuit = chips.MerkleSumTreeCircuit(leaf[0], leaf[1], [e[0] for e in elements], [e[1] for e in elements], indices, 500)
    jobs = [fe.synthesize_job(circuit, 9, [[leaf[0], leaf[1], root[0], 500]]),
            fe.synthesize_job(chips.LessThanCircuit(755), 10, [list(range(800))]),
            fe.synthesize_job(chips.SafeAccumulatorCircuit([1, 3], [0, 0, 14, 13]), 8, [[0, 0, 15, 1]])]
    for job in jobs:
        s, g, pk, want, _ = _prove(orc, job)
        _, gl = orc.params_setup(job.k, s)
        cpk = orc.CppProvingKey(job.cs, job.k, job.fixed, job.map_col, job.map_row)
        wide = orc.XorShiftWide().draw(cpk.rng_draws)
        assert cpk.create_proof(g, gl, job.advice, job.instances, wide, job.transcript_repr) == want
        cpk.close()
    synth = importlib.import_module(zk.__name__ + ".circuits_synth")
    job = synth.small(6)
    bad = np.array(job.advice[5 + 3 + 2])
    bad[3] = orc.ints_to_mont([100000])[0]
    job.advice[5 + 3 + 2] = bad
    s = orc.random_fr(1, 77)[0]
    g, gl = orc.params_setup(job.k, s)
    cpk = orc.CppProvingKey(job.cs, job.k, job.fixed, job.map_col, job.map_row)
    with pytest.raises(ValueError, match="ConstraintSystemFailure"):
        cpk.create_proof(g, gl, job.advice, job.instances, orc.XorShiftWide().draw(cpk.rng_draws), job.transcript_repr)
    cpk.close()
