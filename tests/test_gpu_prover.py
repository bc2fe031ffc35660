"""GPU: create_proof through the C ABI is byte-identical to the oracle's restatement of halo2's
CPU prover on the same circuit, witness, SRS and RNG stream, and the proof verifies (the
reference's own assertion, /root/reference/src/circuits/utils.rs:56-63)."""
import importlib

import numpy as np
import pytest

from oracle import prover as OP

pytestmark = pytest.mark.gpu


def _synth(zk):
    return importlib.import_module(zk.__name__ + ".circuits_synth")


def _mont_ints(orc, vals):
    return orc.ints_to_mont([v % OP.R for v in vals]) if len(vals) else np.zeros((0, 4), dtype=np.uint64)


def _first_diff(a, b):
    for i in range(0, min(len(a), len(b)), 32):
        if a[i:i + 32] != b[i:i + 32]:
            return i // 32
    return None


def _run(zk, backend, orc, job, check_verify=True, pairing=False):
    s = orc.random_fr(1, 4321)[0]
    params = zk.ParamsKZG.setup(backend, job.k, s)
    g, gl = params.read()
    pk_gpu = zk.ProvingKey(params, job.cs, job.k, job.fixed, job.map_col, job.map_row)
    assert pk_gpu.blinding_factors == job.cs.blinding_factors() and pk_gpu.degree == job.cs.degree()
    assert pk_gpu.rng_draws == OP.rng_draws_needed(job.cs, job.k)
    wide = orc.XorShiftWide().draw(pk_gpu.rng_draws)
    inst = [_mont_ints(orc, c) for c in job.instances]
    tr_repr = orc.ints_to_mont([job.transcript_repr])[0]
    got = pk_gpu.create_proof(job.advice, inst, wide, tr_repr)
    pk_cpu = OP.keygen_pk(job.cs, job.k, job.fixed, job.map_col, job.map_row)
    want, _ = OP.create_proof(g, gl, pk_cpu, job.advice, job.instances, wide, job.transcript_repr)
    assert len(got) == len(want) == pk_gpu.proof_size
    assert got == want, f"first differing 32-byte item: {_first_diff(got, want)} of {len(want) // 32}"
    if check_verify:
        assert OP.verify_full(orc.mont_to_ints(s)[0], g, pk_cpu, job.instances, got, job.transcript_repr)
    if pairing:                                 # as halo2's verifier: e(., [s]_2) from the SRS, no secret
        from oracle import pairing as PR
        s_g2 = PR.g2_mul(PR.G2_GEN, orc.mont_to_ints(s)[0])
        assert OP.verify_full(None, g, pk_cpu, job.instances, got, job.transcript_repr, s_g2=s_g2)
    # the product's own verifier (b200zk_verify_proof; verifying-key commitments computed on the device)
    fixed_c, sigma_c = pk_gpu.vk_commitments()
    vk = zk.VerifyingKey(job.cs, job.k, fixed_c, sigma_c, g[0], zk.g2_mul(s))
    assert vk.verify_proof(inst, got, tr_repr)
    bad = bytearray(got); bad[len(got) // 2] ^= 1
    assert not vk.verify_proof(inst, bytes(bad), tr_repr)
    # device-resident entry point gives the same bytes
    d_adv = backend.to_device(np.concatenate([np.ascontiguousarray(a).reshape(-1, 4) for a in job.advice]))
    d_wide = backend.to_device(wide)
    assert pk_gpu.create_proof_dev(d_adv, inst, d_wide, tr_repr) == want
    d_adv.free(); d_wide.free()
    pk_gpu.close(); params.close()
    return got


@pytest.mark.parametrize("name,k", [("small", 5), ("small", 6), ("v3_shaped", 6), ("small", 8), ("mst_shaped", 9), ("v3_shaped", 11),
                                    ("generic_shapes", 6), ("generic_shapes", 10)])
def test_proof_bytes_match_oracle(zk, backend, orc, name, k):
    job = getattr(_synth(zk), name)(k)
    _run(zk, backend, orc, job, check_verify=(k <= 9))


def test_mst_shaped_k12(zk, backend, orc):
    _run(zk, backend, orc, _synth(zk).mst_shaped(12), check_verify=False)


def test_lookup_failure_is_reported(zk, backend, orc):
    """An input value missing from the table is Error::ConstraintSystemFailure upstream."""
    synth = _synth(zk)
    job = synth.small(6)
    byte_col = 5 + 3 + 2
    bad = np.array(job.advice[byte_col])
    bad[3] = orc.ints_to_mont([100000])[0]
    job.advice[byte_col] = bad
    s = orc.random_fr(1, 4321)[0]
    params = zk.ParamsKZG.setup(backend, job.k, s)
    pk = zk.ProvingKey(params, job.cs, job.k, job.fixed, job.map_col, job.map_row)
    wide = orc.XorShiftWide().draw(pk.rng_draws)
    with pytest.raises(zk.B200zkError, match="-5"):
        pk.create_proof(job.advice, [_mont_ints(orc, c) for c in job.instances], wide, orc.ints_to_mont([job.transcript_repr])[0])
    with pytest.raises(Exception):          # the oracle rejects the same witness
        g, gl = params.read()
        pk_cpu = OP.keygen_pk(job.cs, job.k, job.fixed, job.map_col, job.map_row)
        OP.create_proof(g, gl, pk_cpu, job.advice, job.instances, wide, job.transcript_repr)
    pk.close(); params.close()


def _frontend(zk):
    return importlib.import_module(zk.__name__ + ".frontend"), importlib.import_module(zk.__name__ + ".chips")


def test_real_merkle_sum_tree_k9(zk, backend, orc):
    """BASELINE config 1 on the GPU: the reference's MerkleSumTreeCircuit (Poseidon width 5, LtChip,
    u8 lookups) with test_full_prover's inputs at k = 9, proof bytes identical to the oracle and
    accepted by the verifier restatement."""
    fe, chips = _frontend(zk)
    leaf, elements, indices = (10, 100), [(1, 10), (5, 50), (6, 60), (9, 90), (9, 90)], [0] * 5
    root = chips.compute_merkle_sum_root(leaf, elements, indices)
    circuit = chips.MerkleSumTreeCircuit(leaf[0], leaf[1], [e[0] for e in elements], [e[1] for e in elements], indices, 500)
    job = fe.synthesize_job(circuit, 9, [[leaf[0], leaf[1], root[0], 500]])
    _run(zk, backend, orc, job, check_verify=True, pairing=True)


def test_real_merkle_sum_tree_k11(zk, backend, orc):
    """Same circuit with a 9-level path at k = 11: large enough for the batched commits, the
    constant-run columns (grand products equal to one value on every unused row) and the warp-level
    NTT to be the code that runs; proof bytes identical to the oracle."""
    fe, chips = _frontend(zk)
    job = chips.merkle_sum_tree_job(11, levels=9, seed=5)
    assert fe.mock_verify(job, rows=range(job.rows_used + 2)) == []
    _run(zk, backend, orc, job, check_verify=False)


@pytest.mark.parametrize("k,levels", [(10, 5), (14, 13)])
def test_real_merkle_v3(zk, backend, orc, k, levels):
    """BASELINE config 2: Poseidon Merkle tree v3 full prove (k = 14), bytes identical to the CPU path."""
    fe, chips = _frontend(zk)
    rng = np.random.Generator(np.random.PCG64(k))
    elements = [int(x) for x in rng.integers(1, 1 << 62, size=levels)]
    indices = [int(x) for x in rng.integers(0, 2, size=levels)]
    root = chips.compute_merkle_root(99, elements, indices)
    job = fe.synthesize_job(chips.MerkleTreeV3Circuit(99, elements, indices), k, [[99, root]])
    _run(zk, backend, orc, job, check_verify=(k <= 10))


def test_real_merkle_sum_tree_k20_verifies(zk, backend, orc):
    """BASELINE's headline size: the Merkle Sum Tree circuit (16-level path) padded to k = 20 is
    proved on the GPU and the proof is accepted by the verifier restatement (size-independent
    property: the oracle prover does not run at this size in test time).  The verifying key's
    commitments are commit_lagrange of the fixed columns and of the permutation polynomials."""
    fe, chips = _frontend(zk)
    k = 20
    job = chips.merkle_sum_tree_job(k)
    s = orc.random_fr(1, 777)[0]
    params = zk.ParamsKZG.setup(backend, k, s)
    pk = zk.ProvingKey(params, job.cs, k, job.fixed, job.map_col, job.map_row)
    wide = orc.XorShiftWide().draw(pk.rng_draws)
    inst = [_mont_ints(orc, c) for c in job.instances]
    proof = pk.create_proof(job.advice, inst, wide, orc.ints_to_mont([job.transcript_repr])[0])
    assert len(proof) == pk.proof_size
    vk_fixed, vk_sigma = pk.vk_commitments()
    pk.close()
    # full-size `verify_proof(..).is_ok()` by pairing, through the C ABI (/root/reference/src/circuits/utils.rs:56-63)
    pvk = zk.VerifyingKey(job.cs, k, vk_fixed, vk_sigma, orc.g1_generator()[:8], zk.g2_mul(s))
    assert pvk.verify_proof(inst, proof, orc.ints_to_mont([job.transcript_repr])[0])
    inst_wrong = [c.copy() for c in inst]
    inst_wrong[0][3] = inst[0][0]
    assert not pvk.verify_proof(inst_wrong, proof, orc.ints_to_mont([job.transcript_repr])[0])

    def aff(out12):
        return orc.affine_to_ints(np.asarray(out12[:8]).reshape(1, 8))[0]

    fixed_c = [aff(params.commit_lagrange(f)) for f in job.fixed]
    dom = orc.Domain(job.cs.degree(), k)
    sigma_c = [aff(params.commit_lagrange(sg)) for sg in OP.sigma_from_mapping(dom, job.map_col, job.map_row)]
    params.close()
    assert fixed_c == [orc.affine_to_ints(c.reshape(1, 8))[0] for c in vk_fixed]
    assert sigma_c == [orc.affine_to_ints(c.reshape(1, 8))[0] for c in vk_sigma]
    vk = OP.verifying_key(job.cs, k, fixed_c, sigma_c)
    s_int = orc.mont_to_ints(s)[0]
    assert OP.verify_full(s_int, None, vk, job.instances, proof, job.transcript_repr)
    wrong = [list(job.instances[0])]
    wrong[0][3] += 1                                  # assets_sum
    assert not OP.verify_full(s_int, None, vk, wrong, proof, job.transcript_repr)


def test_real_less_than_and_safe_accumulator(zk, backend, orc):
    """The other circuits the backend must prove unchanged: LessThan (dynamic lookup into an advice
    table, 800 public inputs, k = 10) and SafeAccumulator (degree-17 range-check gates, k = 8)."""
    fe, chips = _frontend(zk)
    _run(zk, backend, orc, fe.synthesize_job(chips.LessThanCircuit(755), 10, [list(range(800))]), check_verify=True)
    _run(zk, backend, orc, fe.synthesize_job(chips.SafeAccumulatorCircuit([1, 3], [0, 0, 14, 13]), 8, [[0, 0, 15, 1]]), check_verify=True)


@pytest.mark.parametrize("name", ["small_k5", "mst_k9"])
def test_golden_vectors(zk, backend, orc, name):
    """The committed golden vectors (tests/golden/, made by make_golden.py from the oracle): the GPU's
    verifying-key commitments and proof bytes for the same job, SRS secret and rng stream are the committed ones."""
    import hashlib
    gold = importlib.import_module("tests.golden.make_golden")
    tv = importlib.import_module("tests.test_verifier")
    z, vk, inst, want = tv.load_golden(zk, name)
    job = _synth(zk).small(5) if name == "small_k5" else gold.mst_k9_job(zk)
    assert np.array_equal(np.asarray(job.cs.to_blob(job.k), dtype=np.uint32), z["blob"])
    s = orc.random_fr(1, int(z["seed_s"]))[0]
    params = zk.ParamsKZG.setup(backend, job.k, s)
    pk = zk.ProvingKey(params, job.cs, job.k, job.fixed, job.map_col, job.map_row)
    fixed_c, sigma_c = pk.vk_commitments()
    assert np.array_equal(fixed_c, z["fixed_commitments"]) and np.array_equal(sigma_c, z["sigma_commitments"])
    assert np.array_equal(zk.g2_mul(s), z["s_g2"])
    wide = orc.XorShiftWide().draw(pk.rng_draws)
    got = pk.create_proof(job.advice, inst, wide, z["transcript_repr"])
    assert got == want and hashlib.sha256(got).hexdigest() == tv.SHA[name]
    assert vk.verify_proof(inst, got, z["transcript_repr"])
    pk.close(); params.close()
