"""The product's host verifier (b200zk_verify_proof: plonk::verify_proof + VerifierSHPLONK + the KZG pairing
check) against the committed golden proofs and the oracle — the reference's one hot-path
assertion, `verify_proof(..).is_ok()` (/root/reference/src/circuits/utils.rs:52-63, reached from test_full_prover,
/root/reference/src/circuits/merkle_sum_tree.rs:345-358).  The verifier needs no device, so these run on CPU."""
import hashlib
import importlib
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SHA = {"small_k5": "b1c8af7aee6adfda77aeece7a58c4ddfa17d1f4e6781e4559bc0d50e6c93d302",
       "mst_k9": "98ef5429e8b3cdc6ecdf5804db274602f6f61a870527cabd244ab2fdea633f9d"}


class _Blob:
    """A constraint system known only by its serialised form (what a verifier is handed)."""

    def __init__(self, blob, num_fixed, num_perm):
        self._blob, self.num_fixed, self.permutation = blob, num_fixed, [None] * num_perm

    def to_blob(self, k):
        return self._blob


def load_golden(zk, name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cs = _Blob(z["blob"], z["fixed_commitments"].shape[0], z["sigma_commitments"].shape[0])
    vk = zk.VerifyingKey(cs, int(z["k"]), z["fixed_commitments"], z["sigma_commitments"], z["g1"], z["s_g2"], z["g2"])
    inst, off = [], 0
    for ln in z["instance_lens"]:
        inst.append(z["instances"][off:off + int(ln)]); off += int(ln)
    return z, vk, inst, bytes(z["proof"])


@pytest.mark.parametrize("name", ["small_k5", "mst_k9"])
def test_golden_proof_verifies(zk, orc, name):
    z, vk, inst, proof = load_golden(zk, name)
    assert hashlib.sha256(proof).hexdigest() == SHA[name]
    assert vk.verify_proof(inst, proof, z["transcript_repr"])
    assert vk.verify_proof(inst, proof + b"\x00" * 32, z["transcript_repr"])          # upstream's reader ignores trailing bytes
    # every 32-byte element of the proof matters: flip one bit in each (commitments, evaluations, SHPLONK points)
    step = 1 if name == "small_k5" else 7
    for el in range(0, len(proof) // 32, step):
        bad = bytearray(proof); bad[32 * el + 3] ^= 0x10
        assert not vk.verify_proof(inst, bytes(bad), z["transcript_repr"]), f"element {el}"
    assert not vk.verify_proof(inst, proof[:-32], z["transcript_repr"])                # truncated
    assert not vk.verify_proof(inst, proof[:-1], z["transcript_repr"])
    # wrong public input, wrong transcript_repr, wrong [s]_2, wrong verifying key
    one = np.array(orc.ints_to_mont([1])).reshape(4)
    wrong = [c.copy() for c in inst]
    wrong[0][-1] = one
    assert not vk.verify_proof(wrong, proof, z["transcript_repr"])
    assert not vk.verify_proof(inst, proof, one)
    other = zk.VerifyingKey(vk.cs, vk.k, vk.fixed, vk.sigma, vk.g1, zk.g2_mul(one))
    assert not other.verify_proof(inst, proof, z["transcript_repr"])
    fx = vk.fixed.copy(); fx[0] = vk.g1
    assert not zk.VerifyingKey(vk.cs, vk.k, fx, vk.sigma, vk.g1, vk.s_g2).verify_proof(inst, proof, z["transcript_repr"])
    # non-canonical scalar (r itself) in place of an evaluation, and an x that is not on the curve
    r_bytes = (0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001).to_bytes(32, "little")
    assert not vk.verify_proof(inst, proof[:-96] + r_bytes + proof[-64:], z["transcript_repr"])
    # instance column longer than n - (blinding_factors + 1): Error::InstanceTooLarge
    big = [np.tile(one, ((1 << vk.k), 1))]
    assert not vk.verify_proof(big, proof, z["transcript_repr"])


def test_pairing_and_g2_against_oracle(zk, orc):
    from oracle import pairing as PR
    from oracle import pyref as P

    def g1m(pt):
        return np.array(orc.ints_to_mont([pt[0], pt[1]], which=1)).reshape(8)

    def g2m(q):
        return np.array(orc.ints_to_mont([q[0][0], q[0][1], q[1][0], q[1][1]], which=1)).reshape(16)

    a, b = 123456789, 987654321
    s_g2 = zk.g2_mul(orc.ints_to_mont([b])[0])
    assert np.array_equal(s_g2, g2m(PR.g2_mul(PR.G2_GEN, b)))                          # [b]_2, EIP-197 generator
    assert np.array_equal(zk.g2_mul(orc.ints_to_mont([1])[0]), g2m(PR.G2_GEN))
    r_minus_1 = zk.g2_mul(orc.ints_to_mont([PR.R - 1])[0])
    assert np.array_equal(r_minus_1, g2m(PR.g2_neg(PR.G2_GEN)))                        # r * G2 = identity
    pa, pab, pab1 = (g1m(P.g1_mul(P.G1_GEN, v % PR.R)) for v in (a, -a * b, -a * b + 1))
    assert zk.pairing_check([pa, pab], [s_g2, g2m(PR.G2_GEN)])                         # e(aG, bH) e(-abG, H) = 1
    assert not zk.pairing_check([pa, pab1], [s_g2, g2m(PR.G2_GEN)])
    assert not zk.pairing_check([pa], [s_g2])                                          # non-degenerate
    assert zk.pairing_check([np.zeros(8, dtype=np.uint64)], [s_g2])                    # identity pairs to one
    with pytest.raises(zk.B200zkError):
        zk.pairing_check([pa + np.uint64(1)], [s_g2])                                  # not on the curve


@pytest.mark.parametrize("name,k", [("v3_shaped", 6), ("generic_shapes", 5)])
def test_oracle_proofs_verify(zk, orc, name, k):
    """Fresh oracle proofs of the other circuit shapes (no lookups / 16 quotient cosets + dynamic lookup):
    the product verifier and the oracle's verifier restatement agree on accept and on reject."""
    from oracle import pairing as PR
    from oracle import prover as OP
    gold = importlib.import_module("tests.golden.make_golden")
    synth = importlib.import_module(zk.__name__ + ".circuits_synth")
    job = getattr(synth, name)(k)
    s = orc.random_fr(1, 5)[0]
    g, gl = orc.params_setup(job.k, s)
    pk = OP.keygen_pk(job.cs, job.k, job.fixed, job.map_col, job.map_row)
    wide = orc.XorShiftWide().draw(OP.rng_draws_needed(job.cs, job.k))
    proof, _ = OP.create_proof(g, gl, pk, job.advice, job.instances, wide, job.transcript_repr)
    fixed_c = np.array([gold.g1_mont(OP.commit(g, pk.fixed_polys[c])) for c in range(job.cs.num_fixed)]).reshape(-1, 8)
    sigma_c = np.array([gold.g1_mont(OP.commit(g, p)) for p in pk.perm_polys]).reshape(-1, 8)
    vk = zk.VerifyingKey(job.cs, job.k, fixed_c, sigma_c, np.asarray(g[0]).reshape(8), zk.g2_mul(s))
    inst = [np.array(orc.ints_to_mont([v % PR.R for v in col])).reshape(-1, 4) for col in job.instances]
    tr = np.array(orc.ints_to_mont([job.transcript_repr % PR.R])).reshape(4)
    assert OP.verify_full(orc.mont_to_ints(s)[0], g, pk, job.instances, proof, job.transcript_repr)
    assert vk.verify_proof(inst, proof, tr)
    bad = bytearray(proof); bad[len(proof) // 2] ^= 1
    assert not vk.verify_proof(inst, bytes(bad), tr)


def test_bad_arguments_are_errors_not_verdicts(zk, orc):
    """A malformed constraint-system blob or verifier params are caller errors (B200ZK_EINVAL), not `false`."""
    z, vk, inst, proof = load_golden(zk, "small_k5")
    blob = z["blob"].copy()
    blob[0] ^= 1                                                         # magic word
    broken = zk.VerifyingKey(_Blob(blob, vk.fixed.shape[0], vk.sigma.shape[0]), vk.k, vk.fixed, vk.sigma, vk.g1, vk.s_g2, vk.g2)
    with pytest.raises(zk.B200zkError):
        broken.verify_proof(inst, proof, z["transcript_repr"])
    truncated = zk.VerifyingKey(_Blob(z["blob"][:-3], vk.fixed.shape[0], vk.sigma.shape[0]), vk.k, vk.fixed, vk.sigma, vk.g1, vk.s_g2, vk.g2)
    with pytest.raises(zk.B200zkError):
        truncated.verify_proof(inst, proof, z["transcript_repr"])
    off_curve = vk.s_g2.copy(); off_curve[0] ^= np.uint64(1)
    with pytest.raises(zk.B200zkError):
        zk.VerifyingKey(vk.cs, vk.k, vk.fixed, vk.sigma, vk.g1, off_curve, vk.g2).verify_proof(inst, proof, z["transcript_repr"])
    with pytest.raises(zk.B200zkError):
        zk.g2_mul(z["transcript_repr"], base=off_curve)
    assert zk.VerifyingKey(vk.cs, vk.k, vk.fixed, vk.sigma, vk.g1, vk.s_g2, vk.g2).verify_proof(inst, b"", z["transcript_repr"]) is False


@pytest.mark.parametrize("which", ["less_than", "safe_accumulator", "merkle_v3"])
def test_reference_circuits_verify(zk, orc, which):
    """The other circuits of the reference — LessThan (dynamic lookup, 800 public inputs; /root/reference/src/circuits/
    less_than.rs), SafeAccumulator (degree-17 gates: 16 quotient pieces; circuits/safe_accumulator.rs) and Merkle tree v3
    (circuits/merkle_v3.rs) — proved by the oracle and accepted by the product's verify_proof; a wrong public input is not."""
    from oracle import pairing as PR
    from oracle import prover as OP
    gold = importlib.import_module("tests.golden.make_golden")
    fe = importlib.import_module(zk.__name__ + ".frontend")
    chips = importlib.import_module(zk.__name__ + ".chips")
    if which == "less_than":
        job = fe.synthesize_job(chips.LessThanCircuit(755), 10, [list(range(800))])
    elif which == "safe_accumulator":
        job = fe.synthesize_job(chips.SafeAccumulatorCircuit([1, 3], [0, 0, 14, 13]), 8, [[0, 0, 15, 1]])
    else:
        elements, indices = [3, 5, 8, 13, 21], [0, 1, 0, 1, 1]
        job = fe.synthesize_job(chips.MerkleTreeV3Circuit(99, elements, indices), 10, [[99, chips.compute_merkle_root(99, elements, indices)]])
    s = orc.random_fr(1, 11)[0]
    g, gl = orc.params_setup(job.k, s)
    pk = OP.keygen_pk(job.cs, job.k, job.fixed, job.map_col, job.map_row)
    wide = orc.XorShiftWide().draw(OP.rng_draws_needed(job.cs, job.k))
    proof, _ = OP.create_proof(g, gl, pk, job.advice, job.instances, wide, job.transcript_repr)
    fixed_c = np.array([gold.g1_mont(OP.commit(g, pk.fixed_polys[c])) for c in range(job.cs.num_fixed)]).reshape(-1, 8)
    sigma_c = np.array([gold.g1_mont(OP.commit(g, p)) for p in pk.perm_polys]).reshape(-1, 8)
    vk = zk.VerifyingKey(job.cs, job.k, fixed_c, sigma_c, np.asarray(g[0]).reshape(8), zk.g2_mul(s))
    inst = [np.array(orc.ints_to_mont([v % PR.R for v in col])).reshape(-1, 4) for col in job.instances]
    tr = np.array(orc.ints_to_mont([job.transcript_repr % PR.R])).reshape(4)
    assert vk.verify_proof(inst, proof, tr)
    wrong = [c.copy() for c in inst]
    wrong[0][-1] = np.array(orc.ints_to_mont([12345])).reshape(4)
    assert not vk.verify_proof(wrong, proof, tr)


def test_unsatisfied_witness_is_rejected(zk, orc):
    """A proof made from a witness that breaks a gate is well-formed but must not verify (the expected h(x) check)."""
    from oracle import pairing as PR
    from oracle import prover as OP
    gold = importlib.import_module("tests.golden.make_golden")
    synth = importlib.import_module(zk.__name__ + ".circuits_synth")
    job = synth.small(5)
    adv = np.array(job.advice[2])
    adv[3] = orc.ints_to_mont([5])[0]                                  # breaks the bool gate
    job.advice[2] = adv
    s = orc.random_fr(1, 77)[0]
    g, gl = orc.params_setup(job.k, s)
    pk = OP.keygen_pk(job.cs, job.k, job.fixed, job.map_col, job.map_row)
    wide = orc.XorShiftWide().draw(OP.rng_draws_needed(job.cs, job.k))
    proof, _ = OP.create_proof(g, gl, pk, job.advice, job.instances, wide, job.transcript_repr)
    z, vk, inst, good = load_golden(zk, "small_k5")                      # same circuit, same SRS secret (seed 77)
    assert len(proof) == len(good) and proof != good
    assert vk.verify_proof(inst, good, z["transcript_repr"])
    assert not vk.verify_proof(inst, proof, z["transcript_repr"])
