"""Formats around the hot path: G1Affine to_bytes / from_bytes, the commitments part of
VerifyingKey::write / read, the library-derived vk.transcript_repr (host code: CPU tests), and ParamsKZG::write / read
with the 2n points (de)compressed on the device (GPU tests).  halo2's own serialisers are third-party code that is not in
/root/reference and cannot run here; the layouts are restated from halo2_proofs v2023_02_02 / halo2curves 0.3.1 [M]."""
import ctypes
import hashlib

import numpy as np
import pytest

from tests.test_verifier import load_golden

Q = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
R = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001


def test_g1_bytes_roundtrip_and_rejects(zk, orc):
    z, vk, inst, proof = load_golden(zk, "mst_k9")
    pts = np.concatenate([vk.fixed, vk.sigma, np.zeros((1, 8), dtype=np.uint64)])          # the last one: identity
    out = np.zeros(32 * pts.shape[0], dtype=np.uint8)
    assert zk.lib().b200zk_g1_to_bytes(zk._p(pts), ctypes.c_size_t(pts.shape[0]), zk._p(out)) == 0
    want = orc.g1_compress(pts)                                                              # the oracle's G1Affine::to_bytes
    assert out.tobytes() == np.asarray(want, dtype=np.uint8).tobytes()
    back = np.zeros_like(pts)
    assert zk.lib().b200zk_g1_from_bytes(zk._p(out), ctypes.c_size_t(pts.shape[0]), zk._p(back)) == 0
    assert np.array_equal(back, pts)
    for x in (Q, Q + 1, 5, 4, 7):                           # x >= q is non-canonical; small x: accepted exactly when x^3 + 3 is a square
        b = np.frombuffer(int(x).to_bytes(32, "little"), dtype=np.uint8).copy()
        rc = zk.lib().b200zk_g1_from_bytes(zk._p(b), ctypes.c_size_t(1), zk._p(back))
        on_curve = x < Q and pow((x ** 3 + 3) % Q, (Q - 1) // 2, Q) == 1
        assert (rc == 0) == on_curve
    b = np.zeros(32, dtype=np.uint8); b[31] = 0x80                                          # sign bit on the identity
    assert zk.lib().b200zk_g1_from_bytes(zk._p(b), ctypes.c_size_t(1), zk._p(back)) == zk.EVERIFY


def test_vk_commitments_roundtrip(zk):
    z, vk, inst, proof = load_golden(zk, "mst_k9")
    data = vk.commitments_to_bytes()
    assert len(data) == 4 + 32 * (vk.fixed.shape[0] + vk.sigma.shape[0])
    assert int.from_bytes(data[:4], "big") == vk.fixed.shape[0]
    fixed, sigma = zk.VerifyingKey.commitments_from_bytes(data, vk.sigma.shape[0])
    assert np.array_equal(fixed, vk.fixed) and np.array_equal(sigma, vk.sigma)
    with pytest.raises(zk.B200zkError):
        zk.VerifyingKey.commitments_from_bytes(data[:-1], vk.sigma.shape[0])


def test_vk_transcript_repr_binds_the_key(zk, orc):
    """The derived transcript_repr is Blake2b-512("Halo2-Verify-Key") over len || blob || commitments, reduced mod r
    (recomputed here with hashlib), and changes with k, the constraint system and every commitment."""
    z, vk, inst, proof = load_golden(zk, "mst_k9")
    got = orc.mont_to_ints(vk.transcript_repr)[0]
    blob = np.ascontiguousarray(vk.blob, dtype="<u4").tobytes()
    body = blob
    for pts in (vk.fixed, vk.sigma):
        for x, y in zip(orc.mont_to_ints(pts[:, :4], orc.FQ), orc.mont_to_ints(pts[:, 4:], orc.FQ)):
            body += x.to_bytes(32, "little") + y.to_bytes(32, "little")
    h = hashlib.blake2b(digest_size=64, person=b"Halo2-Verify-Key")
    h.update(len(body).to_bytes(8, "little") + body)
    assert got == int.from_bytes(h.digest(), "little") % R
    fx = vk.fixed.copy(); fx[1] = vk.g1
    assert not np.array_equal(zk.VerifyingKey(vk.cs, vk.k, fx, vk.sigma, vk.g1, vk.s_g2).transcript_repr, vk.transcript_repr)
    sg = vk.sigma.copy(); sg[0] = vk.g1
    assert not np.array_equal(zk.VerifyingKey(vk.cs, vk.k, vk.fixed, sg, vk.g1, vk.s_g2).transcript_repr, vk.transcript_repr)
    # the golden proof was made with a caller-supplied value: with the derived one it must not verify
    assert vk.verify_proof(inst, proof, z["transcript_repr"]) and not vk.verify_proof(inst, proof)


@pytest.mark.gpu
def test_params_write_read_roundtrip(zk, backend, orc):
    """ParamsKZG::write -> ParamsKZG::read on the device: identical bases, identical G2 elements, identical
    commitments; the G1 part of the byte stream equals the oracle's G1Affine::to_bytes of every base."""
    k = 10
    s = orc.random_fr(1, 31)[0]
    params = zk.ParamsKZG.setup(backend, k, s)
    g2, s_g2 = zk.g2_mul(np.array(zk._FR_ONE, dtype=np.uint64)), zk.g2_mul(s)
    data = params.to_bytes(g2, s_g2)
    n = 1 << k
    assert len(data) == 4 + 2 * n * 32 + 128 and int.from_bytes(data[:4], "little") == k
    g, gl = params.read()
    assert data[4:4 + 32 * n] == np.asarray(orc.g1_compress(g), dtype=np.uint8).tobytes()
    assert data[4 + 32 * n:4 + 64 * n] == np.asarray(orc.g1_compress(gl), dtype=np.uint8).tobytes()
    p2, g2b, s_g2b = zk.ParamsKZG.from_bytes(backend, data)
    g_b, gl_b = p2.read()
    assert np.array_equal(g_b, g) and np.array_equal(gl_b, gl)
    assert np.array_equal(g2b, g2) and np.array_equal(s_g2b, s_g2)
    poly = orc.random_fr(n, 32)
    assert np.array_equal(p2.commit(poly), params.commit(poly)) and np.array_equal(p2.commit_lagrange(poly), params.commit_lagrange(poly))
    bad = bytearray(data); bad[4 + 5 * 32 + 2] ^= 1                       # an x with (almost surely) no curve point, or a different point
    try:
        p3, _, _ = zk.ParamsKZG.from_bytes(backend, bytes(bad))
        g_c, _ = p3.read()
        assert not np.array_equal(g_c[5], g[5])
        p3.close()
    except zk.B200zkError:
        pass
    with pytest.raises(zk.B200zkError):
        zk.ParamsKZG.from_bytes(backend, data[:-1])
    p2.close(); params.close()


@pytest.mark.gpu
def test_prove_and_verify_with_derived_transcript_repr(zk, backend, orc):
    """keygen on the device -> VerifyingKey -> its derived transcript_repr -> create_proof -> verify_proof, all through the
    C ABI with no caller-invented constant; a proof made under one key does not verify under a key with another commitment."""
    import importlib
    synth = importlib.import_module(zk.__name__ + ".circuits_synth")
    from oracle import prover as OP
    job = synth.small(7)
    s = orc.random_fr(1, 41)[0]
    params = zk.ParamsKZG.setup(backend, job.k, s)
    g, _ = params.read(lagrange=False)
    pk = zk.ProvingKey(params, job.cs, job.k, job.fixed, job.map_col, job.map_row)
    fixed_c, sigma_c = pk.vk_commitments()
    vk = zk.VerifyingKey(job.cs, job.k, fixed_c, sigma_c, g[0], zk.g2_mul(s))
    inst = [orc.ints_to_mont([v % OP.R for v in c]) for c in job.instances]
    proof = pk.create_proof(job.advice, inst, orc.XorShiftWide().draw(pk.rng_draws), vk.transcript_repr)
    assert vk.verify_proof(inst, proof)
    fx = fixed_c.copy(); fx[0] = g[1]
    assert not zk.VerifyingKey(job.cs, job.k, fx, sigma_c, g[0], zk.g2_mul(s)).verify_proof(inst, proof)
    pk.close(); params.close()
