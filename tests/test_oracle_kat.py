"""Pins the C++ oracle against independent Python big-int KATs (SURVEY.md §8(c) [P] values)
and against oracle/pyref.py.  CPU only."""
import hashlib
import random

import numpy as np

from oracle import pyref as P

R, Q = P.R_MOD, P.Q_MOD


def test_field_constants(orc):
    fr, fq = orc.field_params(orc.FR), orc.field_params(orc.FQ)
    assert orc.raw_to_ints(fr["p"])[0] == R and orc.raw_to_ints(fq["p"])[0] == Q
    assert fr["inv"] == 0xc2e1f593efffffff and fq["inv"] == 0x87d20782e4866389
    assert orc.raw_to_ints(fr["r"])[0] == 0x0e0a77c19a07df2f666ea36f7879462e36fc76959f60cd29ac96341c4ffffffb
    assert orc.raw_to_ints(fr["r2"])[0] == 0x0216d0b17f4e44a58c49833d53bb808553fe3ab1e35c59e31bb8e645ae216da7
    assert orc.raw_to_ints(fq["r"])[0] == 0x0e0a77c19a07df2f666ea36f7879462c0a78eb28f5c70b3dd35d438dc58f0d9d
    assert orc.raw_to_ints(fq["r2"])[0] == 0x06d89f71cab8351f47ab1eff0a417ff6b5e71911d44501fbf32cfc5b538afa89
    root, delta, zeta = orc.fr_constants()
    assert orc.mont_to_ints(root)[0] == 0x03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c == P.FR_ROOT_OF_UNITY
    assert orc.mont_to_ints(delta)[0] == 0x09226b6e22c6f0ca64ec26aad4c86e715b5f898e5e963f25870e56bbe533e9a2 == P.FR_DELTA
    z = orc.mont_to_ints(zeta)[0]
    assert z == P.FR_ZETA and pow(z, 3, R) == 1 and z != 1


def test_field_ops_vs_bigint(orc):
    rnd = random.Random(1)
    for which, p in ((orc.FR, R), (orc.FQ, Q)):
        edge = [0, 1, 2, p - 1, p - 2, (1 << 253), p >> 1]
        a = edge + [rnd.randrange(p) for _ in range(200)]
        b = list(reversed(edge)) + [rnd.randrange(p) for _ in range(200)]
        A, B = orc.ints_to_mont(a, which), orc.ints_to_mont(b, which)
        assert orc.mont_to_ints(A, which) == a
        assert orc.raw_to_ints(A) == [P.to_mont(x, p) for x in a]          # layout = x*2^256 mod p
        assert orc.mont_to_ints(orc.binop("add", A, B, which), which) == [(x + y) % p for x, y in zip(a, b)]
        assert orc.mont_to_ints(orc.binop("sub", A, B, which), which) == [(x - y) % p for x, y in zip(a, b)]
        assert orc.mont_to_ints(orc.binop("mul", A, B, which), which) == [(x * y) % p for x, y in zip(a, b)]
        nz = [x for x in a if x]
        assert orc.mont_to_ints(orc.inv(orc.ints_to_mont(nz, which), which), which) == [pow(x, -1, p) for x in nz]
    # batch inversion keeps zeros
    v = [0, 5, 0, 7, R - 1]
    assert orc.mont_to_ints(orc.batch_invert(orc.ints_to_mont(v))) == [0 if x == 0 else pow(x, -1, R) for x in v]


def test_from_u512(orc):
    rnd = random.Random(2)
    wide = [rnd.getrandbits(512) for _ in range(50)] + [0, (1 << 512) - 1]
    arr = np.array([[(w >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(8)] for w in wide], dtype=np.uint64)
    assert orc.mont_to_ints(orc.from_u512(arr)) == [w % R for w in wide]


def test_g1_kats(orc):
    g = orc.g1_generator()
    assert orc.affine_to_ints(g) == [(1, 2)]
    G = orc.g1_from_affine(g)[0]
    two_g = orc.affine_to_ints(orc.g1_batch_normalize(orc.g1_double(G)))[0]
    assert two_g == (0x030644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd3,
                     0x15ed738c0e0a7c92e7845f96b2ae9c0a68a6a449e3538fc7ff3ebf7a5a18a2c4)
    assert two_g == P.g1_add(P.G1_GEN, P.G1_GEN)
    # r*G = identity  (scalar r-1 then +G)
    m = orc.g1_mul(G, orc.ints_to_mont([R - 1])[0])
    assert orc.affine_to_ints(orc.g1_batch_normalize(m))[0] == (1, Q - 2)
    assert orc.affine_to_ints(orc.g1_batch_normalize(orc.g1_add_affine(m, g)))[0] is None
    assert orc.affine_to_ints(orc.g1_batch_normalize(orc.g1_add(m, G)))[0] is None
    # random scalar muls + adds vs big-int affine arithmetic
    rnd = random.Random(3)
    ks = [rnd.randrange(R) for _ in range(6)] + [1, 2, 3]
    pts = orc.g1_fixed_base_mul(orc.ints_to_mont(ks))
    assert orc.g1_on_curve(pts)
    exp = [P.g1_mul(P.G1_GEN, k) for k in ks]
    assert orc.affine_to_ints(pts) == exp
    for i in range(len(ks) - 1):
        s = orc.g1_add(orc.g1_from_affine(pts[i])[0], orc.g1_from_affine(pts[i + 1])[0])
        assert orc.affine_to_ints(orc.g1_batch_normalize(s))[0] == P.g1_add(exp[i], exp[i + 1])
        s2 = orc.g1_add_affine(orc.g1_from_affine(pts[i])[0], pts[i + 1])
        assert orc.affine_to_ints(orc.g1_batch_normalize(s2))[0] == P.g1_add(exp[i], exp[i + 1])
    # add of equal points falls through to doubling; P + (-P) = identity
    d = orc.g1_add(orc.g1_from_affine(pts[0])[0], orc.g1_from_affine(pts[0])[0])
    assert orc.affine_to_ints(orc.g1_batch_normalize(d))[0] == P.g1_add(exp[0], exp[0])
    # compressed encoding: x LE, sign(y) in bit 7 of byte 31, identity all-zero
    comp = orc.g1_compress(pts)
    assert [bytes(c) for c in comp] == [P.g1_compress(e) for e in exp]
    assert bytes(orc.g1_compress(np.zeros((1, 8), dtype=np.uint64))[0]) == bytes(32)


def test_msm_kat(orc):
    # sum_{i=1..8} i * (i*G) = 204*G   
    ks = list(range(1, 9))
    bases = orc.g1_fixed_base_mul(orc.ints_to_mont(ks))
    out = orc.best_multiexp(orc.ints_to_mont(ks), bases)
    got = orc.affine_to_ints(orc.g1_batch_normalize(out))[0]
    assert got == (0x25b77066961904ca2559ca2cabb0ed8ba45f413407bfedf87880427bc08bccab,
                   0x03a0d9bfb355adfbd25052484a7ad058b5cb3c8d7ebeac1c41075b033d234a26)
    assert got == P.g1_mul(P.G1_GEN, 204)


def test_msm_vs_bigint_and_threads(orc):
    rnd = random.Random(4)
    n = 300
    ks = [rnd.randrange(R) for _ in range(n)]
    ss = [rnd.randrange(R) for _ in range(n - 4)] + [0, 1, R - 1, 0]
    bases = orc.g1_fixed_base_mul(orc.ints_to_mont(ks))
    bases[7] = 0                                                    # identity base
    want = P.g1_mul(P.G1_GEN, sum(k * s for i, (k, s) in enumerate(zip(ks, ss)) if i != 7) % R)
    for t in (1, 3, 8):
        orc.set_threads(t)
        got = orc.affine_to_ints(orc.g1_batch_normalize(orc.best_multiexp(orc.ints_to_mont(ss), bases)))[0]
        assert got == want
    orc.set_threads(0)
    for m in (1, 2, 3, 5, 31, 33):                                   # window-size edge cases (c = 1, 3, ceil(ln n))
        got = orc.affine_to_ints(orc.g1_batch_normalize(orc.best_multiexp(orc.ints_to_mont(ss[:m]), bases[:m])))[0]
        assert got == P.g1_mul(P.G1_GEN, sum(k * s for i, (k, s) in enumerate(zip(ks[:m], ss[:m])) if i != 7) % R)


def test_ntt_kat(orc):
    w8 = P.omega_for_k(3)
    assert w8 == 0x2b337de1c8c14f22ec9b9e2f96afef3652627366f8170a0a948dad4ac1bd5e80
    out = orc.mont_to_ints(orc.best_fft(orc.ints_to_mont(list(range(1, 9))), orc.ints_to_mont([w8])[0], 3))
    assert out[0] == 0x24 and out[4] == R - 4
    assert out[1] == 0x002701a4fd3f1d3e7a309cdc72c7c8fcb5c94af009cb48e6e51461367a2f1796
    assert out == P.ntt_naive(list(range(1, 9)), w8)
    assert P.omega_for_k(20) == 0x2a14464f1ff42de3856402b62520e670745e39fada049d5b2f0e1e3182673378


def test_ntt_vs_bigint(orc):
    rnd = random.Random(5)
    for k in (1, 2, 3, 4, 7, 10):
        a = [rnd.randrange(R) for _ in range(1 << k)]
        w = P.omega_for_k(k)
        for t in (1, 8):
            orc.set_threads(t)
            assert orc.mont_to_ints(orc.best_fft(orc.ints_to_mont(a), orc.ints_to_mont([w])[0], k)) == P.ntt(a, w)
    orc.set_threads(0)
    a = [rnd.randrange(R) for _ in range(16)]
    assert P.ntt(a, P.omega_for_k(4)) == P.ntt_naive(a, P.omega_for_k(4))


def test_domain_vs_bigint(orc):
    rnd = random.Random(6)
    for j, k in ((6, 4), (3, 5), (17, 3), (4, 6), (2, 4)):
        d, pd = orc.Domain(j, k), P.Domain(j, k)
        assert d.extended_k == pd.extended_k
        assert orc.mont_to_ints(d.omega)[0] == pd.omega == P.omega_for_k(k)
        assert orc.mont_to_ints(d.extended_omega)[0] == pd.extended_omega
        assert orc.mont_to_ints(d.t_evaluations) == pd.t_evaluations
        a = [rnd.randrange(R) for _ in range(1 << k)]
        A = orc.ints_to_mont(a)
        coeff = d.lagrange_to_coeff(A)
        assert orc.mont_to_ints(coeff) == pd.lagrange_to_coeff(a)
        ext = d.coeff_to_extended(coeff)
        assert orc.mont_to_ints(ext) == pd.coeff_to_extended(pd.lagrange_to_coeff(a))
        # coset evaluation really is p(zeta * w_ext^i)
        c = pd.lagrange_to_coeff(a)
        x = P.FR_ZETA * pow(pd.extended_omega, 5, R) % R
        assert orc.mont_to_ints(ext)[5] == sum(ci * pow(x, i, R) for i, ci in enumerate(c)) % R
        back = d.extended_to_coeff(ext)
        exp_back = c + [0] * ((1 << k) * (j - 1) - (1 << k))
        assert orc.mont_to_ints(back) == exp_back[: (1 << k) * (j - 1)]
        assert orc.mont_to_ints(d.rotate_omega(A[0], -2))[0] == a[0] * pow(pd.omega_inv, 2, R) % R
        dv = d.divide_by_vanishing_poly(ext)
        m = len(pd.t_evaluations)
        assert orc.mont_to_ints(dv) == [v * pd.t_evaluations[i % m] % R for i, v in enumerate(orc.mont_to_ints(ext))]


def test_eval_and_kate(orc):
    rnd = random.Random(7)
    a = [rnd.randrange(R) for _ in range(33)]
    x = rnd.randrange(R)
    X = orc.ints_to_mont([x])[0]
    assert orc.mont_to_ints(orc.eval_polynomial(orc.ints_to_mont(a), X))[0] == sum(c * pow(x, i, R) for i, c in enumerate(a)) % R
    # (a(X) - a(x)) / (X - x)
    a0 = list(a); a0[0] = (a0[0] - sum(c * pow(x, i, R) for i, c in enumerate(a))) % R
    q = orc.mont_to_ints(orc.kate_division(orc.ints_to_mont(a0), X))
    # multiply back
    prod = [0] * 33
    for i, c in enumerate(q):
        prod[i + 1] = (prod[i + 1] + c) % R
        prod[i] = (prod[i] - c * x) % R
    assert prod == a0


def test_transcript_kats():
    t = P.Blake2bTranscript()
    assert t.squeeze_challenge() == 0x0e89c2c9ef365f095ec7aa36500bb0ba58bf7d5e17194055afb5a1c746f1786a
    t = P.Blake2bTranscript(); t.common_scalar(1)
    assert t.squeeze_challenge() == 0x1ba5cdb93688afe0b4eaa4bf9094a4fce372769e41db9e398206953797569832
    t = P.Blake2bTranscript(); t.common_point(P.G1_GEN)
    assert t.squeeze_challenge() == 0x0c0ba67bd0011941b884c2942b53e055abdea83e47dfb046b33bb810b6760239


def test_params_setup(orc):
    rnd = random.Random(8)
    s = rnd.randrange(R)
    k = 4
    g, gl = orc.params_setup(k, orc.ints_to_mont([s])[0])
    assert orc.affine_to_ints(g) == [P.g1_mul(P.G1_GEN, pow(s, i, R)) for i in range(16)]
    # g_lagrange[i] = [l_i(s)] G  ==> commit_lagrange(evals) == commit(coeffs)
    a = [rnd.randrange(R) for _ in range(16)]
    d = orc.Domain(3, k)
    coeff = d.lagrange_to_coeff(orc.ints_to_mont(a))
    c1 = orc.g1_batch_normalize(orc.best_multiexp(orc.ints_to_mont(a), gl))
    c2 = orc.g1_batch_normalize(orc.best_multiexp(coeff, g))
    assert orc.affine_to_ints(c1) == orc.affine_to_ints(c2)
    assert orc.affine_to_ints(c1)[0] == P.g1_mul(P.G1_GEN, sum(c * pow(s, i, R) for i, c in enumerate(orc.mont_to_ints(coeff))) % R)


def test_xorshift_fr_random():
    rng = P.XorShiftRng()
    a = rng.fr_random()
    assert 0 <= a < R and a != rng.fr_random()
    # deterministic restart
    assert P.XorShiftRng().fr_random() == a
