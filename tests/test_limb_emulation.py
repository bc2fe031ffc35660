"""Checks the product's 8x32-bit limb algorithms and kernel index logic, compiled for the
host with the PTX carry flag / block phases emulated, against the CPU oracle.  CPU only."""
import random

import numpy as np
import pytest

from oracle import pyref as P
from tests.emu import binding as emu


@pytest.mark.parametrize("which", [0, 1])
def test_field_limbs_match_oracle(orc, which):
    p = P.R_MOD if which == 0 else P.Q_MOD
    rnd = random.Random(10 + which)
    edge = [0, 1, 2, p - 1, p - 2, p >> 1, (1 << 253) % p, (1 << 32) - 1, (1 << 64) - 1, 1 << 224]
    a = edge + [rnd.randrange(p) for _ in range(300)]
    b = list(reversed(edge)) + [rnd.randrange(p) for _ in range(300)]
    A, B = orc.ints_to_mont(a, which), orc.ints_to_mont(b, which)
    one, r2 = emu.field_consts(which)
    assert orc.mont_to_ints(one, which) == [1]
    assert orc.raw_to_ints(r2) == orc.raw_to_ints(orc.field_params(which)["r2"])
    for op in ("add", "sub", "mul"):
        assert np.array_equal(emu.field_op(which, op, A, B), orc.binop(op, A, B, which)), op
    assert np.array_equal(emu.field_op(which, "sqr", A), orc.binop("mul", A, A, which))
    assert np.array_equal(emu.field_op(which, "from_mont", A), orc.to_raw(A, which))
    assert np.array_equal(emu.field_op(which, "to_mont", orc.ints_to_raw(a)), A)
    nz = orc.ints_to_mont([x for x in a[:40] if x], which)
    assert np.array_equal(emu.field_op(which, "inv", nz), orc.inv(nz, which))
    # worst-case carries: all limbs 0xffffffff is not a field element, but p-1 squared etc. are
    m = orc.ints_to_mont([p - 1] * 4, which)
    assert orc.mont_to_ints(emu.field_op(which, "mul", m, m), which) == [1] * 4


@pytest.mark.parametrize("which", [0, 1])
def test_constant_operand_multiplication(orc, which):
    """Field::mul_shoup (the NTT's twiddle multiplier): for a constant w with wq = floor(w 2^256 / p) it returns x w mod p
    for every x < p — with x in Montgomery form and w plain that is the Montgomery form of the product — fully reduced.
    Edge cases cover the quotient estimate being one short (result in [2p, 3p) before the two conditional subtractions)."""
    p = P.R_MOD if which == 0 else P.Q_MOD
    rnd = random.Random(20 + which)
    edge = [0, 1, 2, p - 1, p - 2, p >> 1, (p >> 1) + 1, (1 << 253) % p, (1 << 32) - 1, (1 << 224) - 1, 1 << 192]
    xs = [a for a in edge for _ in edge] + [rnd.randrange(p) for _ in range(4000)]
    ws = [b for _ in edge for b in edge] + [rnd.randrange(p) for _ in range(4000)]
    wq = [(w << 256) // p for w in ws]
    got = orc.raw_to_ints(emu.mul_shoup(which, orc.ints_to_raw(xs), orc.ints_to_raw(ws), orc.ints_to_raw(wq)))
    assert got == [x * w % p for x, w in zip(xs, ws)]
    # the Montgomery-domain use: x_mont * w_plain = (x w)_mont
    X = orc.ints_to_mont(xs[:200], which)
    want = orc.ints_to_mont([x * w % p for x, w in zip(xs[:200], ws[:200])], which)
    assert np.array_equal(emu.mul_shoup(which, X, orc.ints_to_raw(ws[:200]), orc.ints_to_raw(wq[:200])), want)


@pytest.mark.parametrize("which", [0, 1])
def test_lazy_range_arithmetic(orc, which):
    """The [0, 2p) arithmetic of a warp-level transform pass: mul_shoup_lazy takes ANY x < 2^256 and returns x w mod p as a
    value below 2p; add_lazy keeps [0, 2p); sub_raw returns a - b + 2p (below 4p, fed to the multiplier unreduced);
    reduce_2p brings [0, 4p) back below 2p."""
    p = P.R_MOD if which == 0 else P.Q_MOD
    rnd = random.Random(60 + which)
    edge = [0, 1, p - 1, p, p + 1, 2 * p - 1, 2 * p, 4 * p - 1, (1 << 256) - 1, (1 << 255), 3 * p]
    ws_e = [0, 1, 2, p - 1, p >> 1, (1 << 253) % p]
    xs = [a for a in edge for _ in ws_e] + [rnd.randrange(1 << 256) for _ in range(3000)] + [rnd.randrange(4 * p) for _ in range(1000)]
    ws = [b for _ in edge for b in ws_e] + [rnd.randrange(p) for _ in range(4000)]
    wq = [(w << 256) // p for w in ws]
    got = orc.raw_to_ints(emu.lazy_op(which, "mul", orc.ints_to_raw(xs), orc.ints_to_raw(ws), orc.ints_to_raw(wq)))
    assert all(g < 2 * p for g in got)
    assert [g % p for g in got] == [x * w % p for x, w in zip(xs, ws)]
    a = [rnd.randrange(2 * p) for _ in range(2000)] + [0, 2 * p - 1, 0, 2 * p - 1, p, p]
    b = [rnd.randrange(2 * p) for _ in range(2000)] + [0, 2 * p - 1, 2 * p - 1, 0, p, 0]
    s = orc.raw_to_ints(emu.lazy_op(which, "add", orc.ints_to_raw(a), orc.ints_to_raw(b)))
    assert all(v < 2 * p for v in s) and [v % p for v in s] == [(x + y) % p for x, y in zip(a, b)]
    d = orc.raw_to_ints(emu.lazy_op(which, "sub", orc.ints_to_raw(a), orc.ints_to_raw(b)))
    assert d == [x - y + 2 * p for x, y in zip(a, b)]
    r = orc.raw_to_ints(emu.lazy_op(which, "reduce", orc.ints_to_raw(d)))
    assert all(v < 2 * p for v in r) and [v % p for v in r] == [v % p for v in d]


@pytest.mark.parametrize("which", [0, 1])
def test_division_step_inversion(orc, which):
    """Field::inv_gcd (the inversion inside batch_invert): x^-1 in Montgomery form, identical to the Fermat chain Field::inv
    and to pow(x, p - 2); 0 -> 0.  Edge values sit at limb boundaries of the 9 x 30-bit state."""
    p = P.R_MOD if which == 0 else P.Q_MOD
    rnd = random.Random(40 + which)
    xs = [0, 1, 2, 3, p - 1, p - 2, p >> 1, (p >> 1) + 1, (1 << 253) % p, (1 << 255) % p, 1 << 30, (1 << 30) - 1, 1 << 60,
          1 << 240, (1 << 240) - 1, (1 << 210) + 1]
    xs += [rnd.randrange(p) for _ in range(3000)] + [rnd.randrange(1 << rnd.randrange(1, 254)) for _ in range(500)]
    X = orc.ints_to_mont(xs, which)
    got = emu.field_op(which, "inv_gcd", X)
    assert np.array_equal(got, orc.ints_to_mont([pow(x, p - 2, p) for x in xs], which))
    assert np.array_equal(got[:600], emu.field_op(which, "inv", X[:600]))


@pytest.mark.parametrize("log_n,max_log_m,max_log_tw,cap", [
    (0, 10, 2, 12), (1, 10, 2, 12), (3, 10, 2, 12), (6, 10, 2, 12),      # single pass
    (6, 3, 1, 12), (7, 4, 2, 12), (8, 4, 3, 5),                           # two passes, tile cap binding
    (6, 2, 1, 12), (9, 3, 2, 12), (8, 3, 0, 12), (10, 4, 2, 5),          # three passes
    (10, 5, 3, 12), (11, 4, 3, 12), (12, 4, 3, 12), (8, 8, 3, 12),        # tiles of 8 columns: j-major butterfly order
])
def test_ntt_pass_structure(orc, log_n, max_log_m, max_log_tw, cap):
    rnd = random.Random(100 + log_n)
    n = 1 << log_n
    a = orc.ints_to_mont([rnd.randrange(P.R_MOD) for _ in range(n)])
    w = orc.ints_to_mont([P.omega_for_k(log_n)])[0]
    got = emu.ntt(a, log_n, w, max_log_m=max_log_m, max_log_tw=max_log_tw, tile_cap_log=cap, nthreads=7)
    assert np.array_equal(got, orc.best_fft(a, w, log_n))


def test_ntt_fused_domain_hooks(orc):
    """coeff_to_extended / extended_to_coeff / lagrange_to_coeff as pre/post hooks."""
    rnd = random.Random(77)
    k, j = 5, 6
    d = orc.Domain(j, k)
    n, ext_k = 1 << k, d.extended_k
    a = orc.ints_to_mont([rnd.randrange(P.R_MOD) for _ in range(n)])
    one = orc.ints_to_mont([1])[0]
    # lagrange_to_coeff: iNTT with post = ifft_divisor
    post = np.stack([d.ifft_divisor] * 3)
    got = emu.ntt(a, k, d.omega_inv, max_log_m=3, max_log_tw=1, post=post)
    coeff = d.lagrange_to_coeff(a)
    assert np.array_equal(got, coeff)
    # coeff_to_extended: pre = [1, zeta, zeta^2], zero padded
    pre = np.stack([one, d.g_coset, d.g_coset_inv])
    ext = emu.ntt(coeff, ext_k, d.extended_omega, n_in=n, max_log_m=3, max_log_tw=2, pre=pre)
    assert np.array_equal(ext, d.coeff_to_extended(coeff))
    # extended_to_coeff: post = ext^-1 * [1, zeta^-1, zeta^-2] = ext^-1 * [1, zeta^2, zeta]
    div = d.extended_ifft_divisor
    post = np.stack([div, orc.binop("mul", div, d.g_coset_inv)[0], orc.binop("mul", div, d.g_coset)[0]])
    back = emu.ntt(ext, ext_k, d.extended_omega_inv, max_log_m=4, max_log_tw=2, post=post)
    assert np.array_equal(back[: n * (j - 1)], d.extended_to_coeff(ext))


@pytest.mark.parametrize("which", [0, 1])
def test_host_field_matches_oracle(orc, which):
    p = P.R_MOD if which == 0 else P.Q_MOD
    rnd = random.Random(20 + which)
    a = [0, 1, p - 1, p - 2] + [rnd.randrange(p) for _ in range(200)]
    b = [p - 1, 0, p - 1, 1] + [rnd.randrange(p) for _ in range(200)]
    A, B = orc.ints_to_mont(a, which), orc.ints_to_mont(b, which)
    for op in ("add", "sub", "mul"):
        assert np.array_equal(emu.host_field_op(which, op, A, B), orc.binop(op, A, B, which)), op
    nz = orc.ints_to_mont([x for x in a[:30] if x], which)
    assert np.array_equal(emu.host_field_op(which, "inv", nz), orc.inv(nz, which))


def test_host_fr_constants(orc):
    wide = np.arange(1, 9, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    r, z, d, w = emu.host_fr_consts(wide)
    R, D, Z = orc.fr_constants()
    assert np.array_equal(r, R) and np.array_equal(z, Z) and np.array_equal(d, D)
    assert np.array_equal(w, orc.from_u512(wide.reshape(1, 8))[0])


@pytest.mark.parametrize("n,c", [(1, 0), (2, 0), (37, 4), (200, 0), (200, 7), (64, 16), (300, 5)])
def test_msm_pipeline_structure(orc, n, c):
    rnd = random.Random(200 + n + c)
    ks = [rnd.randrange(P.R_MOD) for _ in range(n)]
    ss = [rnd.randrange(P.R_MOD) for _ in range(n)]
    if n >= 37:
        ss[0], ss[1], ss[2], ss[3] = 0, 1, P.R_MOD - 1, (1 << 253)       # edge scalars
        ks[5] = ks[4]                                                     # repeated base (P + P inside a bucket)
        ss[5] = ss[4]
        ks[7] = P.R_MOD - ks[6]                                           # P + (-P)
        ss[7] = ss[6]
    bases = orc.g1_fixed_base_mul(orc.ints_to_mont(ks))
    if n >= 37:
        bases[9] = 0                                                      # identity base
    S = orc.ints_to_mont(ss)
    got = emu.msm(S, bases, c)
    want = orc.g1_batch_normalize(orc.best_multiexp(S, bases))[0]
    assert np.array_equal(got, want)


def test_msm_all_same_small_scalars(orc):
    """Skewed digit distribution: every point lands in one bucket."""
    n = 100
    ks = list(range(1, n + 1))
    bases = orc.g1_fixed_base_mul(orc.ints_to_mont(ks))
    for s in (1, 3, 255):
        S = orc.ints_to_mont([s] * n)
        assert np.array_equal(emu.msm(S, bases, 6), orc.g1_batch_normalize(orc.best_multiexp(S, bases))[0])
    Z = orc.ints_to_mont([0] * n)
    assert np.array_equal(emu.msm(Z, bases, 0), np.zeros(8, dtype=np.uint64))


def test_poly_primitives_structure(orc):
    rnd = random.Random(300)
    R = P.R_MOD
    for n, m, T, lanes in ((1, 4, 2, 1), (7, 4, 2, 3), (64, 4, 4, 5), (203, 8, 4, 16), (1000, 16, 8, 40)):
        vals = [rnd.randrange(R) for _ in range(n)]
        if n > 5:
            vals[3] = 0
        A = orc.ints_to_mont(vals)
        assert np.array_equal(emu.batch_invert(0, A, lanes), orc.batch_invert(A))
        b = rnd.randrange(R)
        B = orc.ints_to_mont([b])[0]
        y, head = emu.recurrence(A, B, m, T)
        assert np.array_equal(head, orc.eval_polynomial(A, B))
        _, head2 = emu.recurrence(A, B, m, T, want_y=False)
        assert np.array_equal(head2, head)
        if n > 1:
            # kate_division(a, b) = recurrence on a[1:]
            q, _ = emu.recurrence(A[1:], B, m, T)
            assert np.array_equal(q, orc.kate_division(A, B))
        z0 = rnd.randrange(R)
        want, run = [], z0
        for v in vals:
            want.append(run); run = run * v % R
        assert orc.mont_to_ints(emu.prefix_product(A, orc.ints_to_mont([z0])[0], m, T)) == want
    fq = orc.ints_to_mont([5, 0, 7, P.Q_MOD - 1], orc.FQ)
    assert orc.mont_to_ints(emu.batch_invert(1, fq, 2), orc.FQ) == [pow(5, -1, P.Q_MOD), 0, pow(7, -1, P.Q_MOD), P.Q_MOD - 1]


def test_host_transcript_matches_pyref(orc):
    """The product's BLAKE2b transcript vs hashlib (oracle/pyref.py) incl. the SURVEY KATs."""
    ch, _ = emu.transcript([0], np.zeros(4, dtype=np.uint64), 1)
    assert orc.mont_to_ints(ch)[0] == 0x0e89c2c9ef365f095ec7aa36500bb0ba58bf7d5e17194055afb5a1c746f1786a
    ch, _ = emu.transcript([1, 0], orc.ints_to_mont([1]).reshape(-1), 1)
    assert orc.mont_to_ints(ch)[0] == 0x1ba5cdb93688afe0b4eaa4bf9094a4fce372769e41db9e398206953797569832
    # long mixed script (crosses several 128-byte blocks), compared step by step with pyref
    rnd = random.Random(9)
    t = P.Blake2bTranscript()
    ops, data, want = [], [], []
    pts = [P.g1_mul(P.G1_GEN, rnd.randrange(P.R_MOD)) for _ in range(5)] + [None]
    for i in range(40):
        r = rnd.randrange(4)
        if r == 0:
            ops.append(0); want.append(t.squeeze_challenge())
        elif r == 1:
            s = rnd.randrange(P.R_MOD); ops.append(1); data += list(orc.ints_to_mont([s])[0]); t.common_scalar(s)
        elif r == 2:
            p = pts[rnd.randrange(len(pts))]
            x, y = (0, 0) if p is None else p
            ops.append(2); data += list(orc.ints_to_mont([x], orc.FQ)[0]) + list(orc.ints_to_mont([y], orc.FQ)[0]); t.write_point(p)
        else:
            s = rnd.randrange(P.R_MOD); ops.append(3); data += list(orc.ints_to_mont([s])[0]); t.write_scalar(s)
    ops.append(0); want.append(t.squeeze_challenge())
    ch, proof = emu.transcript(ops, np.array(data, dtype=np.uint64), len(want))
    assert orc.mont_to_ints(ch) == want
    assert proof == bytes(t.proof)


def test_device_from_u512_unreduced_inputs(orc):
    """Fr::random on the device path: 512-bit inputs whose halves exceed r (incl. all-ones)."""
    rng = np.random.Generator(np.random.PCG64(5))
    wide = rng.integers(0, 1 << 64, size=(20000, 8), dtype=np.uint64)
    wide[0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    wide[1, :4] = np.uint64(0xFFFFFFFFFFFFFFFF); wide[1, 4:] = 0
    wide[2] = 0
    wide[3:200, 3] |= np.uint64(0xFFFFFFFF00000000)       # top limb of the low half near 2^64
    wide[3:200, 7] |= np.uint64(0xFFFFFFFFF0000000)
    assert np.array_equal(emu.from_u512(wide), orc.from_u512(wide))


@pytest.mark.parametrize("seg_len", [2, 3, 7])
def test_msm_task_balanced_accumulation(orc, seg_len):
    """General (three-level, task-balanced) bucket accumulation forced on, incl. skewed columns."""
    rnd = random.Random(400 + seg_len)
    n = 150
    ks = [rnd.randrange(P.R_MOD) for _ in range(n)]
    bases = orc.g1_fixed_base_mul(orc.ints_to_mont(ks))
    for name, ss in (("uniform", [rnd.randrange(P.R_MOD) for _ in range(n)]),
                     ("bits", [rnd.randrange(2) for _ in range(n)]),
                     ("bytes", [rnd.randrange(256) for _ in range(n)]),
                     ("one-bucket", [3] * n),
                     ("sparse", [0] * (n - 3) + [5, 0, 7])):
        S = orc.ints_to_mont(ss)
        got = emu.msm(S, bases, force_c=5, fast_max=0, seg_len=seg_len)
        assert np.array_equal(got, orc.g1_batch_normalize(orc.best_multiexp(S, bases))[0]), name


@pytest.mark.parametrize("c", [4, 5, 7])
def test_msm_fixed_base_table_path(orc, c):
    """Fixed-base mode: one bucket set over a table of 2^(cj) multiples, bit-decomposition reduce."""
    rnd = random.Random(500 + c)
    n = 40
    ks = [rnd.randrange(P.R_MOD) for _ in range(n)]
    bases = orc.g1_fixed_base_mul(orc.ints_to_mont(ks))
    bases[3] = 0                                                       # identity base
    for name, ss, m in (("uniform", [rnd.randrange(P.R_MOD) for _ in range(n)], n),
                        ("edge", [0, 1, P.R_MOD - 1, 1 << 253, (1 << 254) % P.R_MOD] + [rnd.randrange(P.R_MOD) for _ in range(n - 5)], n),
                        ("bytes", [rnd.randrange(256) for _ in range(n)], n),
                        ("short", [rnd.randrange(P.R_MOD) for _ in range(17)], 17)):    # commit of a shorter polynomial
        S = orc.ints_to_mont(ss)
        got = emu.msm_pre(S, bases, c)
        assert np.array_equal(got, orc.g1_batch_normalize(orc.best_multiexp(S, bases[:m]))[0]), name


@pytest.mark.parametrize("C,CL", [(5, 4), (16, 4), (16, 6), (3, 3)])
def test_quotient_coset_interpolation_and_lookup_extrapolation(orc, C, CL):
    """prover_kernels.cuh: h(X) = sum_r X^r H_r(X^n) is recovered from its per-coset iNTTs by the C x C inverse Vandermonde
    (coset_interpolate_row), and a part of degree < CL * n that was evaluated on the first CL cosets only is carried to the
    other cosets by Lagrange extrapolation in H-space (lookup_extrapolate_row) — checked against plain integers."""
    rnd = random.Random(1000 + 16 * C + CL)
    R, n = P.R_MOD, 8
    cs = [rnd.randrange(1, R) for _ in range(C)]                    # coset generators (any distinct non-zero values)
    ys = [pow(c, n, R) for c in cs]
    tinv = [rnd.randrange(1, R) for _ in range(C)]                  # stands for 1 / (c_j^n - 1)
    M = lambda v: orc.ints_to_mont([x % R for x in v])
    inv_pow = M([pow(c, -r, R) for c in cs for r in range(n)])
    ev = lambda coeffs, x: sum(cf * pow(x, t, R) for t, cf in enumerate(coeffs)) % R
    # the part known on all C cosets (gates, permutation) and the low-degree part (lookups)
    H_full = [[rnd.randrange(R) for _ in range(C)] for _ in range(n)]
    H_low = [[rnd.randrange(R) for _ in range(CL)] for _ in range(n)]
    g_full = M([pow(cs[j], r, R) * ev(H_full[r], ys[j]) for j in range(C) for r in range(n)])
    g_low = [pow(cs[j], r, R) * ev(H_low[r], ys[j]) if j < CL else rnd.randrange(R) for j in range(C) for r in range(n)]
    lam = [[1] * CL for _ in range(max(C - CL, 1))]
    for jp in range(CL, C):
        for j in range(CL):
            num = den = 1
            for m in range(CL):
                if m != j:
                    num = num * (ys[jp] - ys[m]) % R; den = den * (ys[j] - ys[m]) % R
            lam[jp - CL][j] = num * pow(den, -1, R) % R
    got = emu.lookup_extrapolate(M(g_low), inv_pow, M([x for row in lam for x in row]), M(tinv), CL, C, n)
    want = [ev(H_low[r], ys[j]) * tinv[j] % R for j in range(C) for r in range(n)]
    assert orc.mont_to_ints(got) == want
    # inverse Vandermonde of the y_j by Gauss-Jordan over the integers mod R
    A = [[pow(ys[j], t, R) for t in range(C)] + [int(i == j) for i in range(C)] for j in range(C)]
    for c in range(C):
        piv = next(r for r in range(c, C) if A[r][c])
        A[c], A[piv] = A[piv], A[c]
        inv = pow(A[c][c], -1, R)
        A[c] = [v * inv % R for v in A[c]]
        for r in range(C):
            if r != c and A[r][c]:
                f = A[r][c]
                A[r] = [(a - f * b) % R for a, b in zip(A[r], A[c])]
    vinv = M([A[t][C + j] for t in range(C) for j in range(C)])
    pieces = emu.coset_interpolate(g_full, inv_pow, vinv, C, n)
    assert orc.mont_to_ints(pieces) == [H_full[r][t] for t in range(C) for r in range(n)]
    # both parts together: extra = the extrapolated low part (with t_j = 1 here so that it is H_low itself)
    extra = emu.lookup_extrapolate(M(g_low), inv_pow, M([x for row in lam for x in row]), M([1] * C), CL, C, n)
    both = emu.coset_interpolate(g_full, inv_pow, vinv, C, n, extra)
    assert orc.mont_to_ints(both) == [(H_full[r][t] + (H_low[r][t] if t < CL else 0)) % R for t in range(C) for r in range(n)]
