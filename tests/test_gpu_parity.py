"""GPU parity tests: the CUDA path through the C ABI vs the CPU oracle on the same seeded
inputs (bit-exact: everything here is integer arithmetic), plus size-independent properties
at the sizes the oracle is too slow for."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(orc, k, seed=1):
    s = orc.random_fr(1, 1000 + seed)[0]
    return orc.params_setup(k, s)


def _affine(out12):
    return out12[:8]


@pytest.mark.parametrize("k", [0, 1, 2, 3, 5, 8, 10, 11, 13, 16])
def test_best_fft_matches_oracle(backend, orc, k):
    a = orc.random_fr(1 << k, 10 + k)
    from oracle import pyref
    w = orc.ints_to_mont([pyref.omega_for_k(k)])[0]
    got = backend.best_fft(a, w, k)
    assert np.array_equal(got, orc.best_fft(a, w, k))
    # generic omega (not a root of unity of that order is still a valid best_fft input)
    w2 = orc.random_fr(1, 99)[0]
    if k <= 8:
        assert np.array_equal(backend.best_fft(a, w2, k), orc.best_fft(a, w2, k))


@pytest.mark.parametrize("k", [12, 13, 16, 18])
@pytest.mark.parametrize("kind", ["head", "head+tail", "one", "zero", "tile"])
def test_best_fft_sparse_inputs(zk, backend, orc, k, kind):
    """Padded witness columns: the warp-level kernel skips tiles whose inputs are all zero (ntt_warp.cuh);
    transforms of sparse inputs — and iNTT / coset extension through the domain — stay bit-exact."""
    from oracle import pyref
    n = 1 << k
    a = np.zeros((n, 4), dtype=np.uint64)
    r = orc.random_fr(300, 40 + k)
    if kind in ("head", "head+tail"):
        a[:250] = r[:250]
    if kind == "head+tail":
        a[n - 6:] = r[250:256]
    if kind == "one":
        a[n // 3] = r[0]
    if kind == "tile":                     # exactly one first-pass tile's worth of rows, strided
        a[5::n // 128] = r[:128]
    w = orc.ints_to_mont([pyref.omega_for_k(k)])[0]
    assert np.array_equal(backend.best_fft(a, w, k), orc.best_fft(a, w, k))
    if k <= 13:
        d, od = zk.EvaluationDomain(backend, 6, k), orc.Domain(6, k)
        coeff = d.lagrange_to_coeff(a)
        assert np.array_equal(coeff, od.lagrange_to_coeff(a))
        sp = np.zeros((n, 4), dtype=np.uint64)
        sp[:3] = r[:3]                         # a sparse COEFFICIENT vector (e.g. a constant / low-degree column)
        assert np.array_equal(d.coeff_to_extended(sp), od.coeff_to_extended(sp))
        d.close()


@pytest.mark.parametrize("k", [18, 20, 21, 22, 23, 25])
def test_best_fft_large_roundtrip_and_spot(backend, orc, k):
    from oracle import pyref
    n = 1 << k
    a = orc.random_fr(n, 30 + k)
    w = pyref.omega_for_k(k)
    W, WI = orc.ints_to_mont([w])[0], orc.ints_to_mont([pow(w, -1, pyref.R_MOD)])[0]
    f = backend.best_fft(a, W, k)
    if k <= 20:
        assert np.array_equal(f, orc.best_fft(a, W, k))
    else:
        # spot-check a few outputs by Horner: out[i] = a(omega^i)
        for i in (0, 1, n // 2 + 3, n - 1):
            x = orc.ints_to_mont([pow(w, i, pyref.R_MOD)])[0]
            assert np.array_equal(f[i], orc.eval_polynomial(a, x))
    back = backend.best_fft(f, WI, k)
    ninv = orc.ints_to_mont([pow(n, -1, pyref.R_MOD)])[0]
    assert np.array_equal(orc.binop("mul", back, np.tile(ninv, (n, 1))), a)


@pytest.mark.parametrize("j,k", [(6, 4), (6, 9), (3, 10), (17, 6), (6, 14), (4, 12), (2, 7)])
def test_domain_matches_oracle(zk, backend, orc, j, k):
    d, od = zk.EvaluationDomain(backend, j, k), orc.Domain(j, k)
    assert d.extended_k == od.extended_k and d.quotient_poly_degree == od.quotient_poly_degree
    for name in d._CONSTS:
        assert np.array_equal(getattr(d, name), getattr(od, name)), name
    a = orc.random_fr(1 << k, 50 + k)
    coeff = d.lagrange_to_coeff(a)
    assert np.array_equal(coeff, od.lagrange_to_coeff(a))
    ext = d.coeff_to_extended(coeff)
    assert np.array_equal(ext, od.coeff_to_extended(coeff))
    assert np.array_equal(d.divide_by_vanishing_poly(ext), od.divide_by_vanishing_poly(ext))
    e2 = orc.random_fr(1 << od.extended_k, 60 + k)
    assert np.array_equal(d.extended_to_coeff(e2), od.extended_to_coeff(e2))
    # round trip: extended_to_coeff(coeff_to_extended(p)) = p || 0
    back = d.extended_to_coeff(ext)
    assert np.array_equal(back[: 1 << k], coeff) and not back[1 << k:].any()
    # rotate_omega, l_i_range, rotate_extended
    x = orc.random_fr(1, 70 + k)[0]
    for rot in (0, 1, -1, 5, -6):
        assert np.array_equal(d.rotate_omega(x, rot), od.rotate_omega(x, rot))
    n, R = 1 << k, 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
    xi, wi = orc.mont_to_ints(x)[0], orc.mont_to_ints(od.omega)[0]
    lo, hi = -min(6, n - 1), min(3, n - 1)
    want = [(pow(xi, n, R) - 1) * pow(n, -1, R) * pow(wi, i % n, R) * pow(xi - pow(wi, i % n, R), -1, R) % R for i in range(lo, hi + 1)]
    assert orc.mont_to_ints(d.l_i_range(x, lo, hi)) == want
    full = orc.mont_to_ints(d.l_i_range(x, 0, n - 1)) if k <= 7 else None
    if full is not None:
        assert sum(full) % R == 1                                      # the Lagrange basis sums to one
    scale = 1 << (od.extended_k - k)
    for rot in (1, -1, 3):
        assert np.array_equal(d.rotate_extended(ext, rot), np.roll(ext, -rot * scale, axis=0))
    d.close()


@pytest.mark.parametrize("n", [0, 1, 2, 3, 31, 32, 33, 255, 1000, 4096, 1 << 14, (1 << 16) + 7])
def test_best_multiexp_matches_oracle(backend, orc, n):
    k = max(1, (max(n, 1) - 1).bit_length())
    g, _ = _setup(orc, k)
    coeffs = orc.random_fr(n, 70 + (n % 97))
    got = backend.best_multiexp(coeffs, g[:n])
    if n == 0:
        # G1::identity() = (0, 1, 0)
        assert not got[:4].any() and not got[8:].any() and np.array_equal(got[4:8], orc.ints_to_mont([1], orc.FQ)[0])
        return
    want = orc.g1_batch_normalize(orc.best_multiexp(coeffs, g[:n]))[0]
    assert np.array_equal(_affine(got), want)


@pytest.mark.parametrize("c", [4, 7, 11, 13, 16])
def test_best_multiexp_window_sizes(backend, orc, c):
    n = 3000
    g, _ = _setup(orc, 12)
    coeffs = orc.random_fr(n, 5)
    backend.set_msm_window(c)
    try:
        got = backend.best_multiexp(coeffs, g[:n])
    finally:
        backend.set_msm_window(0)
    assert np.array_equal(_affine(got), orc.g1_batch_normalize(orc.best_multiexp(coeffs, g[:n]))[0])


def test_best_multiexp_edge_inputs(backend, orc):
    """zero / one / r-1 scalars, identity and repeated bases, P + (-P), witness-like sparse columns."""
    from oracle import pyref
    n = 2048
    g, _ = _setup(orc, 11)
    g = np.array(g)
    ss = orc.mont_to_ints(orc.random_fr(n, 6))
    ss[0], ss[1], ss[2], ss[3] = 0, 1, pyref.R_MOD - 1, 1 << 253
    g[5] = g[4]; ss[5] = ss[4]                      # P + P inside one bucket
    g[7, :4] = g[6, :4]                             # -P: same x, negated y
    g[7, 4:] = orc.binop("sub", np.zeros((1, 4), dtype=np.uint64), g[6, 4:].reshape(1, 4), orc.FQ)[0]
    ss[7] = ss[6]
    g[9] = 0                                        # identity base
    S = orc.ints_to_mont(ss)
    assert np.array_equal(_affine(backend.best_multiexp(S, g)), orc.g1_batch_normalize(orc.best_multiexp(S, g))[0])
    # 99.9 % zeros, the rest < 2^64 ("witness-like")
    sparse = [0] * n
    for i in range(0, n, 512):
        sparse[i] = (i * 0x9E3779B97F4A7C15) & ((1 << 64) - 1)
    S = orc.ints_to_mont(sparse)
    assert np.array_equal(_affine(backend.best_multiexp(S, g)), orc.g1_batch_normalize(orc.best_multiexp(S, g))[0])
    # all scalars equal and tiny: every point lands in one bucket
    S = orc.ints_to_mont([3] * n)
    assert np.array_equal(_affine(backend.best_multiexp(S, g)), orc.g1_batch_normalize(orc.best_multiexp(S, g))[0])
    # all-zero column -> identity
    out = backend.best_multiexp(orc.ints_to_mont([0] * n), g)
    assert not out[:4].any() and not out[8:].any()
    # length mismatch is an error, as upstream's assert_eq!
    with pytest.raises(Exception):
        backend.best_multiexp(S[:10], g[:11])


def test_best_multiexp_linearity_large(backend, orc):
    """2^20 points: MSM(a + b) == MSM(a) + MSM(b) and agreement with the oracle on a 2^18 prefix."""
    k = 20
    n = 1 << k
    g, _ = _setup(orc, k)
    a, b = orc.random_fr(n, 80), orc.random_fr(n, 81)
    ra, rb = backend.best_multiexp(a, g), backend.best_multiexp(b, g)
    rab = backend.best_multiexp(orc.binop("add", a, b), g)
    s = orc.g1_add(orc.g1_from_affine(_affine(ra))[0], orc.g1_from_affine(_affine(rb))[0])
    assert np.array_equal(orc.g1_batch_normalize(s)[0], _affine(rab))
    m = 1 << 18
    assert np.array_equal(_affine(backend.best_multiexp(a[:m], g[:m])), orc.g1_batch_normalize(orc.best_multiexp(a[:m], g[:m]))[0])


def test_params_commit_matches_oracle(zk, backend, orc):
    k = 10
    g, gl = _setup(orc, k)
    params = zk.ParamsKZG.load(backend, k, g, gl)
    poly = orc.random_fr(1 << k, 90)
    assert np.array_equal(_affine(params.commit(poly)), orc.g1_batch_normalize(orc.best_multiexp(poly, g))[0])
    assert np.array_equal(_affine(params.commit_lagrange(poly)), orc.g1_batch_normalize(orc.best_multiexp(poly, gl))[0])
    short = poly[:300]
    assert np.array_equal(_affine(params.commit(short)), orc.g1_batch_normalize(orc.best_multiexp(short, g[:300]))[0])
    # commit_lagrange(evals) == commit(lagrange_to_coeff(evals))
    d = zk.EvaluationDomain(backend, 3, k)
    assert np.array_equal(params.commit_lagrange(poly), params.commit(d.lagrange_to_coeff(poly)))
    g2, gl2 = params.read()
    assert np.array_equal(g2, g) and np.array_equal(gl2, gl)
    d.close(); params.close()


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 1000, 4097, 1 << 16, (1 << 18) + 5])
def test_poly_primitives_match_oracle(backend, orc, n):
    from oracle import pyref
    a = np.array(orc.random_fr(n, 120 + n % 13))
    if n > 10:
        a[3] = 0
    x = orc.random_fr(1, 7)[0]
    assert np.array_equal(backend.eval_polynomial(a, x), orc.eval_polynomial(a, x))
    if n > 1:
        assert np.array_equal(backend.kate_division(a, x), orc.kate_division(a, x))
    assert np.array_equal(backend.batch_invert(a), orc.batch_invert(a))
    z = backend.prefix_product(a, x)
    # z[0] = x, z[i] = z[i-1] * a[i-1]
    assert np.array_equal(z[0], x)
    if n > 1:
        assert np.array_equal(z[1:], orc.binop("mul", z[:-1], a[:-1]))


def test_batch_invert_fq(backend, orc):
    a = orc.ints_to_mont([5, 0, 7, 11, 13], orc.FQ)
    assert np.array_equal(backend.batch_invert(a, field=1)[[0, 2, 3, 4]], orc.inv(a[[0, 2, 3, 4]], orc.FQ))


@pytest.mark.parametrize("k", [1, 4, 9, 13])
def test_params_setup_matches_oracle(zk, backend, orc, k):
    """ParamsKZG::setup on the device vs the oracle's restatement, same secret s."""
    s = orc.random_fr(1, 2000 + k)[0]
    params = zk.ParamsKZG.setup(backend, k, s)
    g, gl = params.read()
    og, ogl = orc.params_setup(k, s)
    assert np.array_equal(g, og)
    assert np.array_equal(gl, ogl)
    params.close()


def test_params_setup_large_consistency(zk, backend, orc):
    """k = 18: commit_lagrange(evals) == commit(coeffs) ties both device-generated bases together,
    and g[i] spot-checks against the oracle."""
    k = 18
    s = orc.random_fr(1, 31337)[0]
    params = zk.ParamsKZG.setup(backend, k, s)
    d = zk.EvaluationDomain(backend, 3, k)
    evals = orc.random_fr(1 << k, 55)
    assert np.array_equal(params.commit_lagrange(evals), params.commit(d.lagrange_to_coeff(evals)))
    g, _ = params.read()
    assert orc.g1_on_curve(g[:: 1 << 10])
    sp = orc.mont_to_ints(s)[0]
    from oracle import pyref
    idx = [0, 1, 2, 12345, (1 << k) - 1]
    want = orc.g1_fixed_base_mul(orc.ints_to_mont([pow(sp, i, pyref.R_MOD) for i in idx]))
    assert np.array_equal(g[idx], want)
    d.close(); params.close()


@pytest.mark.parametrize("kind", ["bits", "bytes", "u40", "one-bucket", "mixed", "z-like", "sorted-runs"])
def test_best_multiexp_skewed_columns(backend, orc, kind):
    """Witness-like columns (bits, bytes, small integers) put thousands of points in a few buckets:
    exercises the task-balanced accumulation path."""
    k = 15
    n = 1 << k
    g, _ = _setup(orc, k)
    rng = np.random.Generator(np.random.PCG64(7))
    if kind == "bits":
        vals = rng.integers(0, 2, size=n)
    elif kind == "bytes":
        vals = rng.integers(0, 256, size=n)
    elif kind == "u40":
        vals = rng.integers(0, 1 << 40, size=n)
    elif kind == "one-bucket":
        vals = np.full(n, 3)
    elif kind == "z-like":                               # grand product of a padded circuit: 1 on the unused rows
        vals = np.ones(n, dtype=np.int64)
    elif kind == "sorted-runs":                          # permuted lookup column: sorted bytes in long runs
        vals = np.sort(rng.integers(0, 256, size=n))
    else:
        vals = np.where(rng.integers(0, 4, size=n) == 0, rng.integers(0, 1 << 62, size=n), rng.integers(0, 2, size=n))
    lut_keys, inv = np.unique(vals, return_inverse=True)
    S = orc.ints_to_mont([int(v) for v in lut_keys])[inv]
    if kind == "mixed":
        S[: n // 8] = orc.random_fr(n // 8, 3)           # a dense stretch on top
    if kind == "z-like":
        S[:700] = orc.random_fr(700, 5)                  # the active rows
        S[n - 6:] = orc.random_fr(6, 6)                  # blinding rows
    got = backend.best_multiexp(S, g)
    want = orc.g1_batch_normalize(orc.best_multiexp(S, g))[0]
    assert np.array_equal(_affine(got), want)
    if kind in ("z-like", "sorted-runs", "one-bucket"):  # same column through the fixed-base (SRS commit) path
        import importlib
        zk = importlib.import_module(type(backend).__module__)
        params = zk.ParamsKZG.load(backend, k, g, None)
        assert np.array_equal(_affine(params.commit(S)), want)
        params.close()


@pytest.mark.parametrize("lagrange", [False, True])
def test_commit_many_matches_single_commits(zk, backend, orc, lagrange):
    """Batched commits (one bucket set per column): seven columns of different character — dense,
    zero, constant-run (grand-product like), sorted bytes, sparse, all-equal, bits — against the oracle."""
    k = 14
    n = 1 << k
    g, gl = _setup(orc, k)
    params = zk.ParamsKZG.load(backend, k, g, gl)
    rng = np.random.Generator(np.random.PCG64(21))
    small = lambda v: orc.ints_to_mont([int(x) for x in np.unique(v)])[np.unique(v, return_inverse=True)[1]]
    cols = [orc.random_fr(n, 31)]
    cols.append(np.zeros((n, 4), dtype=np.uint64))
    z = np.repeat(orc.random_fr(1, 32), n, axis=0); z[:500] = orc.random_fr(500, 33); z[n - 6:] = orc.random_fr(6, 34)
    cols.append(z)
    cols.append(small(np.sort(rng.integers(0, 256, size=n))))
    sp = np.zeros((n, 4), dtype=np.uint64); sp[::997] = orc.random_fr(len(sp[::997]), 35)
    cols.append(sp)
    cols.append(small(np.full(n, 7)))
    cols.append(small(rng.integers(0, 2, size=n)))
    d = [backend.to_device(c) for c in cols]
    got = params.commit_many_dev(d, n, lagrange)
    bases = gl if lagrange else g
    for i, c in enumerate(cols):
        want = orc.g1_batch_normalize(orc.best_multiexp(c, bases))[0]
        if not want.any():                                   # identity: G1 (0, 1, 0)
            assert not got[i][:4].any() and not got[i][8:].any(), f"column {i}"
        else:
            assert np.array_equal(_affine(got[i]), want), f"column {i}"
        assert np.array_equal(got[i], params.commit_dev(d[i], n, lagrange)), f"column {i} vs single commit"
    for b in d:
        b.free()
    params.close()


@pytest.mark.parametrize("log_n,log_r,world", [(10, 4, 2), (12, 6, 4), (16, 8, 8), (20, 10, 8), (16, 5, 2), (20, 7, 8)])
def test_four_step_sharded_ntt_kernels(zk, backend, orc, log_n, log_r, world):
    """Column step / row step kernels of the sharded four-step NTT, with the `world` ranks emulated
    one after another on one GPU and the all-to-all done on the host; result = best_fft."""
    import importlib
    from oracle import pyref
    sharded = importlib.import_module(zk.__name__ + ".sharded")
    N, R, C = 1 << log_n, 1 << log_r, 1 << (log_n - log_r)
    a = orc.random_fr(N, 300 + log_n)
    w = pyref.omega_for_k(log_n)
    omega_n = orc.ints_to_mont([w])[0]
    omega_c = orc.ints_to_mont([pow(w, R, pyref.R_MOD)])[0]
    eng = sharded.GpuNttEngine(zk, backend, log_n)
    cg, rg = C // world, R // world
    Y = np.zeros((R, C, 4), dtype=np.uint64)
    for g in range(world):
        fs = sharded.FourStepNTT(eng, log_n, log_r, g, world)
        Y[:, g * cg:(g + 1) * cg] = eng.col_step(fs.local_columns(a), omega_n, log_r, log_n - log_r, g * cg)
    Z = np.zeros((R, C, 4), dtype=np.uint64)
    for g in range(world):
        Z[g * rg:(g + 1) * rg] = eng.row_step(Y[g * rg:(g + 1) * rg], omega_c, log_n - log_r)
    got = np.ascontiguousarray(np.transpose(Z, (1, 0, 2))).reshape(N, 4)          # X[k_r + R k_c] = Z[k_r][k_c]
    want = orc.best_fft(a, omega_n, log_n) if log_n <= 16 else backend.best_fft(a, omega_n, log_n)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("log_n,log_r,world", [(10, 4, 2), (14, 6, 4), (18, 10, 8), (16, 5, 2), (20, 7, 8), (18, 7, 4)])
def test_four_step_fused_scatter_kernel(zk, backend, orc, log_n, log_r, world):
    """b200zk_fft_colstep_scatter_dev: the column step that writes every transformed row into the
    owner's row buffer (NVLink peer stores between processes; here the `world` ranks are emulated
    one after another on one GPU and the "peer" buffers are local allocations).  Result = best_fft."""
    import ctypes
    from oracle import pyref
    N, R, C = 1 << log_n, 1 << log_r, 1 << (log_n - log_r)
    cg, rg = C // world, R // world
    a = orc.random_fr(N, 400 + log_n)
    w = pyref.omega_for_k(log_n)
    omega_n = orc.ints_to_mont([w])[0]
    omega_c = orc.ints_to_mont([pow(w, R, pyref.R_MOD)])[0]
    L = zk.lib()
    rows = [backend.alloc(rg * C * 32) for _ in range(world)]
    peers = (ctypes.c_void_p * world)(*[r.ptr.value for r in rows])
    for g in range(world):
        block = np.ascontiguousarray(a.reshape(R, C, 4)[:, g * cg:(g + 1) * cg])
        d = backend.to_device(block)
        backend._check(L.b200zk_fft_colstep_scatter_dev(backend._ctx, d.ptr, ctypes.c_uint32(log_r), ctypes.c_uint32(int(np.log2(cg))),
                                                        ctypes.c_uint32(g * cg), omega_n.ctypes.data_as(ctypes.c_void_p),
                                                        ctypes.c_uint32(log_n), peers, ctypes.c_uint32(world)))
        backend.sync()
        assert np.array_equal(d.download(block.shape), block)                      # the input block is read only
        d.free()
    Z = np.zeros((R, C, 4), dtype=np.uint64)
    for g in range(world):
        backend._check(L.b200zk_fft_rows_dev(backend._ctx, rows[g].ptr, ctypes.c_uint32(rg), omega_c.ctypes.data_as(ctypes.c_void_p),
                                             ctypes.c_uint32(log_n - log_r)))
        Z[g * rg:(g + 1) * rg] = rows[g].download((rg, C, 4))
        rows[g].free()
    got = np.ascontiguousarray(np.transpose(Z, (1, 0, 2))).reshape(N, 4)
    want = orc.best_fft(a, omega_n, log_n) if log_n <= 16 else backend.best_fft(a, omega_n, log_n)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("env", [{"B200ZK_NTT_SHOUP": "0"}, {"B200ZK_NTT_FULL_TW": "0"}, {"B200ZK_NTT_WARP_MAX": "21"},
                                 {"B200ZK_NTT_SHOUP": "0", "B200ZK_NTT_FULL_TW": "0"}])
def test_ntt_fallback_plans_agree(zk, backend, orc, env):
    """The plan switches of INTEGRATION.md §8 select other kernels / multipliers / twiddle tables for the same transform:
    CIOS instead of the constant-operand multiplier in the warp-level kernel, two-level inter-pass twiddles, the block kernel
    above 2^21.  They are what the library falls back to when the big tables do not fit, so each must give the bits of the
    default plan (which the tests above and test_gpu_full_size.py pin to the oracle): sizes with 1, 2, 3 and 4 passes, plus the
    coset extension (batched transforms, coset powers on load)."""
    import os
    from oracle import pyref
    cases = [(k, orc.random_fr(1 << k, 70 + k), orc.ints_to_mont([pyref.omega_for_k(k)])[0]) for k in (7, 13, 16, 20, 22)]
    want = [backend.best_fft(a, w, k) for k, a, w in cases]
    dom = zk.EvaluationDomain(backend, 6, 16)
    coeffs = orc.random_fr(1 << 16, 5)
    want_ext = dom.coeff_to_extended(coeffs)
    dom.close()
    old = {name: os.environ.get(name) for name in env}
    os.environ.update(env)
    try:
        be = zk.Backend(0)                                     # plans are built per context, with the switches read then
        try:
            for (k, a, w), ref in zip(cases, want):
                assert np.array_equal(be.best_fft(a, w, k), ref), (env, k)
            d2 = zk.EvaluationDomain(be, 6, 16)
            assert np.array_equal(d2.coeff_to_extended(coeffs), want_ext), env
            d2.close()
        finally:
            be.close()
    finally:
        for name, v in old.items():
            if v is None:
                os.environ.pop(name, None)
            else:
                os.environ[name] = v
