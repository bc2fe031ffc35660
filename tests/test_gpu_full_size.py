"""GPU parity at BASELINE.json's full sizes (configs 3, 4, 5), through the C ABI:

  * best_multiexp at 2^22 and 2^24 bit-exact against the oracle — both the arbitrary-bases path (b200zk_msm_dev) and
    the fixed-base ParamsKZG::commit path — and at 2^26 as the sum of its four 2^24 quarters, one of them oracle-checked;
  * best_fft at 2^22 and 2^24 bit-exact against the oracle (sizes above 2^21 run the block kernel, a different kernel
    from the warp-level one most other tests reach), 2^26 by Horner spot checks and the inverse round trip;
  * coeff_to_extended / extended_to_coeff round trip at k = 20, 21, 22;
  * create_proof bytes for the reference's Merkle Sum Tree circuit at k = 20 (the headline configuration), the dense
    MST-shaped circuit at k = 20 and the Merkle Sum Tree circuit at k = 21 against the committed sha256 of the
    oracle's proof (tests/golden/full_size_sha256.json, generator tests/golden/make_full_size_digests.py).
"""
import hashlib
import importlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001


def _scalars(n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] %= np.uint64(0x30644e72e131a029)
    return a


@pytest.mark.parametrize("log_n", [22, 24])
def test_best_multiexp_large_matches_oracle(zk, backend, orc, log_n):
    n = 1 << log_n
    params = zk.ParamsKZG.setup(backend, log_n, orc.random_fr(1, 500 + log_n)[0])     # setup itself is oracle-checked at small k
    g, _ = params.read(lagrange=False)
    coeffs = _scalars(n, 600 + log_n)
    want = orc.g1_batch_normalize(orc.best_multiexp(coeffs, g))[0]
    d_c, d_g = backend.to_device(coeffs), backend.to_device(g)
    got_generic = backend.best_multiexp_dev(d_c, d_g, n)                             # arbitrary bases: windows per call
    got_commit = params.commit_dev(d_c, n, lagrange=False)                           # SRS bases: fixed-base tables when they fit
    d_c.free(); d_g.free(); params.close()
    assert np.array_equal(got_generic[:8], want)
    assert np.array_equal(got_commit[:8], want)


def test_best_multiexp_2p26_split_sum(zk, backend, orc):
    """One 2^26-point MSM equals the sum of its four 2^24 quarters (each computed on the GPU), and the first quarter
    equals the oracle's best_multiexp: the full BASELINE config-3 size without a 2^26 CPU run."""
    log_n, n, q = 26, 1 << 26, 1 << 24
    params = zk.ParamsKZG.setup(backend, log_n, orc.random_fr(1, 526)[0])
    coeffs = _scalars(n, 626)
    d_c = backend.to_device(coeffs)
    full = params.commit_dev(d_c, n, lagrange=False)
    g, _ = params.read(lagrange=False)
    params.close()
    d_g = backend.to_device(g)
    parts = []
    for i in range(4):
        dq_c, dq_g = backend.to_device(coeffs[i * q:(i + 1) * q]), backend.to_device(g[i * q:(i + 1) * q])
        parts.append(backend.best_multiexp_dev(dq_c, dq_g, q))
        dq_c.free(); dq_g.free()
    also = backend.best_multiexp_dev(d_c, d_g, n)
    d_c.free(); d_g.free()
    want0 = orc.g1_batch_normalize(orc.best_multiexp(coeffs[:q], g[:q]))[0]
    assert np.array_equal(parts[0][:8], want0)
    total = np.zeros(12, dtype=np.uint64)
    import ctypes
    pts = np.ascontiguousarray(np.stack(parts))
    assert zk.lib().b200zk_g1_sum(pts.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(4), total.ctypes.data_as(ctypes.c_void_p)) == 0
    assert np.array_equal(total, full) and np.array_equal(also, full)


@pytest.mark.parametrize("log_n", [22, 24])
def test_best_fft_large_matches_oracle(backend, orc, log_n):
    from oracle import pyref
    a = _scalars(1 << log_n, 700 + log_n)
    w = orc.ints_to_mont([pyref.omega_for_k(log_n)])[0]
    got = backend.best_fft(a, w, log_n)
    assert np.array_equal(got, orc.best_fft(a, w, log_n))


def test_best_fft_2p26_spots_and_roundtrip(zk, backend, orc):
    from oracle import pyref
    k, n = 26, 1 << 26
    a = _scalars(n, 726)
    w = pyref.omega_for_k(k)
    W, WI = orc.ints_to_mont([w])[0], orc.ints_to_mont([pow(w, -1, R_MOD)])[0]
    d_a, d_in = backend.to_device(a), backend.to_device(a)
    backend.best_fft_dev(d_a, W, k)
    f = d_a.download((n, 4))
    for i in (0, 1, n // 2 + 3, 12345678, n - 1):                        # out[i] = a(omega^i), Horner on the device (oracle-checked primitive)
        x = orc.ints_to_mont([pow(w, i, R_MOD)])[0]
        assert np.array_equal(f[i], backend.eval_polynomial_dev(d_in, n, x))
    assert np.array_equal(f[0], orc.eval_polynomial(a, orc.ints_to_mont([1])[0]))       # and one against the CPU
    backend.best_fft_dev(d_a, WI, k)
    back = d_a.download((n, 4))
    d_a.free(); d_in.free()
    ninv = orc.ints_to_mont([pow(n, -1, R_MOD)])[0]
    assert np.array_equal(orc.binop("mul", back, np.tile(ninv, (n, 1))), a)


@pytest.mark.parametrize("k", [20, 21, 22])
def test_coeff_to_extended_roundtrip_full_size(zk, backend, orc, k):
    """extended_to_coeff(coeff_to_extended(p)) = p || 0 on the 8n extended domain of a degree-6 circuit, plus two
    evaluations of the extended form against Horner: ext[i] = p(zeta * extended_omega^i)."""
    dom = zk.EvaluationDomain(backend, 6, k)
    n, ext = 1 << k, dom.extended_len()
    p = _scalars(n, 800 + k)
    d_p, d_e, d_back = backend.to_device(p), backend.alloc(ext * 32), backend.alloc(dom.quotient_poly_degree * n * 32)
    dom.coeff_to_extended_dev(d_p, d_e)
    spots = d_e.download((ext, 4))[[0, 1, ext - 1]]
    zeta, ew = orc.mont_to_ints(dom.g_coset)[0], orc.mont_to_ints(dom.extended_omega)[0]
    for i, got in zip((0, 1, ext - 1), spots):
        x = orc.ints_to_mont([zeta * pow(ew, i, R_MOD) % R_MOD])[0]
        assert np.array_equal(got, backend.eval_polynomial_dev(d_p, n, x))
    dom.extended_to_coeff_dev(d_e, d_back)
    back = d_back.download((dom.quotient_poly_degree * n, 4))
    d_p.free(); d_e.free(); d_back.free(); dom.close()
    assert np.array_equal(back[:n], p) and not back[n:].any()


def _full_size(zk, backend, orc, name):
    gold = importlib.import_module("tests.golden.make_full_size_digests")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "full_size_sha256.json")
    want = json.load(open(path)).get(name)
    if want is None:
        pytest.skip(f"no committed digest for {name} (run tests/golden/make_full_size_digests.py {name})")
    from oracle import prover as OP
    job = gold.build_job(zk, name)
    params = zk.ParamsKZG.setup(backend, job.k, orc.random_fr(1, gold.SEED_S)[0])
    pk = zk.ProvingKey(params, job.cs, job.k, job.fixed, job.map_col, job.map_row)
    wide = orc.XorShiftWide().draw(pk.rng_draws)
    inst = [orc.ints_to_mont([v % OP.R for v in c]) for c in job.instances]
    proof = pk.create_proof(job.advice, inst, wide, orc.ints_to_mont([job.transcript_repr])[0])
    pk.close(); params.close()
    assert len(proof) == want["proof_bytes"]
    assert hashlib.sha256(proof).hexdigest() == want["sha256"], f"{name}: GPU proof differs from the oracle's committed digest"


def test_proof_bytes_mst_k20(zk, backend, orc):
    """BASELINE's headline configuration: byte-identical to the CPU restatement's proof (committed digest)."""
    _full_size(zk, backend, orc, "mst_k20")


def test_proof_bytes_mst_dense_k20(zk, backend, orc):
    _full_size(zk, backend, orc, "mst_dense_k20")


def test_proof_bytes_mst_k21(zk, backend, orc):
    _full_size(zk, backend, orc, "mst_k21")
