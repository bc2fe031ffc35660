"""ORACLE — test infrastructure only.  ctypes binding over oracle/_build/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Arrays are numpy uint64 of shape (n, 4) (Fr/Fq,
Montgomery limbs), (n, 8) (G1Affine) or (n, 12) (G1 Jacobian).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

FR, FQ = 0, 1


def build(force=False):
    """Compile the oracle with gcc (no CUDA involved)."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_domain_new.restype = ctypes.c_void_p
        _lib.orc_init()
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _arr(n, w=4):
    return np.zeros((n, w), dtype=np.uint64)


def set_threads(t):
    lib().orc_set_threads(int(t))


def get_threads():
    return int(lib().orc_get_threads())


# ---- conversions ----
def ints_to_raw(vals):
    out = _arr(len(vals))
    for i, v in enumerate(vals):
        for j in range(4):
            out[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def raw_to_ints(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    return [sum(int(a[i, j]) << (64 * j) for j in range(4)) for i in range(a.shape[0])]


def from_raw(raw, which=FR):
    raw = np.ascontiguousarray(raw, dtype=np.uint64).reshape(-1, 4)
    out = _arr(raw.shape[0])
    lib().orc_f_from_raw(which, _p(raw), _p(out), ctypes.c_size_t(raw.shape[0]))
    return out


def to_raw(a, which=FR):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    out = _arr(a.shape[0])
    lib().orc_f_to_raw(which, _p(a), _p(out), ctypes.c_size_t(a.shape[0]))
    return out


def ints_to_mont(vals, which=FR):
    return from_raw(ints_to_raw(vals), which)


def mont_to_ints(a, which=FR):
    return raw_to_ints(to_raw(a, which))


def from_u512(wide, which=FR):
    wide = np.ascontiguousarray(wide, dtype=np.uint64).reshape(-1, 8)
    out = _arr(wide.shape[0])
    lib().orc_f_from_u512(which, _p(wide), _p(out), ctypes.c_size_t(wide.shape[0]))
    return out


def random_fr(n, seed):
    """n uniform Fr elements (Montgomery), numpy PCG64 -> 512-bit -> mod r (bench/test inputs)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    wide = rng.integers(0, 1 << 64, size=(n, 8), dtype=np.uint64)
    return from_u512(wide, FR)


# ---- field ops ----
def binop(op, a, b, which=FR):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    out = _arr(a.shape[0])
    lib().orc_f_binop(which, {"add": 0, "sub": 1, "mul": 2}[op], _p(a), _p(b), _p(out), ctypes.c_size_t(a.shape[0]))
    return out


def inv(a, which=FR):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    out = _arr(a.shape[0])
    lib().orc_f_inv(which, _p(a), _p(out), ctypes.c_size_t(a.shape[0]))
    return out


def batch_invert(a):
    a = np.array(a, dtype=np.uint64).reshape(-1, 4)
    lib().orc_fr_batch_invert(_p(a), ctypes.c_size_t(a.shape[0]))
    return a


def fr_constants():
    r, d, z = _arr(1), _arr(1), _arr(1)
    lib().orc_fr_constants(_p(r), _p(d), _p(z))
    return r[0], d[0], z[0]


def field_params(which):
    p, r, r2, r3 = _arr(1), _arr(1), _arr(1), _arr(1)
    inv_ = np.zeros(1, dtype=np.uint64)
    lib().orc_field_params(which, _p(p), _p(inv_), _p(r), _p(r2), _p(r3))
    return dict(p=p[0], inv=int(inv_[0]), r=r[0], r2=r2[0], r3=r3[0])


# ---- G1 ----
def g1_generator():
    out = _arr(1, 8)
    lib().orc_g1_generator(_p(out))
    return out[0]


def g1_from_affine(aff):
    aff = np.ascontiguousarray(aff, dtype=np.uint64).reshape(-1, 8)
    out = _arr(aff.shape[0], 12)
    lib().orc_g1_from_affine(_p(aff), _p(out), ctypes.c_size_t(aff.shape[0]))
    return out


def g1_add(a, b):
    out = _arr(1, 12)
    lib().orc_g1_add(_p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(out))
    return out[0]


def g1_add_affine(a, b_aff):
    out = _arr(1, 12)
    lib().orc_g1_add_affine(_p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b_aff)), _p(out))
    return out[0]


def g1_double(a):
    out = _arr(1, 12)
    lib().orc_g1_double(_p(np.ascontiguousarray(a)), _p(out))
    return out[0]


def g1_mul(a, fr_scalar):
    out = _arr(1, 12)
    lib().orc_g1_mul(_p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(fr_scalar)), _p(out))
    return out[0]


def g1_batch_normalize(jac):
    jac = np.ascontiguousarray(jac, dtype=np.uint64).reshape(-1, 12)
    out = _arr(jac.shape[0], 8)
    lib().orc_g1_batch_normalize(_p(jac), _p(out), ctypes.c_size_t(jac.shape[0]))
    return out


def g1_compress(aff):
    aff = np.ascontiguousarray(aff, dtype=np.uint64).reshape(-1, 8)
    out = np.zeros((aff.shape[0], 32), dtype=np.uint8)
    lib().orc_g1_compress(_p(aff), _p(out), ctypes.c_size_t(aff.shape[0]))
    return out


def g1_on_curve(aff):
    aff = np.ascontiguousarray(aff, dtype=np.uint64).reshape(-1, 8)
    return bool(lib().orc_g1_on_curve(_p(aff), ctypes.c_size_t(aff.shape[0])))


def g1_fixed_base_mul(scalars):
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    out = _arr(scalars.shape[0], 8)
    lib().orc_g1_fixed_base_mul(_p(scalars), _p(out), ctypes.c_size_t(scalars.shape[0]))
    return out


def affine_to_ints(aff):
    """(n,8) Montgomery affine -> list of (x, y) ints or None for identity."""
    aff = np.ascontiguousarray(aff, dtype=np.uint64).reshape(-1, 8)
    xs = mont_to_ints(aff[:, :4], FQ)
    ys = mont_to_ints(aff[:, 4:], FQ)
    return [None if (x == 0 and y == 0) else (x, y) for x, y in zip(xs, ys)]


def params_setup(k, s_mont, with_lagrange=True):
    n = 1 << k
    g = _arr(n, 8)
    gl = _arr(n, 8) if with_lagrange else None
    lib().orc_params_setup(ctypes.c_uint(k), _p(np.ascontiguousarray(s_mont)), _p(g), _p(gl) if with_lagrange else None)
    return g, gl


# ---- arithmetic.rs ----
def best_multiexp(coeffs, bases):
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
    assert coeffs.shape[0] == bases.shape[0]
    out = _arr(1, 12)
    lib().orc_best_multiexp(_p(coeffs), _p(bases), ctypes.c_size_t(coeffs.shape[0]), _p(out))
    return out[0]


def best_fft(a, omega, log_n):
    a = np.array(a, dtype=np.uint64).reshape(-1, 4)
    assert a.shape[0] == 1 << log_n
    lib().orc_best_fft(_p(a), _p(np.ascontiguousarray(omega)), ctypes.c_uint(log_n))
    return a


def eval_polynomial(poly, x):
    poly = np.ascontiguousarray(poly, dtype=np.uint64).reshape(-1, 4)
    out = _arr(1)
    lib().orc_eval_polynomial(_p(poly), ctypes.c_size_t(poly.shape[0]), _p(np.ascontiguousarray(x)), _p(out))
    return out[0]


def kate_division(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    q = _arr(a.shape[0] - 1)
    lib().orc_kate_division(_p(a), ctypes.c_size_t(a.shape[0]), _p(np.ascontiguousarray(b)), _p(q))
    return q


def binop_scalar(op, a, s):
    """a (op) s with the Fr scalar s broadcast."""
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    out = _arr(a.shape[0])
    lib().orc_fr_binop_scalar({"add": 0, "sub": 1, "mul": 2}[op], _p(a), _p(np.ascontiguousarray(s, dtype=np.uint64)), _p(out),
                              ctypes.c_size_t(a.shape[0]))
    return out


class XorShiftWide:
    """rand_xorshift::XorShiftRng producing the 512-bit inputs of Fr::random, in draw order."""

    def __init__(self, seed=bytes([0x59, 0x62, 0xbe, 0x5d, 0x76, 0x3d, 0x31, 0x8d, 0x17, 0xdb, 0x37, 0x32, 0x54, 0x06, 0xbc, 0xe5])):
        self.state = np.frombuffer(bytes(seed), dtype=np.uint32).copy()

    def draw(self, count):
        out = np.zeros((count, 8), dtype=np.uint64)
        lib().orc_xorshift_fill_wide(_p(self.state), _p(out), ctypes.c_size_t(count))
        return out


def prefix_product(p, z0):
    p = np.ascontiguousarray(p, dtype=np.uint64).reshape(-1, 4)
    z = _arr(p.shape[0])
    lib().orc_prefix_product(_p(p), _p(np.ascontiguousarray(z0, dtype=np.uint64)), _p(z), ctypes.c_size_t(p.shape[0]))
    return z


def powers(base, n):
    out = _arr(n)
    lib().orc_powers(_p(np.ascontiguousarray(base, dtype=np.uint64)), _p(out), ctypes.c_size_t(n))
    return out


def permute_expression_pair(inp, tab, usable):
    inp = np.ascontiguousarray(inp, dtype=np.uint64).reshape(-1, 4)
    tab = np.ascontiguousarray(tab, dtype=np.uint64).reshape(-1, 4)
    a, s = _arr(usable), _arr(usable)
    rc = lib().orc_permute_expression_pair(_p(inp), _p(tab), ctypes.c_size_t(usable), _p(a), _p(s))
    if rc:
        raise ValueError("ConstraintSystemFailure: lookup input not in table")
    return a, s


def sort_canonical(raw):
    raw = np.array(raw, dtype=np.uint64).reshape(-1, 4)
    lib().orc_sort_canonical(_p(raw), ctypes.c_size_t(raw.shape[0]))
    return raw


class Domain:
    """poly::EvaluationDomain<Fr>::new(j, k)."""
    _FIELDS = ["omega", "omega_inv", "extended_omega", "extended_omega_inv", "g_coset", "g_coset_inv",
               "ifft_divisor", "extended_ifft_divisor", "barycentric_weight"]

    def __init__(self, j, k):
        self._h = ctypes.c_void_p(lib().orc_domain_new(ctypes.c_uint(j), ctypes.c_uint(k)))
        self.k, self.n = k, 1 << k
        self.extended_k = int(lib().orc_domain_extended_k(self._h))
        self.quotient_poly_degree = int(lib().orc_domain_quotient_degree(self._h))
        for i, name in enumerate(self._FIELDS):
            v = _arr(1)
            lib().orc_domain_get(self._h, i, _p(v))
            setattr(self, name, v[0])
        t = _arr(1 << (self.extended_k - k))
        lib().orc_domain_t_evaluations(self._h, _p(t))
        self.t_evaluations = t

    def __del__(self):
        try:
            lib().orc_domain_free(self._h)
        except Exception:
            pass

    def extended_len(self):
        return 1 << self.extended_k

    def lagrange_to_coeff(self, a):
        a = np.array(a, dtype=np.uint64).reshape(self.n, 4)
        lib().orc_domain_lagrange_to_coeff(self._h, _p(a))
        return a

    def coeff_to_extended(self, a):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(self.n, 4)
        out = _arr(self.extended_len())
        lib().orc_domain_coeff_to_extended(self._h, _p(a), _p(out))
        return out

    def extended_to_coeff(self, a):
        a = np.array(a, dtype=np.uint64).reshape(self.extended_len(), 4)
        lib().orc_domain_extended_to_coeff(self._h, _p(a))
        return a[: self.n * self.quotient_poly_degree]

    def divide_by_vanishing_poly(self, a):
        a = np.array(a, dtype=np.uint64).reshape(self.extended_len(), 4)
        lib().orc_domain_divide_by_vanishing_poly(self._h, _p(a))
        return a

    def rotate_omega(self, x, rot):
        out = _arr(1)
        lib().orc_domain_rotate_omega(self._h, _p(np.ascontiguousarray(x)), ctypes.c_int(rot), _p(out))
        return out[0]


# ---- threaded C++ restatement of keygen_pk + create_proof (oracle/prover.cpp) ----
class CppProvingKey:
    """keygen_pk of oracle/prover.cpp: same inputs as oracle.prover.keygen_pk (cs must offer to_blob(k))."""

    def __init__(self, cs, k, fixed, map_col, map_row):
        L = lib()
        L.orc_pk_create.restype = ctypes.c_void_p
        L.orc_pk_rng_draws.restype = ctypes.c_size_t
        self.k, self.n = k, 1 << k
        blob = np.ascontiguousarray(cs.to_blob(k), dtype=np.uint32)
        cols = [np.ascontiguousarray(f, dtype=np.uint64).reshape(self.n, 4) for f in fixed]
        ptrs = (ctypes.c_void_p * max(len(cols), 1))(*[c.ctypes.data for c in cols])
        mc = None if map_col is None else np.ascontiguousarray(map_col, dtype=np.uint32)
        mr = None if map_row is None else np.ascontiguousarray(map_row, dtype=np.uint32)
        self._h = L.orc_pk_create(_p(blob), ctypes.c_size_t(blob.shape[0]), ptrs, None if mc is None else _p(mc), None if mr is None else _p(mr))
        if not self._h:
            raise ValueError("orc_pk_create: malformed constraint-system blob")
        self.rng_draws = int(L.orc_pk_rng_draws(ctypes.c_void_p(self._h)))
        self.num_advice, self.num_instance = cs.num_advice, cs.num_instance

    def close(self):
        if self._h:
            lib().orc_pk_destroy(ctypes.c_void_p(self._h))
            self._h = None

    def create_proof(self, g, g_lagrange, advice, instances, rng_wide, transcript_repr):
        """plonk::create_proof; instances = lists of Python ints, transcript_repr = Python int (as oracle.prover)."""
        n = self.n
        adv = [np.ascontiguousarray(a, dtype=np.uint64).reshape(n, 4) for a in advice]
        adv_ptrs = (ctypes.c_void_p * max(len(adv), 1))(*[a.ctypes.data for a in adv])
        P = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
        inst = [ints_to_mont([v % P for v in col]) if len(col) else np.zeros((1, 4), dtype=np.uint64) for col in instances]
        inst_ptrs = (ctypes.c_void_p * max(len(inst), 1))(*[a.ctypes.data for a in inst])
        lens = np.array([len(col) for col in instances] + [0], dtype=np.uint32)
        wide = np.ascontiguousarray(rng_wide, dtype=np.uint64).reshape(-1, 8)
        g = np.ascontiguousarray(g, dtype=np.uint64)
        gl = np.ascontiguousarray(g_lagrange, dtype=np.uint64)
        repr_m = ints_to_mont([transcript_repr % P])
        out = np.zeros(1 << 16, dtype=np.uint8)
        ln = ctypes.c_size_t()
        rc = lib().orc_create_proof(ctypes.c_void_p(self._h), _p(g), _p(gl), adv_ptrs, inst_ptrs, _p(lens), _p(wide), ctypes.c_size_t(wide.shape[0]),
                                    _p(repr_m), _p(out), ctypes.c_size_t(out.shape[0]), ctypes.byref(ln))
        if rc == 1:
            raise ValueError("ConstraintSystemFailure: lookup input not in table")
        if rc:
            raise ValueError(f"orc_create_proof failed with code {rc}")
        return bytes(out[: ln.value])
