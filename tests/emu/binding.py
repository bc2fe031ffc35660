"""ctypes binding for the host-emulation build of the device headers (test aid)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libemu.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
        _lib = ctypes.CDLL(_LIB)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


OPS = {"add": 0, "sub": 1, "mul": 2, "sqr": 3, "inv": 4, "from_mont": 5, "to_mont": 6, "inv_gcd": 7}


def field_op(which, op, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = a if b is None else np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros_like(a)
    lib().emu_field_op(which, OPS[op], _p(a), _p(b), _p(out), ctypes.c_size_t(a.shape[0]))
    return out


def field_consts(which):
    one = np.zeros((1, 4), dtype=np.uint64)
    r2 = np.zeros((1, 4), dtype=np.uint64)
    lib().emu_field_consts(which, _p(one), _p(r2))
    return one[0], r2[0]


def ntt(a, log_n, omega, n_in=None, max_log_m=10, max_log_tw=2, tile_cap_log=12, nthreads=64, pre=None, post=None):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    n_in = a.shape[0] if n_in is None else n_in
    out = np.zeros((1 << log_n, 4), dtype=np.uint64)
    pre = None if pre is None else np.ascontiguousarray(pre, dtype=np.uint64)
    post = None if post is None else np.ascontiguousarray(post, dtype=np.uint64)
    lib().emu_ntt(_p(a), ctypes.c_uint32(n_in), _p(out), ctypes.c_uint32(log_n), _p(np.ascontiguousarray(omega)),
                  ctypes.c_uint32(max_log_m), ctypes.c_uint32(max_log_tw), ctypes.c_uint32(tile_cap_log),
                  ctypes.c_uint32(nthreads), _p(pre), _p(post))
    return out


def msm(scalars, bases, force_c=0, fast_max=96, seg_len=32):
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
    out = np.zeros(8, dtype=np.uint64)
    lib().emu_msm(_p(scalars), _p(bases), ctypes.c_uint32(scalars.shape[0]), ctypes.c_int(force_c),
                  ctypes.c_int(fast_max), ctypes.c_uint32(seg_len), _p(out))
    return out


def host_field_op(which, op, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = a if b is None else np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros_like(a)
    lib().emu_host_field_op(which, {"add": 0, "sub": 1, "mul": 2, "inv": 3}[op], _p(a), _p(b), _p(out), ctypes.c_size_t(a.shape[0]))
    return out


def host_fr_consts(wide):
    r, z, d, w = (np.zeros(4, dtype=np.uint64) for _ in range(4))
    wide = np.ascontiguousarray(wide, dtype=np.uint64)
    lib().emu_host_fr_consts(_p(r), _p(z), _p(d), _p(wide), _p(w))
    return r, z, d, w


def batch_invert(which, a, lanes):
    a = np.array(a, dtype=np.uint64).reshape(-1, 4)
    lib().emu_batch_invert(which, _p(a), ctypes.c_size_t(a.shape[0]), ctypes.c_size_t(lanes))
    return a


def recurrence(a, b, m, T, want_y=True):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    y = np.zeros_like(a) if want_y else None
    head = np.zeros(4, dtype=np.uint64)
    lib().emu_recurrence(_p(a), _p(y), ctypes.c_size_t(a.shape[0]), ctypes.c_size_t(m), ctypes.c_uint32(T),
                         _p(np.ascontiguousarray(b)), _p(head))
    return y, head


def prefix_product(p, z0, m, T):
    p = np.ascontiguousarray(p, dtype=np.uint64).reshape(-1, 4)
    z = np.zeros_like(p)
    lib().emu_prefix_product(_p(p), _p(z), ctypes.c_size_t(p.shape[0]), ctypes.c_size_t(m), ctypes.c_uint32(T),
                             _p(np.ascontiguousarray(z0)))
    return z


def transcript(ops, data, n_squeeze):
    ops = np.asarray(ops, dtype=np.uint8)
    data = np.ascontiguousarray(data, dtype=np.uint64).reshape(-1)
    out = np.zeros((max(n_squeeze, 1), 4), dtype=np.uint64)
    proof = np.zeros(32 * len(ops) + 32, dtype=np.uint8)
    lib().emu_transcript.restype = ctypes.c_size_t
    ln = lib().emu_transcript(_p(ops), ctypes.c_size_t(len(ops)), _p(data), _p(out), _p(proof))
    return out[:n_squeeze], bytes(proof[:ln])


def from_u512(wide):
    wide = np.ascontiguousarray(wide, dtype=np.uint64).reshape(-1, 8)
    out = np.zeros((wide.shape[0], 4), dtype=np.uint64)
    lib().emu_from_u512(_p(wide), _p(out), ctypes.c_size_t(wide.shape[0]))
    return out


def msm_pre(scalars, bases, c, seg_len=3):
    """Fixed-base (precomputed window table) MSM path; len(scalars) <= len(bases) = table stride."""
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
    out = np.zeros(8, dtype=np.uint64)
    lib().emu_msm_pre(_p(scalars), _p(bases), ctypes.c_uint32(scalars.shape[0]), ctypes.c_uint32(bases.shape[0]),
                      ctypes.c_uint32(c), ctypes.c_uint32(seg_len), _p(out))
    return out


def lookup_extrapolate(g, inv_pow, lam, tinv, CL, C, n):
    g = np.array(g, dtype=np.uint64).reshape(C * n, 4)
    lib().emu_lookup_extrapolate(_p(g), _p(np.ascontiguousarray(inv_pow, dtype=np.uint64)), _p(np.ascontiguousarray(lam, dtype=np.uint64)),
                                 _p(np.ascontiguousarray(tinv, dtype=np.uint64)), ctypes.c_uint32(CL), ctypes.c_uint32(C), ctypes.c_size_t(n))
    return g


def coset_interpolate(g, inv_pow, vinv, C, n, extra=None):
    out = np.zeros((C * n, 4), dtype=np.uint64)
    ex = None if extra is None else np.ascontiguousarray(extra, dtype=np.uint64)
    lib().emu_coset_interpolate(_p(np.ascontiguousarray(g, dtype=np.uint64)), _p(np.ascontiguousarray(inv_pow, dtype=np.uint64)),
                                _p(np.ascontiguousarray(vinv, dtype=np.uint64)), ctypes.c_uint32(C), ctypes.c_size_t(n), _p(out), _p(ex) if ex is not None else None)
    return out


def mul_shoup(which, x, w, wq):
    """Field::mul_shoup: x * w mod p with w plain and wq = floor(w * 2^256 / p); raw limb arrays (n, 4) uint64."""
    x = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 4)
    w = np.ascontiguousarray(w, dtype=np.uint64).reshape(-1, 4)
    wq = np.ascontiguousarray(wq, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros_like(x)
    lib().emu_mul_shoup(which, _p(x), _p(w), _p(wq), _p(out), ctypes.c_size_t(x.shape[0]))
    return out


def lazy_op(which, mode, x, w=None, wq=None):
    """The [0, 2p) forms of a transform pass (Field::mul_shoup_lazy / add_lazy / sub_raw / reduce_2p); raw limb arrays."""
    x = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 4)
    w = x if w is None else np.ascontiguousarray(w, dtype=np.uint64).reshape(-1, 4)
    wq = x if wq is None else np.ascontiguousarray(wq, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros_like(x)
    lib().emu_lazy_op(which, {"mul": 0, "add": 1, "sub": 2, "reduce": 3}[mode], _p(x), _p(w), _p(wq), _p(out), ctypes.c_size_t(x.shape[0]))
    return out


def ntt_warp_plan(log_n):
    """Pass plan of the warp-level NTT kernel: list of dicts (log_m, log_l, log_tw, is_last, log_m1, log_mid, log_m3, blocks), or None."""
    out = np.zeros(1 + 8 * 4, dtype=np.uint32)
    if not lib().emu_ntt_warp_plan(ctypes.c_uint32(log_n), _p(out)):
        return None
    keys = ("log_m", "log_l", "log_tw", "is_last", "log_m1", "log_mid", "log_m3", "blocks")
    return [dict(zip(keys, (int(v) for v in out[1 + 8 * p: 9 + 8 * p]))) for p in range(int(out[0]))]
