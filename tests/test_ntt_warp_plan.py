"""Pass plans of the warp-level NTT kernel (csrc/ntt_plan.hpp::ntt_plan_shape_warp, csrc/ntt_warp.cuh), checked on the CPU:
the digit split, and — in a 31-bit prime field with numpy — that the kernel's address algebra (strided non-last passes with
the inter-pass twiddle omega^(l k << tw_shift), last pass over contiguous rows storing k1 + M1 k2 (+ M1 M2 k3) + M1 M_mid k with
the two middle digits of a four-pass plan swapped) is the natural-order transform of halo2's best_fft.  The field arithmetic
itself is covered by test_limb_emulation.py and, on the GPU, bit-exactly against the oracle (test_gpu_parity / test_gpu_full_size)."""
import numpy as np
import pytest

from tests.emu import binding as emu

P = 2013265921            # 15 * 2^27 + 1
G = 31                    # generator of the multiplicative group


def dft_axis(x, w, axis):
    """Natural-order DFT of length M = x.shape[axis] along `axis` with root w (radix-2 DIF + bit reversal), mod P."""
    x = np.moveaxis(x, axis, 0).copy()
    M = x.shape[0]
    logm = M.bit_length() - 1
    half, wl = M // 2, w
    while half >= 1:
        xv = x.reshape((M // (2 * half), 2, half) + x.shape[1:])
        tw = np.array([pow(wl, j, P) for j in range(half)], dtype=np.int64).reshape((1, half) + (1,) * (x.ndim - 1))
        a, b = xv[:, 0].copy(), xv[:, 1].copy()
        xv[:, 0] = (a + b) % P
        xv[:, 1] = (a - b) % P * tw % P
        half //= 2
        wl = wl * wl % P
    rev = np.array([int(format(i, f"0{logm}b")[::-1], 2) if logm else 0 for i in range(M)])
    return np.moveaxis(x[rev], 0, axis)


@pytest.mark.parametrize("log_n", [12, 14, 20, 21, 22, 23, 24, 25, 26])
def test_plan_digits(log_n):
    plan = emu.ntt_warp_plan(log_n)
    assert plan is not None and len(plan) == -(-log_n // 7)
    d = [q["log_m"] for q in plan]
    assert sum(d) == log_n and all(4 <= m <= 7 for m in d)
    log_l = log_n
    for i, q in enumerate(plan):
        log_l -= q["log_m"]
        assert q["log_l"] == log_l and q["log_tw"] == 7 - q["log_m"] and q["blocks"] == 1 << (log_n - 7)
        assert q["is_last"] == (i == len(plan) - 1)
        assert q["log_m1"] == (d[0] if len(plan) > 1 else 0)
        assert q["log_mid"] == sum(d[1:-1]) and q["log_m3"] == (d[2] if len(plan) == 4 else 0)
    assert emu.ntt_warp_plan(11) is None and emu.ntt_warp_plan(27) is None      # default ceiling 2^26 (B200ZK_NTT_WARP_MAX)


@pytest.mark.parametrize("log_n", [14, 20, 22])
def test_address_algebra_is_the_natural_order_transform(log_n):
    plan = emu.ntt_warp_plan(log_n)
    N = 1 << log_n
    omega = pow(G, (P - 1) >> log_n, P)
    rng = np.random.default_rng(log_n)
    a = rng.integers(0, P, size=N, dtype=np.int64)
    buf = a.copy()
    for q in plan[:-1]:                                       # non-last pass: element (h, m, l) at h M L + m L + l
        M, L = 1 << q["log_m"], 1 << q["log_l"]
        x = buf.reshape(-1, M, L)
        x = dft_axis(x, pow(omega, N // M, P), 1)             # tile transform along m; k lands where m was
        shift = log_n - q["log_m"] - q["log_l"]               # a.tw_shift
        wsub = pow(omega, 1 << shift, P)                      # omega^(l k << shift) = wsub^(l k)
        col = np.ones(L, dtype=np.int64)                      # wsub^l
        base = 1
        for l in range(1, L):
            base = base * wsub % P
            col[l] = base
        tw = np.ones((M, L), dtype=np.int64)
        for k in range(1, M):
            tw[k] = tw[k - 1] * col % P
        buf = (x * tw[None] % P).reshape(-1)
    q = plan[-1]                                              # last pass: contiguous rows, stores digit-reversed
    M = 1 << q["log_m"]
    x = dft_axis(buf.reshape(-1, M), pow(omega, N // M, P), 1)
    rho = np.arange(N // M, dtype=np.int64)
    k1, rho_mid = rho >> q["log_mid"], rho & ((1 << q["log_mid"]) - 1)
    base = k1 + ((rho_mid >> q["log_m3"]) << q["log_m1"]) + ((rho_mid & ((1 << q["log_m3"]) - 1)) << (q["log_m1"] + q["log_mid"] - q["log_m3"]))
    out = np.zeros(N, dtype=np.int64)
    for k in range(M):
        out[base + (k << (q["log_m1"] + q["log_mid"]))] = x[:, k]
    # spot checks against the definition X[K] = sum_i a[i] omega^(i K)
    for K in [0, 1, 2, N // 2, N - 1] + [int(v) for v in rng.integers(0, N, size=6)]:
        wk = pow(omega, K, P)
        pw = np.ones(N, dtype=np.int64)                       # wk^i by doubling
        step, filled = wk, 1
        while filled < N:
            pw[filled:2 * filled] = pw[:filled] * step % P
            step = step * step % P
            filled *= 2
        assert int((a * pw % P).sum() % P) == int(out[K])
