"""TEST INFRASTRUCTURE (oracle): optimal ate pairing on bn256 in plain Python integers.

Lets `oracle/prover.py::verify_full` finish the SHPLONK check the way halo2's verifier does
(`poly/kzg/multiopen/shplonk/verifier.rs` + `DualMSM::check`: e(left, [s]_2) = e(right, [1]_2)),
i.e. from the SRS's G2 elements instead of the secret — the reference's `verify_proof` call at
/root/reference/src/circuits/utils.rs:56-63.  The curve, the twist y^2 = x^3 + 3/(9+u) and the G2
generator are the alt_bn128 constants of EIP-197 (which halo2curves' `bn256` implements); the
tower is Fq12 = Fq[w] / (w^12 - 18 w^6 + 82) with the twist map (x, y) -> (x w^2, y w^3).
Pinned by bilinearity / non-degeneracy / generator-order tests in tests/test_oracle_pairing.py;
never imported by the product.
"""
Q = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
R = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
ATE_LOOP_COUNT = 29793968203157093288          # 6u + 2, u = 4965661367192848881
LOG_ATE = 63

G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))


# ---- Fq2 = Fq[u] / (u^2 + 1), elements (c0, c1) ----
def f2_add(a, b): return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)
def f2_sub(a, b): return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)
def f2_neg(a): return ((-a[0]) % Q, (-a[1]) % Q)


def f2_inv(a):
    d = pow((a[0] * a[0] + a[1] * a[1]) % Q, -1, Q)
    return (a[0] * d % Q, (-a[1]) * d % Q)


B2 = f2_mul((3, 0), f2_inv((9, 1)))            # twist coefficient 3 / (9 + u)


# ---- G2 (affine over Fq2, None = identity) ----
def g2_on_curve(p):
    if p is None:
        return True
    x, y = p
    return f2_sub(f2_mul(y, y), f2_mul(f2_mul(x, x), x)) == B2


def g2_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    if a[0] == b[0]:
        if f2_add(a[1], b[1]) == (0, 0):
            return None
        lam = f2_mul(f2_mul((3, 0), f2_mul(a[0], a[0])), f2_inv(f2_mul((2, 0), a[1])))
    else:
        lam = f2_mul(f2_sub(b[1], a[1]), f2_inv(f2_sub(b[0], a[0])))
    x3 = f2_sub(f2_sub(f2_mul(lam, lam), a[0]), b[0])
    return (x3, f2_sub(f2_mul(lam, f2_sub(a[0], x3)), a[1]))


def g2_mul(p, k):
    k %= R
    acc = None
    while k:
        if k & 1:
            acc = g2_add(acc, p)
        p = g2_add(p, p)
        k >>= 1
    return acc


def g2_neg(p):
    return None if p is None else (p[0], f2_neg(p[1]))


# ---- Fq12 = Fq[w] / (w^12 - 18 w^6 + 82), elements = 12 coefficients ----
def f12_mul(a, b):
    t = [0] * 23
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                t[i + j] += x * y
    for k in range(22, 11, -1):                 # w^k = 18 w^(k-6) - 82 w^(k-12)
        c = t[k]
        if c:
            t[k - 6] += 18 * c
            t[k - 12] -= 82 * c
    return [v % Q for v in t[:12]]


F12_ONE = [1] + [0] * 11


def f12_pow(a, e):
    acc, base = F12_ONE, a
    while e:
        if e & 1:
            acc = f12_mul(acc, base)
        base = f12_mul(base, base)
        e >>= 1
    return acc


def f12_inv(a):
    """Extended Euclid in Fq[w] against the modulus polynomial."""
    def deg(p):
        d = len(p) - 1
        while d and p[d] == 0:
            d -= 1
        return d

    def poly_divmod_step(lm, low, hm, high):
        r = [0] * 13
        dl, dh = deg(low), deg(high)
        tmp = list(high)
        for i in range(dh - dl, -1, -1):
            r[i] = tmp[dl + i] * pow(low[dl], -1, Q) % Q
            for c in range(dl + 1):
                tmp[c + i] = (tmp[c + i] - r[i] * low[c]) % Q
        return r

    lm, hm = [1] + [0] * 12, [0] * 13
    low, high = list(a) + [0], [82, 0, 0, 0, 0, 0, (-18) % Q, 0, 0, 0, 0, 0, 1]
    while deg(low):
        r = poly_divmod_step(lm, low, hm, high)
        nm, new = list(hm), list(high)
        for i in range(13):
            for j in range(13 - i):
                nm[i + j] = (nm[i + j] - lm[i] * r[j]) % Q
                new[i + j] = (new[i + j] - low[i] * r[j]) % Q
        lm, low, hm, high = nm, new, lm, low
    inv0 = pow(low[0], -1, Q)
    return [c * inv0 % Q for c in lm[:12]]


def f12_sub(a, b): return [(x - y) % Q for x, y in zip(a, b)]
def f12_add(a, b): return [(x + y) % Q for x, y in zip(a, b)]
def f12_scalar(c): return [c % Q] + [0] * 11


def twist(p):
    """G2 point over Fq2 -> point over Fq12 on y^2 = x^3 + 3."""
    (x0, x1), (y0, y1) = p
    nx = [0] * 12
    ny = [0] * 12
    nx[0], nx[6] = (x0 - 9 * x1) % Q, x1         # a + b u  ->  (a - 9b) + b w^6
    ny[0], ny[6] = (y0 - 9 * y1) % Q, y1
    w2 = [0, 0, 1] + [0] * 9
    w3 = [0, 0, 0, 1] + [0] * 8
    return (f12_mul(nx, w2), f12_mul(ny, w3))


def _p12_double(p):
    x, y = p
    lam = f12_mul(f12_mul(f12_scalar(3), f12_mul(x, x)), f12_inv(f12_mul(f12_scalar(2), y)))
    nx = f12_sub(f12_sub(f12_mul(lam, lam), x), x)
    return (nx, f12_sub(f12_mul(lam, f12_sub(x, nx)), y))


def _p12_add(a, b):
    if a[0] == b[0]:
        return _p12_double(a)
    lam = f12_mul(f12_sub(b[1], a[1]), f12_inv(f12_sub(b[0], a[0])))
    nx = f12_sub(f12_sub(f12_mul(lam, lam), a[0]), b[0])
    return (nx, f12_sub(f12_mul(lam, f12_sub(a[0], nx)), a[1]))


def _linefunc(p1, p2, t):
    """Line through p1, p2 (tangent if equal) evaluated at t; all over Fq12."""
    x1, y1 = p1
    x2, y2 = p2
    xt, yt = t
    if x1 != x2:
        m = f12_mul(f12_sub(y2, y1), f12_inv(f12_sub(x2, x1)))
        return f12_sub(f12_mul(m, f12_sub(xt, x1)), f12_sub(yt, y1))
    if y1 == y2:
        m = f12_mul(f12_mul(f12_scalar(3), f12_mul(x1, x1)), f12_inv(f12_mul(f12_scalar(2), y1)))
        return f12_sub(f12_mul(m, f12_sub(xt, x1)), f12_sub(yt, y1))
    return f12_sub(xt, x1)


def miller_loop(q2, p1):
    """q2 in G2 (affine over Fq2), p1 in G1 (affine integers); no final exponentiation."""
    if q2 is None or p1 is None:
        return F12_ONE
    Qt = twist(q2)
    Pt = (f12_scalar(p1[0]), f12_scalar(p1[1]))
    Rt, f = Qt, F12_ONE
    for i in range(LOG_ATE, -1, -1):
        f = f12_mul(f12_mul(f, f), _linefunc(Rt, Rt, Pt))
        Rt = _p12_double(Rt)
        if ATE_LOOP_COUNT & (1 << i):
            f = f12_mul(f, _linefunc(Rt, Qt, Pt))
            Rt = _p12_add(Rt, Qt)
    Q1 = (f12_pow(Qt[0], Q), f12_pow(Qt[1], Q))
    nQ2 = (f12_pow(Q1[0], Q), [(-c) % Q for c in f12_pow(Q1[1], Q)])
    f = f12_mul(f, _linefunc(Rt, Q1, Pt))
    Rt = _p12_add(Rt, Q1)
    f = f12_mul(f, _linefunc(Rt, nQ2, Pt))
    return f


def final_exponentiate(f):
    return f12_pow(f, (Q ** 12 - 1) // R)


def pairing(q2, p1):
    return final_exponentiate(miller_loop(q2, p1))


def pairing_product_is_one(pairs):
    """prod e(P_i, Q_i) == 1 for pairs (p1, q2), with one shared final exponentiation."""
    f = F12_ONE
    for p1, q2 in pairs:
        f = f12_mul(f, miller_loop(q2, p1))
    return final_exponentiate(f) == F12_ONE
