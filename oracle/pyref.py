"""ORACLE PIN — test infrastructure only.

Pure-Python big-int restatement of the bn256 arithmetic that the C++ oracle
(oracle/*.cpp) is pinned against.  Independent of the C++ code on purpose: it
shares no limb arithmetic with it, so agreement of the two is evidence that the
Montgomery/limb code is right.  PARITY UNPINNED against the Rust reference
(halo2curves 0.3.1 / halo2_proofs v2023_02_02 are not available here); the
constants below are the published bn256 parameters.

Reference call sites: /root/reference/src/circuits/utils.rs:2 (Fr, G1Affine).
"""
import hashlib

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001  # Fr modulus
Q_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47  # Fq modulus
MONT_R = 1 << 256
FR_S = 28
FR_GENERATOR = 7
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (R_MOD - 1) >> FR_S, R_MOD)
FR_DELTA = pow(FR_GENERATOR, 1 << FR_S, R_MOD)
FR_ZETA = 0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23
G1_GEN = (1, 2)
CURVE_B = 3


def to_mont(v, p):
    return (v * MONT_R) % p


def from_mont(v, p):
    return (v * pow(MONT_R, -1, p)) % p


def limbs(v):
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def from_limbs(l):
    return sum(int(x) << (64 * i) for i, x in enumerate(l))


# ---- G1 (affine tuples, None = identity) ----
def g1_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % Q_MOD == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, Q_MOD) % Q_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, Q_MOD) % Q_MOD
    x3 = (lam * lam - x1 - x2) % Q_MOD
    return (x3, (lam * (x1 - x3) - y1) % Q_MOD)


def g1_mul(a, k):
    acc = None
    k %= R_MOD
    while k:
        if k & 1:
            acc = g1_add(acc, a)
        a = g1_add(a, a)
        k >>= 1
    return acc


def g1_compress(a):
    if a is None:
        return bytes(32)
    b = bytearray(a[0].to_bytes(32, "little"))
    b[31] |= (a[1] & 1) << 7
    return bytes(b)


def msm(scalars, points):
    acc = None
    for s, p in zip(scalars, points):
        acc = g1_add(acc, g1_mul(p, s))
    return acc


# ---- Fr NTT (definition: out[i] = sum_j a[j] * omega^(i*j)) ----
def omega_for_k(k):
    return pow(FR_ROOT_OF_UNITY, 1 << (FR_S - k), R_MOD)


def ntt_naive(a, omega):
    n = len(a)
    return [sum(a[j] * pow(omega, i * j, R_MOD) for j in range(n)) % R_MOD for i in range(n)]


def ntt(a, omega):
    n = len(a)
    if n == 1:
        return list(a)
    w2 = omega * omega % R_MOD
    ev, od = ntt(a[0::2], w2), ntt(a[1::2], w2)
    out = [0] * n
    w = 1
    for i in range(n // 2):
        t = w * od[i] % R_MOD
        out[i] = (ev[i] + t) % R_MOD
        out[i + n // 2] = (ev[i] - t) % R_MOD
        w = w * omega % R_MOD
    return out


# ---- EvaluationDomain (poly/domain.rs) ----
class Domain:
    def __init__(self, j, k):
        self.k = k
        self.n = 1 << k
        self.q = j - 1
        ek = k
        while (1 << ek) < self.n * self.q:
            ek += 1
        self.extended_k = ek
        self.extended_omega = omega_for_k(ek)
        self.omega = pow(self.extended_omega, 1 << (ek - k), R_MOD)
        self.omega_inv = pow(self.omega, -1, R_MOD)
        self.extended_omega_inv = pow(self.extended_omega, -1, R_MOD)
        self.g_coset = FR_ZETA
        self.g_coset_inv = FR_ZETA * FR_ZETA % R_MOD
        t = []
        orig = pow(FR_ZETA, self.n, R_MOD)
        step = pow(self.extended_omega, self.n, R_MOD)
        cur = orig
        while True:
            t.append(cur)
            cur = cur * step % R_MOD
            if cur == orig:
                break
        self.t_evaluations = [pow(x - 1, -1, R_MOD) for x in t]

    def lagrange_to_coeff(self, a):
        ninv = pow(self.n, -1, R_MOD)
        return [x * ninv % R_MOD for x in ntt(a, self.omega_inv)]

    def coeff_to_extended(self, a):
        z = [1, self.g_coset, self.g_coset_inv]
        v = [x * z[i % 3] % R_MOD for i, x in enumerate(a)] + [0] * ((1 << self.extended_k) - len(a))
        return ntt(v, self.extended_omega)

    def extended_to_coeff(self, a):
        einv = pow(1 << self.extended_k, -1, R_MOD)
        z = [1, self.g_coset_inv, self.g_coset]
        v = [x * einv % R_MOD for x in ntt(a, self.extended_omega_inv)]
        v = [x * z[i % 3] % R_MOD for i, x in enumerate(v)]
        return v[: self.n * self.q]


# ---- transcript (transcript.rs: Blake2bWrite + Challenge255) ----
class Blake2bTranscript:
    def __init__(self):
        self.h = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
        self.proof = bytearray()

    def common_point(self, pt):
        self.h.update(b"\x01")
        x, y = (0, 0) if pt is None else pt
        self.h.update(x.to_bytes(32, "little") + y.to_bytes(32, "little"))

    def common_scalar(self, s):
        self.h.update(b"\x02" + s.to_bytes(32, "little"))

    def write_point(self, pt):
        self.common_point(pt)
        self.proof += g1_compress(pt)

    def write_scalar(self, s):
        self.common_scalar(s)
        self.proof += s.to_bytes(32, "little")

    def squeeze_challenge(self):
        self.h.update(b"\x00")
        d = self.h.copy().digest()
        return int.from_bytes(d, "little") % R_MOD


# ---- XorShiftRng (rand_xorshift) + Fr::random ----
HALO2_TEST_SEED = bytes([0x59, 0x62, 0xbe, 0x5d, 0x76, 0x3d, 0x31, 0x8d, 0x17, 0xdb, 0x37, 0x32, 0x54, 0x06, 0xbc, 0xe5])


class XorShiftRng:
    def __init__(self, seed=HALO2_TEST_SEED):
        self.x, self.y, self.z, self.w = (int.from_bytes(seed[4 * i:4 * i + 4], "little") for i in range(4))

    def next_u32(self):
        t = (self.x ^ (self.x << 11)) & 0xFFFFFFFF
        self.x, self.y, self.z = self.y, self.z, self.w
        self.w = (self.w ^ (self.w >> 19) ^ (t ^ (t >> 8))) & 0xFFFFFFFF
        return self.w

    def next_u64(self):
        lo = self.next_u32()
        return lo | (self.next_u32() << 32)

    def fr_random(self):
        """Fr::random = from_u512 of 8 x next_u64 (LE)."""
        v = sum(self.next_u64() << (64 * i) for i in range(8))
        return v % R_MOD
