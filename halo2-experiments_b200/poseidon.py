"""Poseidon over bn256 Fr as the reference's chips use it (host side, Python integers).

Restates `halo2_gadgets::poseidon::primitives` at the PSE tag the reference pins
(/root/reference/Cargo.toml:11): the Grain LFSR parameter generation (`grain.rs`), the Cauchy MDS
matrix (`mds.rs`), the permutation and the `ConstantLength<L>` sponge (`primitives.rs`), for any
`Spec` (WIDTH, RATE, R_F, R_P).  The reference's spec is `MySpec<F, WIDTH, RATE>`:
R_F = 8, R_P = 56, x^5 S-box, secure_mds = 0 (/root/reference/src/chips/poseidon/spec.rs:16-32).

This is the native hash the reference's tests use to build Merkle roots
(/root/reference/src/circuits/merkle_sum_tree.rs:150-170) and the source of the round constants
the Pow5 chip loads into its fixed columns.  Witness generation only — nothing here is on the
GPU path.  The generator is pinned in tests/test_frontend.py against the published bn254
constants (first round constant, first MDS entry and poseidon([1, 2]) for t = 3, R_P = 57), which
come from the same Grain procedure.
"""
from functools import lru_cache

from .circuit import R_MOD

NUM_BITS = 254


class Grain:
    """grain.rs: 80-bit LFSR, self-shrinking output."""

    def __init__(self, t, r_f, r_p, sbox_tag=0):
        bits = []

        def put(value, length):
            bits.extend((value >> (length - 1 - i)) & 1 for i in range(length))

        put(1, 2)                  # FieldType::PrimeOrder
        put(sbox_tag, 4)           # SboxType::Pow = 0, Inv = 1
        put(NUM_BITS, 12)
        put(t, 12)
        put(r_f, 10)
        put(r_p, 10)
        bits.extend([1] * 30)
        assert len(bits) == 80
        self.state = bits
        for _ in range(160):
            self._new_bit()

    def _new_bit(self):
        s = self.state
        b = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0)
        s.append(b)
        return b

    def next_bit(self):
        while not self._new_bit():
            self._new_bit()
        return self._new_bit()

    def _next_int(self):
        v = 0
        for _ in range(NUM_BITS):
            v = (v << 1) | self.next_bit()         # MSB first, like the reference script
        return v

    def next_field_element(self):
        while True:
            v = self._next_int()
            if v < R_MOD:
                return v

    def next_field_element_without_rejection(self):
        return self._next_int() % R_MOD


def _mat_inverse(m):
    t = len(m)
    a = [list(row) + [1 if i == j else 0 for j in range(t)] for i, row in enumerate(m)]
    for c in range(t):
        p = next(r for r in range(c, t) if a[r][c])
        a[c], a[p] = a[p], a[c]
        inv = pow(a[c][c], -1, R_MOD)
        a[c] = [x * inv % R_MOD for x in a[c]]
        for r in range(t):
            if r != c and a[r][c]:
                f = a[r][c]
                a[r] = [(x - f * y) % R_MOD for x, y in zip(a[r], a[c])]
    return [row[t:] for row in a]


def generate_mds(grain, t, select=0):
    """mds.rs::generate_mds: first `select`-skipped Cauchy matrix 1/(x_i + y_j)."""
    while True:
        while True:
            vals = [grain.next_field_element_without_rejection() for _ in range(2 * t)]
            if len(set(vals)) == len(vals):
                break
        if select:
            select -= 1
            continue
        xs, ys = vals[:t], vals[t:]
        mds = [[pow((xs[i] + ys[j]) % R_MOD, -1, R_MOD) for j in range(t)] for i in range(t)]
        return mds, _mat_inverse(mds)


class Spec:
    """primitives::Spec: constants() = (round_constants, mds, mds_inv)."""

    def __init__(self, width, rate, full_rounds=8, partial_rounds=56, secure_mds=0):
        self.width, self.rate = width, rate
        self.full_rounds, self.partial_rounds, self.secure_mds = full_rounds, partial_rounds, secure_mds
        self._consts = None

    def constants(self):
        if self._consts is None:
            g = Grain(self.width, self.full_rounds, self.partial_rounds)
            rc = [[g.next_field_element() for _ in range(self.width)]
                  for _ in range(self.full_rounds + self.partial_rounds)]
            mds, mds_inv = generate_mds(g, self.width, self.secure_mds)
            self._consts = (rc, mds, mds_inv)
        return self._consts


@lru_cache(maxsize=None)
def my_spec(width, rate):
    """MySpec<Fr, WIDTH, RATE> (/root/reference/src/chips/poseidon/spec.rs)."""
    return Spec(width, rate, 8, 56, 0)


def pow5(v):
    v2 = v * v % R_MOD
    return v2 * v2 % R_MOD * v % R_MOD


def mat_vec(m, v):
    return [sum(a * b for a, b in zip(row, v)) % R_MOD for row in m]


def permute(state, spec):
    """primitives::permute: R_F/2 full, R_P partial, R_F/2 full rounds."""
    rc, mds, _ = spec.constants()
    rf, rp = spec.full_rounds // 2, spec.partial_rounds
    state = list(state)
    r = 0
    for phase, count in (("f", rf), ("p", rp), ("f", rf)):
        for _ in range(count):
            state = [(s + c) % R_MOD for s, c in zip(state, rc[r])]
            if phase == "f":
                state = [pow5(s) for s in state]
            else:
                state[0] = pow5(state[0])
            state = mat_vec(mds, state)
            r += 1
    return state


def hash_constant_length(message, spec):
    """primitives::Hash::<_, S, ConstantLength<L>, WIDTH, RATE>::init().hash(message)."""
    L, rate, width = len(message), spec.rate, spec.width
    state = [0] * width
    state[rate] = (L << 64) % R_MOD                         # ConstantLength::initial_capacity_element
    k = (L + rate - 1) // rate
    padded = [int(m) % R_MOD for m in message] + [0] * (k * rate - L)
    for c in range(k):
        for i in range(rate):
            state[i] = (state[i] + padded[c * rate + i]) % R_MOD
        state = permute(state, spec)
    return state[0]
