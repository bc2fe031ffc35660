"""Synthetic prove jobs with the shape of the reference's circuits.

`mst_shaped(k)` has the column / gate / lookup / permutation counts SURVEY.md §8 derives for the
Merkle Sum Tree circuit (/root/reference/src/chips/merkle_sum_tree.rs:32-137 + Poseidon
Pow5Chip<5,4> + LtChip<8>): 20 advice columns, 8 single-column lookups into a u8 table, 16
permutation columns (a..e, the instance column, 5 state, 5 rc_b), one degree-6 gate
(selector * (a^5 - b), the Pow5 S-box shape), so d = 6, extended domain 8n, 4 permutation sets,
blinding factors 5.  The witness is trivially satisfiable (no Poseidon needed) and keeps the
dense/sparse mix of the real circuit: "state" columns hold full-width field elements, the
byte columns hold values < 256.  `small(k)` is a reduced variant for fast parity tests and
`v3_shaped(k)` has the shape of Merkle v3 (7 advice, no lookups, 10 permutation columns).

No field arithmetic is done here: small integers are converted to Montgomery form through a
lookup table built with Python integers, dense columns are drawn directly as Montgomery limbs.
"""
import numpy as np

from .circuit import ADVICE, FIXED, INSTANCE, ConstraintSystem, Expr, PermutationAssembly, R_MOD

_MONT_R = (1 << 256) % R_MOD


def mont_from_ints(vals):
    """np.ndarray (n,4) uint64 Montgomery limbs of the given Python/NumPy integers."""
    vals = np.asarray(vals, dtype=object).reshape(-1)
    uniq = {}
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        v = int(v) % R_MOD
        if v not in uniq:
            m = v * _MONT_R % R_MOD
            uniq[v] = [(m >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)]
        out[i] = uniq[v]
    return out


_R_LIMBS = np.array([(R_MOD >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)], dtype=np.uint64)
_CHUNK_LUTS = None


def _add_mod_r(a, b):
    """Vectorised (a + b) mod r on (n,4) uint64 limb arrays with a, b < r (no field multiplication)."""
    out = np.empty_like(a)
    carry = np.zeros(a.shape[0], dtype=np.uint64)
    for j in range(4):
        s = a[:, j] + b[:, j]
        c1 = s < a[:, j]
        s2 = s + carry
        c2 = s2 < s
        out[:, j] = s2
        carry = (c1 | c2).astype(np.uint64)
    # r < 2^254, so a + b < 2^255 never carries out; subtract r where out >= r
    ge = np.ones(a.shape[0], dtype=bool)
    decided = np.zeros(a.shape[0], dtype=bool)
    for j in (3, 2, 1, 0):
        gt, lt = out[:, j] > _R_LIMBS[j], out[:, j] < _R_LIMBS[j]
        ge = np.where(~decided & lt, False, ge)
        decided |= gt | lt
    borrow = np.zeros(a.shape[0], dtype=np.uint64)
    sub = np.empty_like(out)
    for j in range(4):
        d = out[:, j] - _R_LIMBS[j]
        b1 = out[:, j] < _R_LIMBS[j]
        d2 = d - borrow
        b2 = d < borrow
        sub[:, j] = d2
        borrow = (b1 | b2).astype(np.uint64)
    out[ge] = sub[ge]
    return out


def mont_from_u64(arr):
    """Montgomery form of non-negative integers < 2^64, vectorised: the value is split into four
    16-bit chunks, each chunk position has a 65536-entry table of Montgomery forms (Python ints,
    built once), and the four table rows are added mod r with limb arithmetic."""
    global _CHUNK_LUTS
    if _CHUNK_LUTS is None:
        _CHUNK_LUTS = [mont_from_ints([(d << (16 * p)) for d in range(65536)]) for p in range(4)]
    v = np.asarray(arr).astype(np.uint64)
    acc = _CHUNK_LUTS[0][(v & np.uint64(0xFFFF)).astype(np.int64)]
    for p in range(1, 4):
        chunk = ((v >> np.uint64(16 * p)) & np.uint64(0xFFFF)).astype(np.int64)
        if chunk.any():
            acc = _add_mod_r(acc, _CHUNK_LUTS[p][chunk])
    return acc


def mont_from_small(arr, lut_fn=None):
    """arr of non-negative ints < 2^63.  With lut_fn the value is lut_fn(key) (few distinct keys)."""
    arr = np.asarray(arr, dtype=np.int64)
    if lut_fn is None:
        return mont_from_u64(arr)
    keys, inv = np.unique(arr, return_inverse=True)
    lut = mont_from_ints([lut_fn(int(k)) for k in keys])
    return lut[inv]


def random_field(n, rng):
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] %= np.uint64(0x30644e72e131a029)
    return a


class Job:
    """Everything b200zk_pk_create / b200zk_create_proof consume."""

    def __init__(self, cs, k):
        self.cs, self.k, self.n = cs, k, 1 << k
        self.fixed = []          # list of (n,4) Montgomery arrays
        self.advice = []         # list of (n,4) Montgomery arrays (unblinded)
        self.instances = []      # list of lists of Python ints
        self.map_col = self.map_row = None
        self.transcript_repr = 0x0123456789abcdef0123456789abcdef0123456789abcdef0123456789abcdef % R_MOD


def _build(k, seed, n_state, n_bytes, n_rc, with_sum_gate=True):
    """Common builder.  Advice layout: [a, b, c, d, e] + state[n_state] + [partial] + [lt] + bytes[n_bytes]."""
    n = 1 << k
    rng = np.random.Generator(np.random.PCG64(seed))
    A = 5 + n_state + 1 + (1 + n_bytes if n_bytes else 0)
    F = 2 + n_rc + 2                      # q_pow5, u8_table, rc..., q_bool, q_sum
    cs = ConstraintSystem(A, F, 1)
    Q_POW5, U8, RC0, Q_BOOL, Q_SUM = 0, 1, 2, 2 + n_rc, 3 + n_rc
    ST = 5                                # first state column
    # permutation: a..e, instance, state columns, rc_b-like fixed columns (second half of rc)
    for c in range(5):
        cs.enable_equality(ADVICE, c)
    cs.enable_equality(INSTANCE, 0)
    for c in range(n_state):
        cs.enable_equality(ADVICE, ST + c)
    for c in range(n_rc // 2):
        cs.enable_equality(FIXED, RC0 + n_rc // 2 + c)
    # gates
    s0, s1 = cs.query_advice(ST, 0), cs.query_advice(ST + 1, 0)
    qp = cs.query_fixed(Q_POW5, 0)
    cs.create_gate([qp * (s0 * s0 * s0 * s0 * s0 - s1)])                              # degree 6
    if n_rc:
        cs.create_gate([qp * (cs.query_advice(ST + 2, 0) + cs.query_fixed(RC0, 0) - cs.query_advice(ST + 2, 1))])
    c_bit = cs.query_advice(2, 0)
    cs.create_gate([cs.query_fixed(Q_BOOL, 0) * (c_bit * (1 - c_bit))])                # bool
    if with_sum_gate:
        cs.create_gate([cs.query_fixed(Q_SUM, 0) * (cs.query_advice(0, 0) + cs.query_advice(1, 0) - cs.query_advice(3, 1)),
                        cs.query_fixed(Q_SUM, 0) * (cs.query_advice(3, -1) * 0 + cs.query_advice(4, 0) - cs.query_advice(4, 0))])
    BY = 5 + n_state + 2
    for b in range(n_bytes):
        cs.lookup([(cs.query_advice(BY + b, 0), cs.query_fixed(U8, 0))])
    bf = cs.blinding_factors()
    usable = n - (bf + 1)
    job = Job(cs, k)
    # ---- fixed columns
    rows = np.arange(n)
    act = (rows < usable).astype(np.int64)
    act_next = (rows < usable - 1).astype(np.int64)
    fixed = [None] * F
    fixed[Q_POW5] = mont_from_small(act_next)
    T = min(256, usable)                                         # table size (u8 when the domain allows)
    fixed[U8] = mont_from_small(np.where(rows < T, rows, 0))
    rc_small = rng.integers(0, 1 << 30, size=(n_rc, n), dtype=np.int64)
    for c in range(n_rc):
        fixed[RC0 + c] = mont_from_small(rc_small[c] % 251) if c else mont_from_small(rc_small[c])
    fixed[Q_BOOL] = mont_from_small(act)
    fixed[Q_SUM] = mont_from_small(act_next if with_sum_gate else np.zeros(n, dtype=np.int64))
    job.fixed = fixed
    # ---- advice columns
    adv = [None] * A
    a0 = rng.integers(0, 1 << 40, size=n, dtype=np.int64)
    a1 = rng.integers(0, 1 << 40, size=n, dtype=np.int64)
    a3 = np.zeros(n, dtype=np.int64)
    a3[1:] = a0[:-1] + a1[:-1]
    bits = rng.integers(0, 2, size=n, dtype=np.int64)
    adv[0], adv[1], adv[2], adv[3] = mont_from_small(a0), mont_from_small(a1), mont_from_small(bits), mont_from_small(a3)
    adv[4] = mont_from_small(a0)                                 # e == a (copy-constrained below)
    base = rng.integers(0, 256, size=n, dtype=np.int64)
    adv[ST] = mont_from_small(base)
    adv[ST + 1] = mont_from_small(base, lambda v: v ** 5)
    if n_state > 2:
        # state2[r+1] = state2[r] + rc0[r]  (running sum of small constants stays small)
        rc0 = rc_small[0] if n_rc else np.zeros(n, dtype=np.int64)
        cum = np.concatenate([[7], 7 + np.cumsum(rc0[:-1].astype(np.int64))])   # < 2^55 for k <= 24
        adv[ST + 2] = mont_from_u64(cum)
    for c in range(3, n_state):
        adv[ST + c] = random_field(n, rng)                       # dense, Poseidon-state-like
    adv[5 + n_state] = random_field(n, rng)                      # partial_sbox: dense
    if n_bytes:
        adv[5 + n_state + 1] = mont_from_small(bits)             # lt
        for b in range(n_bytes):
            adv[BY + b] = mont_from_small(rng.integers(0, T, size=n, dtype=np.int64))
    job.advice = adv
    # ---- instance + copy constraints
    pub = [int(a0[r]) for r in range(4)]
    job.instances = [pub]
    P = len(cs.permutation)
    asm = PermutationAssembly(P, n)
    pidx = {col: i for i, col in enumerate(cs.permutation)}
    for r in range(4):                                           # instance[0][r] == a[r]
        asm.copy(pidx[(INSTANCE, 0)], r, pidx[(ADVICE, 0)], r)
    rr = np.arange(8, usable - 1)
    asm.copy_pairs(pidx[(ADVICE, 4)], rr, pidx[(ADVICE, 0)], rr)  # e[r] == a[r]
    job.map_col, job.map_row = asm.map_col, asm.map_row
    return job


def generic_shapes(k, seed=4):
    """Regression job for the shapes of the reference's other experiments:
    a degree-17 gate like SafeAccumulator's range_check(v, 16) times a selector
    (/root/reference/src/chips/safe_accumulator.rs:118,158-159 -> extended domain 16n, permutation
    chunks of 15 columns) and a two-expression *dynamic* lookup into an advice table column like
    LessThan v1 (/root/reference/src/chips/less_than.rs:62-88), with products inside the lookup
    expressions and rotations on both sides."""
    n = 1 << k
    rng = np.random.Generator(np.random.PCG64(seed))
    cs = ConstraintSystem(4, 3, 1)
    Q_RANGE, TAB, Q_LK = 0, 1, 2
    cs.enable_equality(ADVICE, 0)
    cs.enable_equality(ADVICE, 1)
    cs.enable_equality(INSTANCE, 0)
    cs.enable_equality(FIXED, TAB)
    a0 = cs.query_advice(0, 0)
    prod = a0
    for i in range(1, 16):
        prod = prod * (a0 - i)
    cs.create_gate([cs.query_fixed(Q_RANGE, 0) * prod])                                  # degree 17
    ql = cs.query_fixed(Q_LK, 0)
    cs.lookup([(ql * cs.query_advice(1, 0), cs.query_advice(3, 0)),                       # dynamic (advice) table
               (ql * cs.query_advice(2, 1), cs.query_fixed(TAB, 0))])
    assert cs.degree() == 17 and cs.blinding_factors() == 5
    bf = cs.blinding_factors()
    usable = n - (bf + 1)
    job = Job(cs, k)
    rows = np.arange(n)
    act_next = (rows < usable - 1).astype(np.int64)
    T = min(64, usable)
    tab_a = np.where(rows < T, rows * 3, 0)              # advice table column: (3t, 7t) pairs, row 0 = (0, 0)
    tab_f = np.where(rows < T, rows * 7, 0)
    job.fixed = [mont_from_small((rows < usable).astype(np.int64)), mont_from_small(tab_f), mont_from_small(act_next)]
    v0 = rng.integers(0, 16, size=n, dtype=np.int64)
    pick = rng.integers(0, T, size=n, dtype=np.int64)
    a1 = pick * 3 * act_next                            # rows with q_lookup = 0 look up (0, 0)
    a2 = np.zeros(n, dtype=np.int64)
    a2[1:] = (pick * 7 * act_next)[:-1]                  # queried at rotation +1
    job.advice = [mont_from_small(v0), mont_from_small(a1), mont_from_small(a2), mont_from_small(tab_a)]
    job.instances = [[int(v0[0]), int(v0[1])]]
    asm = PermutationAssembly(len(cs.permutation), n)
    pidx = {col: i for i, col in enumerate(cs.permutation)}
    for r in range(2):
        asm.copy(pidx[(INSTANCE, 0)], r, pidx[(ADVICE, 0)], r)
    job.map_col, job.map_row = asm.map_col, asm.map_row
    return job


def mst_shaped(k, seed=1):
    """A = 20 advice, 8 lookups, 16 permutation columns, d = 6 (MST row)."""
    return _build(k, seed, n_state=5, n_bytes=8, n_rc=10)


def v3_shaped(k, seed=2):
    """Merkle v3 shape: 7 advice (3 + 3 state + partial), no lookups, 10 permutation columns."""
    return _build(k, seed, n_state=3, n_bytes=0, n_rc=6, with_sum_gate=False)


def small(k, seed=3):
    """Reduced variant for fast tests: 3 state columns, 2 lookups, 2 rc columns."""
    return _build(k, seed, n_state=3, n_bytes=2, n_rc=2)
