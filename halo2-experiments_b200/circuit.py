"""Constraint-system description consumed by the prover (b200zk_pk_create / b200zk_create_proof).

Mirrors the parts of halo2_proofs v2023_02_02 `plonk::ConstraintSystem` that `create_proof`
reads (src/plonk/circuit.rs): column counts, the advice/fixed/instance query lists (their order
fixes the order of evaluations in the proof), gate polynomials, lookup arguments, permutation
columns, `blinding_factors()` and `degree()`.  Selectors are assumed already compressed into
fixed columns, as they are in a `ProvingKey`.  The reference builds this object through the
chips' `configure` functions (e.g. /root/reference/src/chips/merkle_sum_tree.rs:32-137).

Pure Python + numpy, no field arithmetic: expressions carry constants as canonical integers and
the library converts them.
"""
import numpy as np

ADVICE, FIXED, INSTANCE = 0, 1, 2
OP_CONST, OP_FIXED, OP_ADVICE, OP_INSTANCE, OP_NEG, OP_ADD, OP_MUL, OP_SCALE = range(8)
R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001


class Expr:
    """plonk::Expression<Fr> (Constant, Fixed, Advice, Instance, Negated, Sum, Product, Scaled)."""
    __slots__ = ("kind", "a", "b")

    def __init__(self, kind, a=None, b=None):
        self.kind, self.a, self.b = kind, a, b

    @staticmethod
    def const(v):
        return Expr("const", int(v) % R_MOD)

    def __add__(self, o):
        return Expr("sum", self, _lift(o))

    def __radd__(self, o):
        return Expr("sum", _lift(o), self)

    def __sub__(self, o):
        return Expr("sum", self, Expr("neg", _lift(o)))

    def __rsub__(self, o):
        return Expr("sum", _lift(o), Expr("neg", self))

    def __mul__(self, o):
        if isinstance(o, int):
            return Expr("scaled", self, o % R_MOD)
        return Expr("prod", self, o)

    def __rmul__(self, o):
        return self.__mul__(o)

    def __neg__(self):
        return Expr("neg", self)

    def degree(self):
        k = self.kind
        if k == "const":
            return 0
        if k in ("fixed", "advice", "instance"):
            return 1
        if k in ("neg", "scaled"):
            return self.a.degree()
        if k == "sum":
            return max(self.a.degree(), self.b.degree())
        return self.a.degree() + self.b.degree()


def _lift(o):
    return o if isinstance(o, Expr) else Expr.const(o)


class ConstraintSystem:
    def __init__(self, num_advice, num_fixed, num_instance):
        self.num_advice, self.num_fixed, self.num_instance = num_advice, num_fixed, num_instance
        self.advice_queries, self.fixed_queries, self.instance_queries = [], [], []
        self.gates = []            # flat list of gate polynomials (gate.polynomials() concatenated)
        self.lookups = []          # (input_expressions, table_expressions)
        self.permutation = []      # (column_type, index) in enable_equality order
        self.minimum_degree = None

    # -- queries (dedupe like query_*_index) --
    def _query(self, lst, kind, col, rot):
        key = (col, rot)
        if key not in lst:
            lst.append(key)
        return Expr(kind, lst.index(key))

    def query_advice(self, col, rot=0):
        assert 0 <= col < self.num_advice
        return self._query(self.advice_queries, "advice", col, rot)

    def query_fixed(self, col, rot=0):
        assert 0 <= col < self.num_fixed
        return self._query(self.fixed_queries, "fixed", col, rot)

    def query_instance(self, col, rot=0):
        assert 0 <= col < self.num_instance
        return self._query(self.instance_queries, "instance", col, rot)

    def enable_equality(self, col_type, col):
        """meta.enable_equality: query the column at Rotation::cur() and add it to the permutation."""
        {ADVICE: self.query_advice, FIXED: self.query_fixed, INSTANCE: self.query_instance}[col_type](col, 0)
        if (col_type, col) not in self.permutation:
            self.permutation.append((col_type, col))

    def create_gate(self, polys):
        self.gates.extend(polys)

    def lookup(self, pairs):
        """meta.lookup / lookup_any: list of (input_expression, table_expression)."""
        self.lookups.append(([p[0] for p in pairs], [p[1] for p in pairs]))

    # -- derived quantities (circuit.rs) --
    def blinding_factors(self):
        per_col = {}
        for col, _ in self.advice_queries:
            per_col[col] = per_col.get(col, 0) + 1
        factors = max([0] + list(per_col.values()))
        return max(3, factors) + 2

    def degree(self):
        degree = 3                                                   # permutation.required_degree()
        for ins, tabs in self.lookups:
            di = max([1] + [e.degree() for e in ins])
            dt = max([1] + [e.degree() for e in tabs])
            degree = max(degree, max(4, 2 + di + dt))
        degree = max([degree] + [g.degree() for g in self.gates])
        return max(degree, self.minimum_degree or 1)

    def permutation_chunk_len(self):
        return self.degree() - 2

    def num_permutation_sets(self):
        c = self.permutation_chunk_len()
        return (len(self.permutation) + c - 1) // c

    # -- serialisation for the C ABI --
    def to_blob(self, k):
        consts, prog = [], []

        def const_idx(v):
            if v not in consts:
                consts.append(v)
            return consts.index(v)

        def emit(e):
            kd = e.kind
            if kd == "const":
                prog.append(OP_CONST | (const_idx(e.a) << 8))
            elif kd == "fixed":
                prog.append(OP_FIXED | (e.a << 8))
            elif kd == "advice":
                prog.append(OP_ADVICE | (e.a << 8))
            elif kd == "instance":
                prog.append(OP_INSTANCE | (e.a << 8))
            elif kd == "neg":
                emit(e.a); prog.append(OP_NEG)
            elif kd == "scaled":
                emit(e.a); prog.append(OP_SCALE | (const_idx(e.b) << 8))
            elif kd == "sum":
                emit(e.a); emit(e.b); prog.append(OP_ADD)
            elif kd == "prod":
                emit(e.a); emit(e.b); prog.append(OP_MUL)
            else:
                raise ValueError(kd)

        def emit_expr(e):
            off = len(prog)
            emit(e)
            return (off, len(prog) - off)

        gate_tab = [emit_expr(g) for g in self.gates]
        lookup_tab = []
        for ins, tabs in self.lookups:
            lookup_tab.append(([emit_expr(e) for e in ins], [emit_expr(e) for e in tabs]))
        w = [0x324B5A42, 1, k, self.num_advice, self.num_fixed, self.num_instance,
             len(self.advice_queries), len(self.fixed_queries), len(self.instance_queries),
             len(self.gates), len(self.lookups), len(self.permutation), len(consts), len(prog),
             self.blinding_factors(), self.degree()]
        for lst in (self.advice_queries, self.fixed_queries, self.instance_queries):
            for col, rot in lst:
                w += [col, rot & 0xFFFFFFFF]
        for t, c in self.permutation:
            w += [t, c]
        for off, ln in gate_tab:
            w += [off, ln]
        for ins, tabs in lookup_tab:
            w.append(len(ins))
            for off, ln in ins + tabs:
                w += [off, ln]
        for v in consts:
            w += [(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)]
        w += prog
        return np.array(w, dtype=np.uint32)


class PermutationAssembly:
    """permutation::keygen::Assembly: copy constraints -> sigma mapping (src/plonk/permutation/keygen.rs)."""

    def __init__(self, num_columns, n):
        self.m, self.n = num_columns, n
        cols = np.repeat(np.arange(num_columns, dtype=np.uint32), n).reshape(num_columns, n)
        rows = np.tile(np.arange(n, dtype=np.uint32), (num_columns, 1))
        self.map_col, self.map_row = cols.copy(), rows.copy()
        self.aux_col, self.aux_row = cols.copy(), rows.copy()
        self.sizes = np.ones((num_columns, n), dtype=np.uint32)

    def copy(self, lc, lr, rc, rr):
        left = (int(self.aux_col[lc, lr]), int(self.aux_row[lc, lr]))
        right = (int(self.aux_col[rc, rr]), int(self.aux_row[rc, rr]))
        if left == right:
            return
        if self.sizes[left] < self.sizes[right]:
            left, right = right, left
        self.sizes[left] += self.sizes[right]
        i = right
        while True:
            self.aux_col[i], self.aux_row[i] = left
            i = (int(self.map_col[i]), int(self.map_row[i]))
            if i == right:
                break
        tmp = (self.map_col[lc, lr], self.map_row[lc, lr])
        self.map_col[lc, lr], self.map_row[lc, lr] = self.map_col[rc, rr], self.map_row[rc, rr]
        self.map_col[rc, rr], self.map_row[rc, rr] = tmp

    def copy_pairs(self, lc, lrows, rc, rrows):
        """Vectorised copy() for pairs of so-far-untouched cells: each pair becomes a 2-cycle,
        exactly what the sequential algorithm produces for singleton cycles."""
        lrows, rrows = np.asarray(lrows), np.asarray(rrows)
        self.map_col[lc, lrows], self.map_row[lc, lrows] = rc, rrows
        self.map_col[rc, rrows], self.map_row[rc, rrows] = lc, lrows
        self.aux_col[rc, rrows], self.aux_row[rc, rrows] = lc, lrows
        self.sizes[lc, lrows] = 2
