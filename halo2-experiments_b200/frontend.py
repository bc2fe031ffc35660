"""Circuit front end: what the reference's chips need from `halo2_proofs::{plonk, circuit}`.

The reference's circuits are written against halo2's `ConstraintSystem` (configure) and
`Layouter`/`Region` (synthesize) with `SimpleFloorPlanner`
(/root/reference/src/circuits/merkle_sum_tree.rs:16-110).  This module restates the parts of
halo2_proofs v2023_02_02 those calls reach, so that the same circuits can be described here and
turned into a prove job (constraint system, fixed columns, permutation mapping, advice columns)
for `b200zk_pk_create` / `b200zk_create_proof`:

  * `Meta`            plonk/circuit.rs `ConstraintSystem`: columns, selectors, `enable_equality`,
                      `enable_constant`, `create_gate`, `lookup_any`, and `compress_selectors`
                      (plonk/circuit/compress_selectors.rs) run at key generation;
  * `SimpleLayouter`  circuit/floor_planner/single_pass.rs: each region is measured, placed at the
                      first row free in all of its columns, then assigned; constants go to the
                      first constants column right after their region;
  * `Region`, `Cell`  circuit.rs `Region` / `AssignedCell` (`assign_advice`, `assign_fixed`,
                      `assign_advice_from_constant`, `assign_advice_from_instance`, `copy_advice`,
                      `constrain_equal`, `Selector::enable`);
  * `synthesize_job`  keygen_vk/keygen_pk + the prover's witness collection in one pass
                      (plonk/keygen.rs, plonk/prover.rs): the three upstream passes are
                      deterministic and assign the same cells;
  * `mock_verify`     dev.rs `MockProver::verify`: gates on every usable row, lookups, copy
                      constraints and instance cells — what the reference's negative tests assert
                      (/root/reference/src/circuits/merkle_sum_tree.rs:229-343).

Host-side Python integers only; nothing here runs on the GPU or imports test infrastructure.
"""
import numpy as np

from .circuit import ADVICE, FIXED, INSTANCE, ConstraintSystem, Expr, PermutationAssembly, R_MOD
from .circuits_synth import Job, mont_from_ints


class Column(tuple):
    """(column_type, index)."""
    __slots__ = ()

    def __new__(cls, ctype, index):
        return tuple.__new__(cls, (ctype, index))

    ctype = property(lambda s: s[0])
    index = property(lambda s: s[1])


class Selector:
    def __init__(self, index, simple=True):
        self.index, self.simple = index, simple


class SynthesisError(Exception):
    """plonk::Error as raised by keygen / synthesis (NotEnoughRowsAvailable, ...)."""


# ---------------------------------------------------------------- expressions with selectors


def _sel_expr(sel):
    return Expr("selector", sel)


def _expr_degree(e):
    if e.kind == "selector":
        return 1
    k = e.kind
    if k == "const":
        return 0
    if k in ("fixed", "advice", "instance"):
        return 1
    if k in ("neg", "scaled"):
        return _expr_degree(e.a)
    if k == "sum":
        return max(_expr_degree(e.a), _expr_degree(e.b))
    return _expr_degree(e.a) + _expr_degree(e.b)


Expr.degree = _expr_degree      # one definition for both the plain and the selector-carrying trees


def extract_simple_selector(e):
    """Expression::extract_simple_selector: the one simple selector of the expression, or None."""
    k = e.kind
    if k == "selector":
        return e.a if e.a.simple else None
    if k in ("const", "fixed", "advice", "instance"):
        return None
    if k in ("neg", "scaled"):
        return extract_simple_selector(e.a)
    a, b = extract_simple_selector(e.a), extract_simple_selector(e.b)
    if a is not None and b is not None:
        raise AssertionError("two simple selectors cannot be in the same expression")
    return a if a is not None else b


def _substitute(e, repl, must_be_nonsimple):
    k = e.kind
    if k == "selector":
        if must_be_nonsimple:
            assert not e.a.simple, "simple selectors are not allowed in lookup arguments"
        return repl[e.a.index]
    if k in ("const", "fixed", "advice", "instance"):
        return e
    if k == "neg":
        return Expr("neg", _substitute(e.a, repl, must_be_nonsimple))
    if k == "scaled":
        return Expr("scaled", _substitute(e.a, repl, must_be_nonsimple), e.b)
    return Expr(k, _substitute(e.a, repl, must_be_nonsimple), _substitute(e.b, repl, must_be_nonsimple))


# ---------------------------------------------------------------- ConstraintSystem (configure side)


class Meta(ConstraintSystem):
    """plonk::ConstraintSystem<Fr> as `Circuit::configure` sees it."""

    def __init__(self):
        super().__init__(0, 0, 0)
        self.num_selectors = 0
        self.selectors = []
        self.constants = []          # fixed columns usable for constants (enable_constant order)
        self.gate_names = []

    def advice_column(self):
        self.num_advice += 1
        return Column(ADVICE, self.num_advice - 1)

    def fixed_column(self):
        self.num_fixed += 1
        return Column(FIXED, self.num_fixed - 1)

    def instance_column(self):
        self.num_instance += 1
        return Column(INSTANCE, self.num_instance - 1)

    def selector(self):
        s = Selector(self.num_selectors, True)
        self.num_selectors += 1
        self.selectors.append(s)
        return s

    def complex_selector(self):
        s = self.selector()
        s.simple = False
        return s

    def enable_equality(self, column, _idx=None):
        if _idx is not None:                                   # ConstraintSystem-style (type, index) call
            column = Column(column, _idx)
        super().enable_equality(column.ctype, column.index)

    def enable_constant(self, column):
        assert column.ctype == FIXED
        if column not in self.constants:
            self.constants.append(column)
            self.enable_equality(column)

    def query_selector(self, sel):
        return _sel_expr(sel)

    def query(self, column, rot=0):
        if column.ctype == ADVICE:
            return self.query_advice(column.index, rot)
        if column.ctype == FIXED:
            return self.query_fixed(column.index, rot)
        return self.query_instance(column.index, rot)

    def create_gate(self, name, polys=None):
        if polys is None:                                      # ConstraintSystem-style call
            name, polys = "", name
        polys = list(polys)
        assert polys, "Gates must contain at least one constraint."
        self.gate_names.extend([name] * len(polys))
        self.gates.extend(polys)

    def lookup_any(self, name, pairs):
        self.lookups.append(([p[0] for p in pairs], [p[1] for p in pairs]))

    def minimum_rows(self):
        return self.blinding_factors() + 1 + 1 + 1             # bf + l_last + one usable row + ...

    # -- compress_selectors (plonk/circuit/compress_selectors.rs + ConstraintSystem::compress_selectors)
    def compress_selectors(self, activations):
        """activations[s] = set of rows where selector s is enabled.  Returns the list of new fixed
        columns as {row: value} dicts, in allocation order; substitutes the selectors in the gates."""
        assert len(activations) == self.num_selectors
        degrees = [0] * self.num_selectors
        for poly in self.gates:
            s = extract_simple_selector(poly)
            if s is not None:
                degrees[s.index] = max(degrees[s.index], _expr_degree(poly))
        max_degree = self.degree()
        new_columns, repl = [], {}

        def allocate_fixed_column():
            col = self.fixed_column()
            return self.query_fixed(col.index, 0)

        sels = [(i, activations[i], degrees[i]) for i in range(self.num_selectors)]
        # complex selectors, and selectors that do not appear in any gate, get a column each
        rest = []
        for i, act, deg in sels:
            if deg == 0:
                repl[i] = allocate_fixed_column()
                new_columns.append({r: 1 for r in act})
            else:
                rest.append((i, act, deg))
        added = [False] * len(rest)
        for a, (i, act, deg) in enumerate(rest):
            if added[a]:
                continue
            added[a] = True
            assert deg <= max_degree
            d = deg - 1
            combination, combination_added = [(i, act)], [a]
            for b in range(a + 1, len(rest)):
                if d + len(combination) == max_degree:
                    break
                if added[b]:
                    continue
                j, act_j, deg_j = rest[b]
                if any(not act_j.isdisjoint(rest[c][1]) for c in combination_added):
                    continue
                new_d = max(d, deg_j - 1)
                if new_d + len(combination) + 1 > max_degree:
                    continue
                d = new_d
                combination.append((j, act_j))
                combination_added.append(b)
                added[b] = True
            query = allocate_fixed_column()
            assignment = {}
            for root0, (j, act_j) in enumerate(combination):
                assigned_root = root0 + 1
                e = query
                for root in range(1, len(combination) + 1):
                    if root != assigned_root:
                        e = Expr("prod", e, Expr("sum", Expr.const(root), Expr("neg", query)))
                repl[j] = e
                for r in act_j:
                    assignment[r] = assigned_root
            new_columns.append(assignment)
        self.gates = [_substitute(g, repl, False) for g in self.gates]
        self.lookups = [([_substitute(e, repl, True) for e in ins], [_substitute(e, repl, True) for e in tabs])
                        for ins, tabs in self.lookups]
        return new_columns


# ---------------------------------------------------------------- Layouter / Region (synthesize side)


class Cell:
    """circuit::AssignedCell: where the value lives plus the value itself."""
    __slots__ = ("region_index", "row_offset", "column", "value")

    def __init__(self, region_index, row_offset, column, value):
        self.region_index, self.row_offset, self.column, self.value = region_index, row_offset, column, int(value) % R_MOD

    def copy_advice(self, region, column, offset):
        """AssignedCell::copy_advice: assign the same value, then constrain_equal(new, self)."""
        new = region.assign_advice(column, offset, self.value)
        region.constrain_equal(new, self)
        return new


class _Shape:
    """RegionShape: first pass, records the columns a region touches and its height."""

    def __init__(self, layouter, index):
        self.layouter, self.index = layouter, index
        self.columns, self.row_count = [], 0

    def _touch(self, key, offset):
        if key not in self.columns:
            self.columns.append(key)
        self.row_count = max(self.row_count, offset + 1)

    def enable_selector(self, sel, offset):
        self._touch(("sel", sel.index), offset)

    def assign_advice(self, column, offset, value):
        self._touch(column, offset)
        return Cell(self.index, offset, column, value)

    def assign_advice_from_constant(self, column, offset, constant):
        return self.assign_advice(column, offset, constant)

    def assign_advice_from_instance(self, instance, row, advice, offset):
        self._touch(advice, offset)
        return Cell(self.index, offset, advice, self.layouter.instance_value(instance, row))

    def assign_fixed(self, column, offset, value):
        self._touch(column, offset)
        return Cell(self.index, offset, column, value)

    def constrain_equal(self, left, right):
        pass


class _Assign:
    """SingleChipLayouterRegion: second pass, writes through to the assembly."""

    def __init__(self, layouter, index):
        self.layouter, self.index = layouter, index
        self.constants = []

    def _row(self, cell_region, offset):
        return self.layouter.regions[cell_region] + offset

    def enable_selector(self, sel, offset):
        self.layouter.asm.enable_selector(sel, self._row(self.index, offset))

    def assign_advice(self, column, offset, value):
        self.layouter.asm.assign_advice(column, self._row(self.index, offset), value)
        return Cell(self.index, offset, column, value)

    def assign_advice_from_constant(self, column, offset, constant):
        cell = self.assign_advice(column, offset, constant)
        self.constants.append((constant, cell))
        return cell

    def assign_advice_from_instance(self, instance, row, advice, offset):
        value = self.layouter.instance_value(instance, row)
        cell = self.assign_advice(advice, offset, value)
        self.layouter.asm.copy(cell.column, self._row(self.index, offset), instance, row)
        return cell

    def assign_fixed(self, column, offset, value):
        self.layouter.asm.assign_fixed(column, self._row(self.index, offset), value)
        return Cell(self.index, offset, column, value)

    def constrain_equal(self, left, right):
        self.layouter.asm.copy(left.column, self._row(left.region_index, left.row_offset),
                               right.column, self._row(right.region_index, right.row_offset))


class SimpleLayouter:
    """floor_planner::single_pass::SingleChipLayouter."""

    def __init__(self, asm, constants):
        self.asm, self.constant_columns = asm, constants
        self.regions = []            # start row of each region
        self.columns = {}            # column / selector -> next free row

    def instance_value(self, instance, row):
        return self.asm.query_instance(instance, row)

    def namespace(self, _name=None):
        return self

    def assign_region(self, name, assignment):
        index = len(self.regions)
        shape = _Shape(self, index)
        assignment(shape)
        start = max([0] + [self.columns.get(c, 0) for c in shape.columns])
        self.regions.append(start)
        for c in shape.columns:
            self.columns[c] = start + shape.row_count
        region = _Assign(self, index)
        result = assignment(region)
        if region.constants:
            if not self.constant_columns:
                raise SynthesisError("NotEnoughColumnsForConstants")
            ccol = self.constant_columns[0]
            nxt = self.columns.get(ccol, 0)
            for constant, cell in region.constants:
                self.asm.assign_fixed(ccol, nxt, constant)
                self.asm.copy(ccol, nxt, cell.column, self.regions[cell.region_index] + cell.row_offset)
                nxt += 1
            self.columns[ccol] = nxt
        return result

    def constrain_instance(self, cell, instance, row):
        self.asm.copy(cell.column, self.regions[cell.region_index] + cell.row_offset, instance, row)


class Assembly:
    """keygen::Assembly + prover::WitnessCollection in one: fixed cells, selector activations, copy
    constraints and advice cells, with the usable-row check both apply."""

    def __init__(self, cs, k, instances):
        self.cs, self.k, self.n = cs, k, 1 << k
        if self.n < cs.minimum_rows():
            raise SynthesisError(f"NotEnoughRowsAvailable(current_k={k})")
        self.usable_rows = self.n - (cs.blinding_factors() + 1)
        self.fixed = [dict() for _ in range(cs.num_fixed)]
        self.advice = [dict() for _ in range(cs.num_advice)]
        self.selectors = [set() for _ in range(cs.num_selectors)]
        self.instances = [[int(v) % R_MOD for v in col] for col in instances]
        for col in self.instances:
            if len(col) > self.usable_rows:
                raise SynthesisError("InstanceTooLarge")
        self.perm = PermutationAssembly(len(cs.permutation), self.n)
        self.pidx = {col: i for i, col in enumerate(cs.permutation)}
        self.copies = []

    def _check(self, row):
        if not 0 <= row < self.usable_rows:
            raise SynthesisError(f"NotEnoughRowsAvailable(current_k={self.k})")

    def enable_selector(self, sel, row):
        self._check(row)
        self.selectors[sel.index].add(row)

    def query_instance(self, column, row):
        self._check(row)
        col = self.instances[column.index]
        return col[row] if row < len(col) else 0

    def assign_advice(self, column, row, value):
        self._check(row)
        self.advice[column.index][row] = int(value) % R_MOD

    def assign_fixed(self, column, row, value):
        self._check(row)
        self.fixed[column.index][row] = int(value) % R_MOD

    def copy(self, lcol, lrow, rcol, rrow):
        self._check(lrow)
        self._check(rrow)
        for c in (lcol, rcol):
            if tuple(c) not in self.pidx:
                raise SynthesisError(f"ColumnNotInPermutation({tuple(c)})")
        self.copies.append((lcol, lrow, rcol, rrow))
        self.perm.copy(self.pidx[tuple(lcol)], lrow, self.pidx[tuple(rcol)], rrow)


def _column_array(n, cells):
    out = np.zeros((n, 4), dtype=np.uint64)
    if cells:
        rows = np.fromiter(cells.keys(), dtype=np.int64, count=len(cells))
        out[rows] = mont_from_ints(list(cells.values()))
    return out


class FrontendJob(Job):
    """Job plus what mock_verify needs (sparse integer views of the columns, the copy list)."""


def synthesize_job(circuit, k, instances):
    """keygen_vk + keygen_pk + witness synthesis for one circuit instance.

    `circuit` has `configure(meta) -> config` and `synthesize(config, layouter)`;
    `instances` is a list of instance columns (lists of integers)."""
    meta = Meta()
    config = circuit.configure(meta)
    asm = Assembly(meta, k, instances)
    layouter = SimpleLayouter(asm, meta.constants)
    circuit.synthesize(config, layouter)
    n_fixed_before = meta.num_fixed
    selector_columns = meta.compress_selectors(asm.selectors)
    fixed_cells = asm.fixed[:n_fixed_before] + selector_columns
    assert len(fixed_cells) == meta.num_fixed
    job = FrontendJob(meta, k)
    n = 1 << k
    job.fixed = [_column_array(n, c) for c in fixed_cells]
    job.advice = [_column_array(n, c) for c in asm.advice]
    job.instances = [list(col) for col in asm.instances]
    job.map_col, job.map_row = asm.perm.map_col, asm.perm.map_row
    job.fixed_cells, job.advice_cells, job.copies = fixed_cells, asm.advice, asm.copies
    job.usable_rows = asm.usable_rows
    job.rows_used = max([0] + list(layouter.columns.values()))
    return job


# ---------------------------------------------------------------- MockProver::verify


def _eval(e, row, job, n):
    k = e.kind
    if k == "const":
        return e.a
    if k in ("fixed", "advice", "instance"):
        cs = job.cs
        if k == "fixed":
            col, rot = cs.fixed_queries[e.a]
            return job.fixed_cells[col].get((row + rot) % n, 0)
        if k == "advice":
            col, rot = cs.advice_queries[e.a]
            return job.advice_cells[col].get((row + rot) % n, 0)
        col, rot = cs.instance_queries[e.a]
        r = (row + rot) % n
        inst = job.instances[col]
        return inst[r] if r < len(inst) else 0
    if k == "neg":
        return -_eval(e.a, row, job, n) % R_MOD
    if k == "scaled":
        return _eval(e.a, row, job, n) * e.b % R_MOD
    if k == "sum":
        return (_eval(e.a, row, job, n) + _eval(e.b, row, job, n)) % R_MOD
    a = _eval(e.a, row, job, n)
    if a == 0:
        return 0
    return a * _eval(e.b, row, job, n) % R_MOD


def mock_verify(job, rows=None):
    """Returns the list of failures (empty = satisfied), like MockProver::verify."""
    cs, n = job.cs, job.n
    failures = []
    rows = range(job.usable_rows) if rows is None else rows
    for gi, poly in enumerate(cs.gates):
        name = cs.gate_names[gi] if gi < len(getattr(cs, "gate_names", [])) else str(gi)
        for row in rows:
            if _eval(poly, row, job, n) != 0:
                failures.append(("ConstraintNotSatisfied", name, gi, row))
    for li, (ins, tabs) in enumerate(cs.lookups):
        table = {tuple(_eval(e, row, job, n) for e in tabs) for row in range(job.usable_rows)}
        for row in rows:
            if tuple(_eval(e, row, job, n) for e in ins) not in table:
                failures.append(("Lookup", li, row))

    def cell_value(col, row):
        ct, ci = col
        if ct == ADVICE:
            return job.advice_cells[ci].get(row, 0)
        if ct == FIXED:
            return job.fixed_cells[ci].get(row, 0)
        inst = job.instances[ci]
        return inst[row] if row < len(inst) else 0

    for lcol, lrow, rcol, rrow in job.copies:
        if cell_value(lcol, lrow) != cell_value(rcol, rrow):
            failures.append(("Permutation", tuple(lcol), lrow, tuple(rcol), rrow))
    return failures
