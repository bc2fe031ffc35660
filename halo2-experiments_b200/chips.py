"""The reference's hot-path chips and circuits on the front end of `frontend.py`.

  * `Pow5Chip`             halo2_gadgets::poseidon::pow5 (PSE tag v2023_02_02): gates "full round",
                           "partial rounds", "pad-and-add"; regions "initial state", "add input",
                           "permute state" (1 + R_F + R_P/2 = 37 rows for MySpec);
  * `PoseidonChip`         /root/reference/src/chips/poseidon/hash.rs:46-89;
  * `LtChip`               zkevm-circuits `gadgets::less_than` at rev 37b8aca
                           (/root/reference/Cargo.toml:18): lt + N_BYTES diff columns, u8 table
                           column loaded by `load`, one `lookup_any` per diff byte;
  * `MerkleSumTreeChip`    /root/reference/src/chips/merkle_sum_tree.rs:32-366;
  * `MerkleSumTreeCircuit` /root/reference/src/circuits/merkle_sum_tree.rs:16-110;
  * `MerkleTreeV3Chip`     /root/reference/src/chips/merkle_v3.rs:27-172;
  * `MerkleTreeV3Circuit`  /root/reference/src/circuits/merkle_v3.rs:11-62;
  * `LessThanChip/Circuit` /root/reference/src/chips/less_than.rs:33-88, circuits/less_than.rs:8-41
                           (dynamic lookup into an advice table filled from the instance column);
  * `IsZeroChip`           /root/reference/src/chips/is_zero.rs;
  * `SafeAccumulatorChip/Circuit`  /root/reference/src/chips/safe_accumulator.rs:33-271,
                           circuits/safe_accumulator.rs:7-91 (degree-17 range-check gates).

Column allocation order, query order, gate order, region order and copy-constraint order follow
the Rust sources statement by statement: they determine the query lists, the selector compression,
the permutation polynomials and therefore the proof bytes.
"""
from . import poseidon as P
from .circuit import R_MOD, Expr


def _const(v):
    return Expr.const(v)


# ---------------------------------------------------------------- Pow5Chip


class Pow5Config:
    pass


class Pow5Chip:
    def __init__(self, config):
        self.config = config

    @staticmethod
    def configure(meta, spec, state, partial_sbox, rc_a, rc_b):
        width, rate = spec.width, spec.rate
        assert rate == width - 1 and len(state) == len(rc_a) == len(rc_b) == width
        assert spec.full_rounds % 2 == 0 and spec.partial_rounds % 2 == 0
        round_constants, m_reg, m_inv = spec.constants()
        for column in list(state) + list(rc_b):
            meta.enable_equality(column)
        s_full, s_partial, s_pad_and_add = meta.selector(), meta.selector(), meta.selector()

        def pow_5(v):
            v2 = v * v
            return v2 * v2 * v

        # "full round"
        s = meta.query_selector(s_full)
        polys = []
        for next_idx in range(width):
            state_next = meta.query(state[next_idx], 1)
            expr = None
            for idx in range(width):
                state_cur = meta.query(state[idx], 0)
                rc = meta.query(rc_a[idx], 0)
                term = pow_5(state_cur + rc) * m_reg[next_idx][idx]
                expr = term if expr is None else expr + term
            polys.append(s * (expr - state_next))
        meta.create_gate("full round", polys)

        # "partial rounds"
        cur_0 = meta.query(state[0], 0)
        mid_0 = meta.query(partial_sbox, 0)
        rc_a0 = meta.query(rc_a[0], 0)
        rc_b0 = meta.query(rc_b[0], 0)
        s = meta.query_selector(s_partial)

        def mid(idx):
            acc = mid_0 * m_reg[idx][0]
            for cur_idx in range(1, width):
                cur = meta.query(state[cur_idx], 0)
                rc = meta.query(rc_a[cur_idx], 0)
                acc = acc + (cur + rc) * m_reg[idx][cur_idx]
            return acc

        def nxt(idx):
            acc = None
            for next_idx in range(width):
                term = meta.query(state[next_idx], 1) * m_inv[idx][next_idx]
                acc = term if acc is None else acc + term
            return acc

        def partial_round_linear(idx):
            rc = meta.query(rc_b[idx], 0)
            return mid(idx) + rc - nxt(idx)

        polys = [pow_5(cur_0 + rc_a0) - mid_0, pow_5(mid(0) + rc_b0) - nxt(0)]
        polys += [partial_round_linear(idx) for idx in range(1, width)]
        meta.create_gate("partial rounds", [s * p for p in polys])

        # "pad-and-add"
        initial_state_rate = meta.query(state[rate], -1)
        output_state_rate = meta.query(state[rate], 1)
        s = meta.query_selector(s_pad_and_add)
        polys = []
        for idx in range(rate):
            initial_state = meta.query(state[idx], -1)
            inp = meta.query(state[idx], 0)
            output_state = meta.query(state[idx], 1)
            polys.append(initial_state + inp - output_state)
        polys.append(initial_state_rate - output_state_rate)
        meta.create_gate("pad-and-add", [s * p for p in polys])

        cfg = Pow5Config()
        cfg.spec, cfg.state, cfg.partial_sbox, cfg.rc_a, cfg.rc_b = spec, list(state), partial_sbox, list(rc_a), list(rc_b)
        cfg.s_full, cfg.s_partial, cfg.s_pad_and_add = s_full, s_partial, s_pad_and_add
        cfg.half_full_rounds, cfg.half_partial_rounds = spec.full_rounds // 2, spec.partial_rounds // 2
        cfg.round_constants, cfg.m_reg, cfg.m_inv = round_constants, m_reg, m_inv
        return cfg

    # PoseidonSpongeInstructions::initial_state (domain ConstantLength<L>)
    def initial_state(self, layouter, L):
        cfg = self.config
        rate = cfg.spec.rate

        def assign(region):
            state = [region.assign_advice_from_constant(cfg.state[i], 0, 0) for i in range(rate)]
            state.append(region.assign_advice_from_constant(cfg.state[rate], 0, (L << 64) % R_MOD))
            return state

        return layouter.assign_region("initial state for domain ConstantLength", assign)

    # PoseidonSpongeInstructions::add_input; `inputs` = RATE entries, each a Cell (message word)
    # or ("pad", value)
    def add_input(self, layouter, initial_state, inputs):
        cfg = self.config
        width, rate = cfg.spec.width, cfg.spec.rate

        def assign(region):
            region.enable_selector(cfg.s_pad_and_add, 1)
            init = [initial_state[i].copy_advice(region, cfg.state[i], 0) for i in range(width)]
            words = []
            for i in range(rate):
                w = inputs[i]
                if isinstance(w, tuple):
                    w = region.assign_fixed(cfg.rc_b[i], 1, w[1])
                words.append(w.copy_advice(region, cfg.state[i], 1))
            out = []
            for i in range(width):
                v = (init[i].value + (words[i].value if i < rate else 0)) % R_MOD
                out.append(region.assign_advice(cfg.state[i], 2, v))
            return out

        return layouter.assign_region("add input for domain ConstantLength", assign)

    # PoseidonInstructions::permute
    def permute(self, layouter, initial_state):
        cfg = self.config
        width = cfg.spec.width
        rc, m = cfg.round_constants, cfg.m_reg

        def do_round(region, state, rnd, offset, gate, round_fn):
            region.enable_selector(gate, offset)
            for i in range(width):
                region.assign_fixed(cfg.rc_a[i], offset, rc[rnd][i])
            next_state = round_fn(region, state)
            return [region.assign_advice(cfg.state[i], offset + 1, next_state[i]) for i in range(width)]

        def full_round(region, state, rnd, offset):
            def fn(_region, st):
                r = [P.pow5((w.value + c) % R_MOD) for w, c in zip(st, rc[rnd])]
                return P.mat_vec(m, r)
            return do_round(region, state, rnd, offset, cfg.s_full, fn)

        def partial_round(region, state, rnd, offset):
            def fn(region_, st):
                p = [w.value for w in st]
                r = [P.pow5((p[0] + rc[rnd][0]) % R_MOD)] + [(p[i] + rc[rnd][i]) % R_MOD for i in range(1, width)]
                region_.assign_advice(cfg.partial_sbox, offset, r[0])
                p_mid = P.mat_vec(m, r)
                for i in range(width):
                    region_.assign_fixed(cfg.rc_b[i], offset, rc[rnd + 1][i])
                r_mid = [P.pow5((p_mid[0] + rc[rnd + 1][0]) % R_MOD)] + \
                        [(p_mid[i] + rc[rnd + 1][i]) % R_MOD for i in range(1, width)]
                return P.mat_vec(m, r_mid)
            return do_round(region, state, rnd, offset, cfg.s_partial, fn)

        def assign(region):
            state = [initial_state[i].copy_advice(region, cfg.state[i], 0) for i in range(width)]
            hf, hp = cfg.half_full_rounds, cfg.half_partial_rounds
            for r in range(hf):
                state = full_round(region, state, r, r)
            for r in range(hp):
                state = partial_round(region, state, hf + 2 * r, hf + r)
            for r in range(hf):
                state = full_round(region, state, hf + 2 * hp + r, hf + hp + r)
            return state

        return layouter.assign_region("permute state", assign)

    # gadget: Hash::<_, _, S, ConstantLength<L>, WIDTH, RATE>::init(chip, layouter).hash(layouter, message)
    def hash(self, layouter, message):
        rate = self.config.spec.rate
        L = len(message)
        state = self.initial_state(layouter, L)                       # Hash::init -> Sponge::new
        padded = list(message) + [("pad", 0)] * ((-L) % rate)         # ConstantLength::padding
        buf = []
        for word in padded:                                           # Sponge::absorb
            if len(buf) == rate:
                state = self.permute(layouter, self.add_input(layouter, state, buf))
                buf = []
            buf.append(word)
        buf += [("pad", 0)] * (rate - len(buf))                        # (never needed for ConstantLength)
        state = self.permute(layouter, self.add_input(layouter, state, buf))   # finish_absorbing
        return state[0]                                               # squeeze


class PoseidonChip:
    """/root/reference/src/chips/poseidon/hash.rs."""

    def __init__(self, config):
        self.config = config

    @staticmethod
    def configure(meta, spec, hash_inputs):
        width = spec.width
        partial_sbox = meta.advice_column()
        rc_a = [meta.fixed_column() for _ in range(width)]
        rc_b = [meta.fixed_column() for _ in range(width)]
        for i in range(width):
            meta.enable_equality(hash_inputs[i])
        meta.enable_constant(rc_b[0])
        return Pow5Chip.configure(meta, spec, hash_inputs, partial_sbox, rc_a, rc_b)

    def hash(self, layouter, input_cells):
        return Pow5Chip(self.config).hash(layouter, input_cells)


# ---------------------------------------------------------------- LtChip (gadgets::less_than)


class LtConfig:
    def is_lt(self, meta, rotation=None):
        return meta.query(self.lt, 0 if rotation is None else rotation)


class LtChip:
    def __init__(self, config):
        self.config = config

    @staticmethod
    def configure(meta, n_bytes, q_enable, lhs, rhs):
        cfg = LtConfig()
        cfg.lt = meta.advice_column()
        cfg.diff = [meta.advice_column() for _ in range(n_bytes)]
        cfg.range = 1 << (8 * n_bytes)
        cfg.u8 = meta.fixed_column()
        q = q_enable(meta)
        lt = meta.query(cfg.lt, 0)
        diff_bytes = [meta.query(c, 0) for c in cfg.diff]
        value, mult = _const(0), 1                                     # expr_from_bytes
        for b in diff_bytes:
            value = value + b * mult
            mult = mult * 256 % R_MOD
        check_a = lhs(meta) - rhs(meta) - value + lt * cfg.range
        check_b = lt * (_const(1) - lt)                                # bool_check
        meta.create_gate("lt gate", [q * check_a, q * check_b])
        for cell_column in cfg.diff:
            meta.lookup_any("range check for u8", [(meta.query(cell_column, 0), meta.query(cfg.u8, 0))])
        return cfg

    def assign(self, region, offset, lhs, rhs):
        cfg = self.config
        lt = lhs < rhs
        region.assign_advice(cfg.lt, offset, int(lt))
        diff = (lhs - rhs + (cfg.range if lt else 0)) % R_MOD
        for idx, col in enumerate(cfg.diff):
            region.assign_advice(col, offset, (diff >> (8 * idx)) & 0xFF)

    def load(self, layouter):
        def assign(region):
            for i in range(256):
                region.assign_fixed(self.config.u8, i, i)
        layouter.assign_region("load u8 range check table", assign)


# ---------------------------------------------------------------- Merkle Sum Tree


class MerkleSumTreeConfig:
    pass


class MerkleSumTreeChip:
    WIDTH, RATE, L = 5, 4, 4

    def __init__(self, config):
        self.config = config

    @classmethod
    def configure(cls, meta, advice, instance):
        col_a, col_b, col_c, col_d, col_e = advice
        cfg = MerkleSumTreeConfig()
        cfg.bool_selector, cfg.swap_selector = meta.selector(), meta.selector()
        cfg.sum_selector, cfg.lt_selector = meta.selector(), meta.selector()
        for c in advice:
            meta.enable_equality(c)
        meta.enable_equality(instance)

        s = meta.query_selector(cfg.bool_selector)
        e = meta.query(col_e, 0)
        meta.create_gate("bool constraint", [s * e * (_const(1) - e)])

        s = meta.query_selector(cfg.swap_selector)
        a, b, c, d, e = (meta.query(col, 0) for col in (col_a, col_b, col_c, col_d, col_e))
        l1, l2, r1, r2 = (meta.query(col, 1) for col in (col_a, col_b, col_c, col_d))
        meta.create_gate("swap constraint", [
            s * (e * _const(2) * (c - a) - (l1 - a) - (c - r1)),
            s * (e * _const(2) * (d - b) - (l2 - b) - (d - r2))])

        s = meta.query_selector(cfg.sum_selector)
        left_balance, right_balance, computed_sum = meta.query(col_b, 0), meta.query(col_d, 0), meta.query(col_e, 0)
        meta.create_gate("sum constraint", [s * (left_balance + right_balance - computed_sum)])

        hash_inputs = [meta.advice_column() for _ in range(cls.WIDTH)]
        cfg.poseidon_config = PoseidonChip.configure(meta, P.my_spec(cls.WIDTH, cls.RATE), hash_inputs)
        cfg.lt_config = LtChip.configure(meta, 8,
                                         lambda m: m.query_selector(cfg.lt_selector),
                                         lambda m: m.query(col_a, 0),
                                         lambda m: m.query(col_b, 0))
        cfg.advice, cfg.instance = list(advice), instance

        q_enable = meta.query_selector(cfg.lt_selector)
        check = meta.query(col_c, 0)
        meta.create_gate("verifies that `check` from current config equal to is_lt from LtChip ",
                         [q_enable * (cfg.lt_config.is_lt(meta, None) - check)])
        return cfg

    def assing_leaf_hash_and_balance(self, layouter, leaf_hash, leaf_balance):
        adv = self.config.advice
        h = layouter.assign_region("assign leaf hash", lambda region: region.assign_advice(adv[0], 0, leaf_hash))
        b = layouter.assign_region("assign leaf balance", lambda region: region.assign_advice(adv[1], 0, leaf_balance))
        return h, b

    def merkle_prove_layer(self, layouter, prev_hash, prev_balance, element_hash, element_balance, index):
        cfg = self.config
        adv = cfg.advice

        def assign(region):
            region.enable_selector(cfg.bool_selector, 0)
            region.enable_selector(cfg.swap_selector, 0)
            l1 = prev_hash.copy_advice(region, adv[0], 0)
            l2 = prev_balance.copy_advice(region, adv[1], 0)
            r1 = region.assign_advice(adv[2], 0, element_hash)
            r2 = region.assign_advice(adv[3], 0, element_balance)
            idx = region.assign_advice(adv[4], 0, index)
            vals = (l1.value, l2.value, r1.value, r2.value)
            region.enable_selector(cfg.sum_selector, 1)
            if idx.value % R_MOD != 0:
                vals = (vals[2], vals[3], vals[0], vals[1])
            left_hash = region.assign_advice(adv[0], 1, vals[0])
            left_balance = region.assign_advice(adv[1], 1, vals[1])
            right_hash = region.assign_advice(adv[2], 1, vals[2])
            right_balance = region.assign_advice(adv[3], 1, vals[3])
            computed_sum = region.assign_advice(adv[4], 1, (left_balance.value + right_balance.value) % R_MOD)
            return left_hash, left_balance, right_hash, right_balance, computed_sum

        lh, lb, rh, rb, computed_sum = layouter.assign_region("merkle prove layer", assign)
        computed_hash = PoseidonChip(cfg.poseidon_config).hash(layouter, [lh, lb, rh, rb])
        return computed_hash, computed_sum

    def enforce_less_than(self, layouter, prev_computed_sum_cell, computed_sum, total_assets):
        cfg = self.config
        chip = LtChip(cfg.lt_config)
        chip.load(layouter)

        def assign(region):
            prev_computed_sum_cell.copy_advice(region, cfg.advice[0], 0)
            region.assign_advice_from_instance(cfg.instance, 3, cfg.advice[1], 0)
            region.assign_advice(cfg.advice[2], 0, 1)
            region.enable_selector(cfg.lt_selector, 0)
            chip.assign(region, 0, computed_sum, total_assets)

        layouter.assign_region("enforce sum to be less than total assets", assign)

    def expose_public(self, layouter, cell, row):
        layouter.constrain_instance(cell, self.config.instance, row)


class MerkleSumTreeCircuit:
    """/root/reference/src/circuits/merkle_sum_tree.rs:6-110.  Public inputs:
    [leaf_hash, leaf_balance, root_hash, assets_sum]."""

    def __init__(self, leaf_hash, leaf_balance, path_element_hashes, path_element_balances, path_indices, assets_sum):
        self.leaf_hash, self.leaf_balance = leaf_hash, leaf_balance
        self.path_element_hashes, self.path_element_balances = list(path_element_hashes), list(path_element_balances)
        self.path_indices, self.assets_sum = list(path_indices), assets_sum

    @staticmethod
    def configure(meta):
        advice = [meta.advice_column() for _ in range(5)]
        instance = meta.instance_column()
        return MerkleSumTreeChip.configure(meta, advice, instance)

    def synthesize(self, config, layouter):
        chip = MerkleSumTreeChip(config)
        leaf_hash, leaf_balance = chip.assing_leaf_hash_and_balance(layouter, self.leaf_hash, self.leaf_balance)
        chip.expose_public(layouter, leaf_hash, 0)
        chip.expose_public(layouter, leaf_balance, 1)
        next_hash, next_sum = leaf_hash, leaf_balance
        for i in range(len(self.path_element_balances)):
            next_hash, next_sum = chip.merkle_prove_layer(layouter, next_hash, next_sum, self.path_element_hashes[i],
                                                          self.path_element_balances[i], self.path_indices[i])
        computed_sum = (self.leaf_balance + sum(self.path_element_balances)) % R_MOD
        chip.enforce_less_than(layouter, next_sum, computed_sum, self.assets_sum)
        chip.expose_public(layouter, next_hash, 2)


def compute_merkle_sum_root(leaf, elements, indices):
    """tests::compute_merkle_sum_root (/root/reference/src/circuits/merkle_sum_tree.rs:134-165);
    nodes are (hash, balance) pairs."""
    spec = P.my_spec(5, 4)
    h, bal = leaf
    for (eh, eb), idx in zip(elements, indices):
        msg = [h, bal, eh, eb] if idx == 0 else [eh, eb, h, bal]
        h = P.hash_constant_length(msg, spec)
        bal = (bal + eb) % R_MOD
    return h, bal


# ---------------------------------------------------------------- Merkle tree v3


class MerkleTreeV3Chip:
    WIDTH, RATE, L = 3, 2, 2

    def __init__(self, config):
        self.config = config

    @classmethod
    def configure(cls, meta, advice, instance):
        col_a, col_b, col_c = advice
        cfg = MerkleSumTreeConfig()
        cfg.bool_selector, cfg.swap_selector = meta.selector(), meta.selector()
        for c in advice:
            meta.enable_equality(c)
        meta.enable_equality(instance)
        s = meta.query_selector(cfg.bool_selector)
        c = meta.query(col_c, 0)
        meta.create_gate("bool constraint", [s * c * (_const(1) - c)])
        s = meta.query_selector(cfg.swap_selector)
        a, b, c = meta.query(col_a, 0), meta.query(col_b, 0), meta.query(col_c, 0)
        l, r = meta.query(col_a, 1), meta.query(col_b, 1)
        meta.create_gate("swap constraint", [s * (c * _const(2) * (b - a) - (l - a) - (b - r))])
        hash_inputs = [meta.advice_column() for _ in range(cls.WIDTH)]
        cfg.poseidon_config = PoseidonChip.configure(meta, P.my_spec(cls.WIDTH, cls.RATE), hash_inputs)
        cfg.advice, cfg.instance = list(advice), instance
        return cfg

    def assing_leaf(self, layouter, leaf):
        return layouter.assign_region("assign leaf", lambda region: region.assign_advice(self.config.advice[0], 0, leaf))

    def merkle_prove_layer(self, layouter, node_cell, path_element, index):
        cfg = self.config
        adv = cfg.advice

        def assign(region):
            region.enable_selector(cfg.bool_selector, 0)
            region.enable_selector(cfg.swap_selector, 0)
            node_cell.copy_advice(region, adv[0], 0)
            region.assign_advice(adv[1], 0, path_element)
            region.assign_advice(adv[2], 0, index)
            l, r = node_cell.value, path_element % R_MOD
            if index % R_MOD != 0:
                l, r = r, l
            return region.assign_advice(adv[0], 1, l), region.assign_advice(adv[1], 1, r)

        left, right = layouter.assign_region("merkle prove layer", assign)
        return PoseidonChip(cfg.poseidon_config).hash(layouter, [left, right])

    def expose_public(self, layouter, cell, row):
        layouter.constrain_instance(cell, self.config.instance, row)


class MerkleTreeV3Circuit:
    """/root/reference/src/circuits/merkle_v3.rs:4-62.  Public inputs: [leaf, root]."""

    def __init__(self, leaf, path_elements, path_indices):
        self.leaf, self.path_elements, self.path_indices = leaf, list(path_elements), list(path_indices)

    @staticmethod
    def configure(meta):
        advice = [meta.advice_column() for _ in range(3)]
        instance = meta.instance_column()
        return MerkleTreeV3Chip.configure(meta, advice, instance)

    def synthesize(self, config, layouter):
        chip = MerkleTreeV3Chip(config)
        leaf_cell = chip.assing_leaf(layouter, self.leaf)
        chip.expose_public(layouter, leaf_cell, 0)
        digest = leaf_cell
        for el, idx in zip(self.path_elements, self.path_indices):
            digest = chip.merkle_prove_layer(layouter, digest, el, idx)
        chip.expose_public(layouter, digest, 1)


def compute_merkle_root(leaf, elements, indices):
    """tests::compute_merkle_root (/root/reference/src/circuits/merkle_v3.rs:73-89)."""
    spec = P.my_spec(3, 2)
    digest = leaf % R_MOD
    for el, idx in zip(elements, indices):
        digest = P.hash_constant_length([digest, el] if idx == 0 else [el, digest], spec)
    return digest


# ---------------------------------------------------------------- prove jobs for the BASELINE configs


def merkle_sum_tree_job(k, levels=16, seed=1):
    """BASELINE config 5: a Merkle Sum Tree inclusion proof for a tree of 2^levels leaves, padded
    to 2^k rows.  Sibling hashes and balances (< 2^48, so every partial sum fits LtChip's 8 bytes)
    and the path bits come from a seeded generator; assets_sum = total + 1."""
    import numpy as np
    from .frontend import synthesize_job
    rng = np.random.Generator(np.random.PCG64(seed))
    leaf = (int(rng.integers(1, 1 << 62)), int(rng.integers(1, 1 << 48)))
    elements = [(int(rng.integers(1, 1 << 62)), int(rng.integers(1, 1 << 48))) for _ in range(levels)]
    indices = [int(b) for b in rng.integers(0, 2, size=levels)]
    root = compute_merkle_sum_root(leaf, elements, indices)
    assets_sum = root[1] + 1
    circuit = MerkleSumTreeCircuit(leaf[0], leaf[1], [e[0] for e in elements], [e[1] for e in elements], indices, assets_sum)
    return synthesize_job(circuit, k, [[leaf[0], leaf[1], root[0], assets_sum]])


def merkle_v3_job(k, levels=13, seed=2):
    """BASELINE config 2: Poseidon Merkle tree v3 inclusion proof (k = 14 upstream)."""
    import numpy as np
    from .frontend import synthesize_job
    rng = np.random.Generator(np.random.PCG64(seed))
    leaf = int(rng.integers(1, 1 << 62))
    elements = [int(x) for x in rng.integers(1, 1 << 62, size=levels)]
    indices = [int(b) for b in rng.integers(0, 2, size=levels)]
    root = compute_merkle_root(leaf, elements, indices)
    return synthesize_job(MerkleTreeV3Circuit(leaf, elements, indices), k, [[leaf, root]])


# ---------------------------------------------------------------- LessThan (dynamic lookup)


class LessThanConfig:
    pass


class LessThanChip:
    """/root/reference/src/chips/less_than.rs: `input` must appear in `advice_table`, an advice column
    filled from the instance column (a dynamic lookup, no gates)."""

    def __init__(self, config):
        self.config = config

    @staticmethod
    def configure(meta, input_col, table):
        cfg = LessThanConfig()
        cfg.input, cfg.table = input_col, table
        cfg.advice_table = meta.advice_column()
        meta.enable_equality(table)
        meta.enable_equality(cfg.advice_table)
        meta.lookup_any("dynamic lookup check", [(meta.query(input_col, 0), meta.query(cfg.advice_table, 0))])
        return cfg

    def assign(self, layouter, value):
        cfg = self.config

        def assign(region):
            for i in range(1000):
                region.assign_advice_from_instance(cfg.table, i, cfg.advice_table, i)
            region.assign_advice(cfg.input, 0, value)

        layouter.assign_region("less than assignment", assign)


class LessThanCircuit:
    """/root/reference/src/circuits/less_than.rs:8-41; instance column = 0 .. target-1."""

    def __init__(self, value):
        self.value = value

    @staticmethod
    def configure(meta):
        input_col = meta.advice_column()
        table = meta.instance_column()
        return LessThanChip.configure(meta, input_col, table)

    def synthesize(self, config, layouter):
        LessThanChip(config).assign(layouter, self.value)


# ---------------------------------------------------------------- SafeAccumulator


def range_check(value, rng):
    """chips/utils.rs::range_check: value * (1 - value) * ... * (range-1 - value), degree `range`."""
    acc = value
    for i in range(1, rng):
        acc = acc * (_const(i) - value)
    return acc


class IsZeroConfig:
    def expr(self):
        return self.is_zero_expr


class IsZeroChip:
    """/root/reference/src/chips/is_zero.rs."""

    def __init__(self, config):
        self.config = config

    @staticmethod
    def configure(meta, q_enable, value, value_inv):
        cfg = IsZeroConfig()
        cfg.value_inv = value_inv
        v = value(meta)
        q = q_enable(meta)
        inv = meta.query(value_inv, 0)
        cfg.is_zero_expr = _const(1) - v * inv
        meta.create_gate("is_zero", [q * v * cfg.is_zero_expr])
        return cfg

    def assign(self, region, offset, value):
        value %= R_MOD
        region.assign_advice(self.config.value_inv, offset, pow(value, -1, R_MOD) if value else 0)


class SafeAccumulatorConfig:
    pass


class SafeAccumulatorChip:
    """/root/reference/src/chips/safe_accumulator.rs (MAX_BITS-bit limbs in ACC_COLS columns)."""

    def __init__(self, config):
        self.config = config

    @staticmethod
    def configure(meta, max_bits, update_value, left_most_inv, add_carries, accumulate, selector, instance):
        acc_cols = len(accumulate)
        bool_selector, add_carry_selector, overflow_check_selector = selector
        cfg = SafeAccumulatorConfig()
        cfg.is_zero = IsZeroChip.configure(meta, lambda m: m.query_selector(overflow_check_selector),
                                           lambda m: m.query(accumulate[0], 0), left_most_inv)
        for col in accumulate:
            meta.enable_equality(col)
        for col in add_carries:
            meta.enable_equality(col)
        meta.enable_equality(instance)

        s = meta.query_selector(bool_selector)
        polys = []
        for carries in add_carries:
            a = meta.query(carries, 0)
            polys.append(s * a * (_const(1) - a))
        meta.create_gate("bool constraint", polys)

        s_add = meta.query_selector(add_carry_selector)
        s_over = meta.query_selector(overflow_check_selector)
        value = meta.query(update_value, 0)
        previous_acc = [meta.query(accumulate[i], -1) for i in range(acc_cols)]
        carries_acc = [meta.query(add_carries[i], 0) for i in range(acc_cols)]
        updated_acc = [meta.query(accumulate[i], 0) for i in range(acc_cols)]
        shift = _const(1 << max_bits)
        rng = 1 << max_bits
        polys = [s_add * ((value + previous_acc[-1]) - ((carries_acc[-1] * shift) + updated_acc[-1]))]
        polys.append(s_add * range_check(value, rng))
        polys += [s_add * ((updated_acc[i] + (carries_acc[i] * shift)) - (previous_acc[i] + carries_acc[i + 1]))
                  for i in range(acc_cols - 1)]
        polys.append(s_over * (_const(1) - cfg.is_zero.expr()))
        polys += [s_over * range_check(w, rng) for w in previous_acc]
        polys += [s_over * range_check(w, rng) for w in updated_acc]
        meta.create_gate("accumulation constraint", polys)

        cfg.max_bits, cfg.update_value, cfg.left_most_inv = max_bits, update_value, left_most_inv
        cfg.add_carries, cfg.accumulate, cfg.instance = list(add_carries), list(accumulate), instance
        cfg.selector = [add_carry_selector, overflow_check_selector]
        return cfg

    def assign(self, layouter, offset, update_value, accumulated_values):
        cfg = self.config
        acc_cols, max_bits = len(cfg.accumulate), cfg.max_bits
        is_zero_chip = IsZeroChip(cfg.is_zero)

        def assign(region):
            region.enable_selector(cfg.selector[0], offset + 1)
            region.enable_selector(cfg.selector[1], offset + 1)
            total = update_value % R_MOD
            region.assign_advice(cfg.update_value, 1, update_value)
            for idx, val in enumerate(accumulated_values):
                region.assign_advice(cfg.accumulate[idx], 0, val)
            for idx in reversed(range(acc_cols)):
                shift_bits = max_bits * ((acc_cols - 1) - idx)
                total += (accumulated_values[idx] % R_MOD) << shift_bits
                carry = 1 if (total >= (1 << (max_bits + shift_bits)) and idx > 0) else 0
                region.assign_advice(cfg.add_carries[idx], offset + 1, carry)
            limbs = [(total >> (max_bits * i)) & ((1 << max_bits) - 1) for i in range(acc_cols)]   # decompose_bigInt_to_ubits
            cells, updated = [], [0] * acc_cols
            left_most_idx = acc_cols - 1
            for i, v in enumerate(limbs):
                if i == left_most_idx:
                    is_zero_chip.assign(region, 1, v)
                cells.append(region.assign_advice(cfg.accumulate[left_most_idx - i], offset + 1, v))
                updated[left_most_idx - i] = v
            return cells, updated

        return layouter.assign_region("calculate accumulates", assign)

    def expose_public(self, layouter, cell, row):
        layouter.constrain_instance(cell, self.config.instance, row)


class SafeAccumulatorCircuit:
    """/root/reference/src/circuits/safe_accumulator.rs:7-91: 4-bit limbs in 4 columns; public
    inputs = the updated accumulator limbs."""

    def __init__(self, values, accumulated_value):
        self.values, self.accumulated_value = list(values), list(accumulated_value)

    @staticmethod
    def configure(meta):
        new_value = meta.advice_column()
        left_most_acc_inv = meta.advice_column()
        carry_cols = [meta.advice_column() for _ in range(4)]
        acc_cols = [meta.advice_column() for _ in range(4)]
        add_selector, overflow_selector, boolean_selector = meta.selector(), meta.selector(), meta.selector()
        instance = meta.instance_column()
        return SafeAccumulatorChip.configure(meta, 4, new_value, left_most_acc_inv, carry_cols, acc_cols,
                                             [boolean_selector, add_selector, overflow_selector], instance)

    def synthesize(self, config, layouter):
        chip = SafeAccumulatorChip(config)
        cells, previous = chip.assign(layouter, 0, self.values[0], self.accumulated_value)
        for i, v in enumerate(self.values[1:]):
            cells, previous = chip.assign(layouter, i, v, previous)
        for i, cell in enumerate(reversed(cells)):
            chip.expose_public(layouter, cell, i)
