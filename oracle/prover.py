"""ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED against the Rust reference.

CPU restatement of the prover path of halo2_proofs (PSE fork, tag v2023_02_02), for one circuit
instance, KZG + SHPLONK + Blake2b transcript — the exact configuration `full_prover` selects at
/root/reference/src/circuits/utils.rs:38-49:

    plonk/keygen.rs            keygen_pk (fixed/sigma polys and cosets, l0, l_last, l_active_row)
    plonk/prover.rs            create_proof orchestration, blinding rows and RNG draw order
    plonk/lookup/prover.rs     commit_permuted / permute_expression_pair / commit_product / evaluate / open
    plonk/permutation/prover.rs commit / evaluate / open
    plonk/vanishing/prover.rs  commit / construct / evaluate / open
    plonk/evaluation.rs        evaluate_h (restated term by term; the result is a unique polynomial)
    poly/kzg/multiopen/shplonk{.rs,/prover.rs}  construct_intermediate_sets, ProverSHPLONK::create_proof
    transcript.rs              Blake2bWrite / Challenge255 (oracle/pyref.py)

The halo2 source is not available in this environment (no Rust toolchain, no crates, no network);
this is written from the published algorithm as recorded in SURVEY.md §3.2, §8(a), Appendix A.
It is pinned by (i) the primitive KATs in tests/test_oracle_kat.py and (ii) `verify_proof` below,
an independent restatement of plonk/verifier.rs + shplonk/verifier.rs that checks every proof this
prover emits (the pairing is replaced by the same identity checked in G1 with the known SRS secret).

Vectors are numpy uint64 (n, 4) Montgomery limbs driven through oracle/binding.py.
"""
import numpy as np

from . import binding as B
from . import pyref as P

R = P.R_MOD


def M(x):
    return B.ints_to_mont([x % R])[0]


def I(a):
    return B.mont_to_ints(np.asarray(a).reshape(1, 4))[0]


def vadd(a, b): return B.binop("add", a, b)
def vsub(a, b): return B.binop("sub", a, b)
def vmul(a, b): return B.binop("mul", a, b)
def vscale(a, x): return B.binop_scalar("mul", a, M(x))
def vadds(a, x): return B.binop_scalar("add", a, M(x))


def const_vec(n, x):
    return np.tile(M(x), (n, 1))


class Rng:
    """Sequential consumer of pre-drawn Fr::random inputs (512-bit wide values)."""

    def __init__(self, wide):
        self.wide, self.pos = np.ascontiguousarray(wide, dtype=np.uint64).reshape(-1, 8), 0

    def take(self, count):
        assert self.pos + count <= self.wide.shape[0], "rng stream exhausted"
        out = B.from_u512(self.wide[self.pos:self.pos + count])
        self.pos += count
        return out

    def one(self):
        return self.take(1)[0]


def rng_draws_needed(cs, k):
    """Number of Fr::random calls create_proof makes for one circuit."""
    n, bf = 1 << k, cs.blinding_factors()
    A, L, S = cs.num_advice, len(cs.lookups), cs.num_permutation_sets()
    q = cs.degree() - 1
    return A * (bf + 1) + A + L * (2 * (bf + 1) + 2) + S * (bf + 1) + L * (bf + 1) + n + 1 + q


# ------------------------------------------------------------------ expressions
def eval_expr(e, fixed, advice, instance, cs, rot_scale, size):
    """plonk::evaluation::evaluate / Expression::evaluate over whole columns."""
    def col(src, q):
        c, rot = q
        return np.roll(src[c], -rot * rot_scale, axis=0)
    k = e.kind
    if k == "const":
        return const_vec(size, e.a)
    if k == "fixed":
        return col(fixed, cs.fixed_queries[e.a])
    if k == "advice":
        return col(advice, cs.advice_queries[e.a])
    if k == "instance":
        return col(instance, cs.instance_queries[e.a])
    if k == "neg":
        x = eval_expr(e.a, fixed, advice, instance, cs, rot_scale, size)
        return vsub(np.zeros_like(x), x)
    if k == "scaled":
        return vscale(eval_expr(e.a, fixed, advice, instance, cs, rot_scale, size), e.b)
    x = eval_expr(e.a, fixed, advice, instance, cs, rot_scale, size)
    y = eval_expr(e.b, fixed, advice, instance, cs, rot_scale, size)
    return vadd(x, y) if k == "sum" else vmul(x, y)


def eval_expr_at(e, fixed_evals, advice_evals, instance_evals):
    """Expression::evaluate on scalar evaluations (verifier side); Python ints."""
    k = e.kind
    if k == "const":
        return e.a
    if k == "fixed":
        return fixed_evals[e.a]
    if k == "advice":
        return advice_evals[e.a]
    if k == "instance":
        return instance_evals[e.a]
    if k == "neg":
        return -eval_expr_at(e.a, fixed_evals, advice_evals, instance_evals) % R
    if k == "scaled":
        return eval_expr_at(e.a, fixed_evals, advice_evals, instance_evals) * e.b % R
    x = eval_expr_at(e.a, fixed_evals, advice_evals, instance_evals)
    y = eval_expr_at(e.b, fixed_evals, advice_evals, instance_evals)
    return (x + y) % R if k == "sum" else x * y % R


# ------------------------------------------------------------------ keygen
class ProvingKey:
    pass


def sigma_from_mapping(dom, map_col, map_row):
    """permutation::keygen::Assembly::build_pk: sigma_i[j] = delta^col * omega^row of the mapped cell."""
    n = dom.n
    d = P.FR_DELTA
    om = B.powers(dom.omega, n)
    out = []
    for c in range(map_col.shape[0]):
        col = om[map_row[c]]
        dl = B.ints_to_mont([pow(d, int(x), R) for x in range(int(map_col.max()) + 1)])
        out.append(vmul(col, dl[map_col[c]]))
    return out


def keygen_pk(cs, k, fixed, map_col, map_row):
    """keygen_pk restated (the vk commitments are not needed by the prover)."""
    pk = ProvingKey()
    pk.cs, pk.k, pk.n = cs, k, 1 << k
    pk.dom = dom = B.Domain(cs.degree(), k)
    n, bf = pk.n, cs.blinding_factors()
    pk.fixed_values = [np.ascontiguousarray(f, dtype=np.uint64).reshape(n, 4) for f in fixed]
    pk.fixed_polys = [dom.lagrange_to_coeff(f) for f in pk.fixed_values]
    pk.fixed_cosets = [dom.coeff_to_extended(f) for f in pk.fixed_polys]
    pk.perm_values = sigma_from_mapping(dom, map_col, map_row) if len(cs.permutation) else []
    pk.perm_polys = [dom.lagrange_to_coeff(s) for s in pk.perm_values]
    pk.perm_cosets = [dom.coeff_to_extended(s) for s in pk.perm_polys]
    one = M(1)
    l0 = np.zeros((n, 4), dtype=np.uint64); l0[0] = one
    l_blind = np.zeros((n, 4), dtype=np.uint64); l_blind[n - bf:] = one
    l_last = np.zeros((n, 4), dtype=np.uint64); l_last[n - bf - 1] = one
    ext = lambda v: dom.coeff_to_extended(dom.lagrange_to_coeff(v))
    pk.l0, pk.l_last = ext(l0), ext(l_last)
    lb = ext(l_blind)
    pk.l_active_row = vsub(const_vec(dom.extended_len(), 1), vadd(pk.l_last, lb))
    return pk


# ------------------------------------------------------------------ helpers
def commit(bases, poly):
    """ParamsKZG::commit / commit_lagrange -> affine ints (None = identity)."""
    out = B.best_multiexp(poly, bases[: poly.shape[0]])
    return B.affine_to_ints(B.g1_batch_normalize(out))[0]


def eval_poly(poly, x):
    return I(B.eval_polynomial(poly, M(x)))


def rotate_omega(dom, x, rot):
    return I(dom.rotate_omega(M(x), rot))


def lagrange_interpolate(points, evals):
    """arithmetic::lagrange_interpolate: the unique polynomial of degree < len(points); ints."""
    m = len(points)
    poly = [0] * m
    for j in range(m):
        num = [1]
        den = 1
        for kx in range(m):
            if kx == j:
                continue
            new = [0] * (len(num) + 1)
            for i, c in enumerate(num):
                new[i] = (new[i] - c * points[kx]) % R
                new[i + 1] = (new[i + 1] + c) % R
            num = new
            den = den * (points[j] - points[kx]) % R
        s = evals[j] * pow(den, -1, R) % R
        for i, c in enumerate(num):
            poly[i] = (poly[i] + c * s) % R
    return poly


def permute_expression_pair(inp, tab, usable, rng, bf):
    """lookup::prover::permute_expression_pair; C++ restatement for speed."""
    a_sorted, s_vals = B.permute_expression_pair(inp, tab, usable)
    a_full = np.concatenate([a_sorted, rng.take(bf + 1)])
    s_full = np.concatenate([s_vals, rng.take(bf + 1)])
    return a_full, s_full


def permute_expression_pair_py(inp, tab, usable):
    """The same algorithm in plain Python (cross-check of the C++ restatement in tests)."""
    a_sorted_raw = B.sort_canonical(B.to_raw(inp[:usable]))
    a_sorted = B.from_raw(a_sorted_raw)
    keys = [tuple(r) for r in a_sorted_raw[:, ::-1].tolist()]          # most significant limb first
    t_raw = B.to_raw(tab[:usable])
    leftover = {}
    for r in t_raw[:, ::-1].tolist():
        leftover[tuple(r)] = leftover.get(tuple(r), 0) + 1
    s_raw = np.zeros((usable, 4), dtype=np.uint64)
    repeated = []
    for row in range(usable):
        if row == 0 or keys[row] != keys[row - 1]:
            s_raw[row] = a_sorted_raw[row]
            if keys[row] not in leftover or leftover[keys[row]] == 0:
                raise ValueError("ConstraintSystemFailure: lookup input not in table")
            leftover[keys[row]] -= 1
        else:
            repeated.append(row)
    for key in sorted(leftover):                                          # BTreeMap ascending order
        for _ in range(leftover[key]):
            s_raw[repeated.pop()] = np.array(key[::-1], dtype=np.uint64)
    assert not repeated
    return a_sorted, B.from_raw(s_raw)


# ------------------------------------------------------------------ create_proof
def create_proof(params_g, params_g_lagrange, pk, advice_in, instances, rng_wide, transcript_repr):
    """plonk::create_proof for one circuit, single phase.  Returns (proof bytes, trace dict)."""
    cs, dom, n, k = pk.cs, pk.dom, pk.n, pk.k
    bf = cs.blinding_factors()
    usable = n - (bf + 1)
    rng = Rng(rng_wide)
    tr = P.Blake2bTranscript()
    trace = {}
    rot_scale = 1 << (dom.extended_k - k)
    ext_n = dom.extended_len()

    tr.common_scalar(transcript_repr)                                    # vk.hash_into
    # -- instances
    inst_values = []
    for col in instances:
        assert len(col) <= usable, "InstanceTooLarge"
        for v in col:
            tr.common_scalar(v % R)
        arr = np.zeros((n, 4), dtype=np.uint64)
        if len(col):
            arr[: len(col)] = B.ints_to_mont([v % R for v in col])
        inst_values.append(arr)
    inst_polys = [dom.lagrange_to_coeff(v) for v in inst_values]
    # -- advice: blinding rows, blinds, commitments
    advice = [np.array(a, dtype=np.uint64).reshape(n, 4) for a in advice_in]
    for a in advice:
        a[usable:] = rng.take(bf + 1)
    for _ in advice:
        rng.one()                                                         # Blind (unused by KZG)
    for a in advice:
        tr.write_point(commit(params_g_lagrange, a))
    theta = tr.squeeze_challenge()
    # -- lookups: commit_permuted
    lookups = []
    for ins, tabs in cs.lookups:
        def compress(exprs):
            acc = np.zeros((n, 4), dtype=np.uint64)
            for e in exprs:
                acc = vadd(vscale(acc, theta), eval_expr(e, pk.fixed_values, advice, inst_values, cs, 1, n))
            return acc
        lk = {"ins": ins, "tabs": tabs}
        lk["cin"], lk["ctab"] = compress(ins), compress(tabs)
        lk["pin"], lk["ptab"] = permute_expression_pair(lk["cin"], lk["ctab"], usable, rng, bf)
        lk["pin_poly"] = dom.lagrange_to_coeff(lk["pin"]); rng.one()
        c_in = commit(params_g_lagrange, lk["pin"])
        lk["ptab_poly"] = dom.lagrange_to_coeff(lk["ptab"]); rng.one()
        c_tab = commit(params_g_lagrange, lk["ptab"])
        tr.write_point(c_in); tr.write_point(c_tab)
        lookups.append(lk)
    beta = tr.squeeze_challenge()
    gamma = tr.squeeze_challenge()
    # -- permutation commit
    def perm_column_values(ct, ci):
        return {0: advice, 1: pk.fixed_values, 2: inst_values}[ct][ci]
    sets = []
    chunk = cs.permutation_chunk_len()
    omega_pows = None
    if cs.permutation:
        omega_pows = B.powers(dom.omega, n)
    last_z, deltaomega = 1, 1
    for s0 in range(0, len(cs.permutation), chunk):
        cols = cs.permutation[s0:s0 + chunk]
        mod = const_vec(n, 1)
        for j, (ct, ci) in enumerate(cols):
            v = perm_column_values(ct, ci)
            mod = vmul(mod, vadd(vadds(vscale(pk.perm_values[s0 + j], beta), gamma), v))
        mod = B.batch_invert(mod)
        for ct, ci in cols:
            v = perm_column_values(ct, ci)
            mod = vmul(mod, vadd(vadds(vscale(omega_pows, deltaomega * beta % R), gamma), v))
            deltaomega = deltaomega * P.FR_DELTA % R
        z = B.prefix_product(mod, M(last_z))
        z[n - bf:] = rng.take(bf)
        last_z = I(z[n - (bf + 1)])
        rng.one()
        tr.write_point(commit(params_g_lagrange, z))
        zp = dom.lagrange_to_coeff(z)
        sets.append({"poly": zp, "coset": dom.coeff_to_extended(zp)})
    # -- lookups: commit_product
    for lk in lookups:
        den = vmul(vadds(lk["pin"], beta), vadds(lk["ptab"], gamma))
        den = B.batch_invert(den)
        prod = vmul(vmul(den, vadds(lk["cin"], beta)), vadds(lk["ctab"], gamma))
        z = np.concatenate([B.prefix_product(prod, M(1))[: n - bf], rng.take(bf)])
        rng.one()
        tr.write_point(commit(params_g_lagrange, z))
        lk["z_poly"] = dom.lagrange_to_coeff(z)
        trace.setdefault("lookup_z", []).append(z)
    # -- vanishing commit
    random_poly = rng.take(n); rng.one()
    tr.write_point(commit(params_g, random_poly))
    y = tr.squeeze_challenge()
    # -- advice polys
    advice_polys = [dom.lagrange_to_coeff(a) for a in advice]
    # -- evaluate_h
    adv_cos = [dom.coeff_to_extended(p) for p in advice_polys]
    inst_cos = [dom.coeff_to_extended(p) for p in inst_polys]
    h = np.zeros((ext_n, 4), dtype=np.uint64)
    def fold(hv, term):
        return vadd(vscale(hv, y), term)
    for g in cs.gates:
        h = fold(h, eval_expr(g, pk.fixed_cosets, adv_cos, inst_cos, cs, rot_scale, ext_n))
    one_v = const_vec(ext_n, 1)
    if sets:
        last_rot = -(bf + 1)
        first, last = sets[0]["coset"], sets[-1]["coset"]
        h = fold(h, vmul(vsub(one_v, first), pk.l0))
        h = fold(h, vmul(vsub(vmul(last, last), last), pk.l_last))
        for si in range(1, len(sets)):
            prev_last = np.roll(sets[si - 1]["coset"], -last_rot * rot_scale, axis=0)
            h = fold(h, vmul(vsub(sets[si]["coset"], prev_last), pk.l0))
        # beta_term[idx] = extended_omega^idx ; current_delta = beta * zeta * beta_term * delta^j
        ext_pows = B.powers(dom.extended_omega, ext_n)
        delta_pow = 1
        cos_cols = {0: adv_cos, 1: pk.fixed_cosets, 2: inst_cos}
        for si, s0 in enumerate(range(0, len(cs.permutation), chunk)):
            cols = cs.permutation[s0:s0 + chunk]
            zc = sets[si]["coset"]
            left = np.roll(zc, -rot_scale, axis=0)
            for j, (ct, ci) in enumerate(cols):
                left = vmul(left, vadds(vadd(cos_cols[ct][ci], vscale(pk.perm_cosets[s0 + j], beta)), gamma))
            right = zc
            for ct, ci in cols:
                cd = vscale(ext_pows, beta * P.FR_ZETA % R * delta_pow % R)
                right = vmul(right, vadds(vadd(cos_cols[ct][ci], cd), gamma))
                delta_pow = delta_pow * P.FR_DELTA % R
            h = fold(h, vmul(vsub(left, right), pk.l_active_row))
    for lk in lookups:
        zc = dom.coeff_to_extended(lk["z_poly"])
        ac = dom.coeff_to_extended(lk["pin_poly"])
        sc = dom.coeff_to_extended(lk["ptab_poly"])
        def compress_ext(exprs):
            acc = np.zeros((ext_n, 4), dtype=np.uint64)
            for e in exprs:
                acc = vadd(vscale(acc, theta), eval_expr(e, pk.fixed_cosets, adv_cos, inst_cos, cs, rot_scale, ext_n))
            return acc
        table_value = vmul(vadds(compress_ext(lk["ins"]), beta), vadds(compress_ext(lk["tabs"]), gamma))
        z_next = np.roll(zc, -rot_scale, axis=0)
        a_prev = np.roll(ac, rot_scale, axis=0)
        a_minus_s = vsub(ac, sc)
        h = fold(h, vmul(vsub(one_v, zc), pk.l0))
        h = fold(h, vmul(vsub(vmul(zc, zc), zc), pk.l_last))
        h = fold(h, vmul(vsub(vmul(vmul(z_next, vadds(ac, beta)), vadds(sc, gamma)), vmul(zc, table_value)), pk.l_active_row))
        h = fold(h, vmul(a_minus_s, pk.l0))
        h = fold(h, vmul(vmul(a_minus_s, vsub(ac, a_prev)), pk.l_active_row))
    trace["h_extended"] = h
    # -- vanishing construct
    h = dom.divide_by_vanishing_poly(h)
    h_coeffs = dom.extended_to_coeff(h)
    q = dom.quotient_poly_degree
    h_pieces = [np.ascontiguousarray(h_coeffs[i * n:(i + 1) * n]) for i in range(q)]
    for _ in h_pieces:
        rng.one()
    for hp in h_pieces:
        tr.write_point(commit(params_g, hp))
    x = tr.squeeze_challenge()
    xn = pow(x, n, R)
    # -- evaluations
    for col, rot in cs.advice_queries:
        tr.write_scalar(eval_poly(advice_polys[col], rotate_omega(dom, x, rot)))
    for col, rot in cs.fixed_queries:
        tr.write_scalar(eval_poly(pk.fixed_polys[col], rotate_omega(dom, x, rot)))
    h_poly = np.zeros((n, 4), dtype=np.uint64)
    for hp in reversed(h_pieces):
        h_poly = vadd(vscale(h_poly, xn), hp)
    tr.write_scalar(eval_poly(random_poly, x))
    for sp in pk.perm_polys:
        tr.write_scalar(eval_poly(sp, x))
    x_next, x_prev, x_last = rotate_omega(dom, x, 1), rotate_omega(dom, x, -1), rotate_omega(dom, x, -(bf + 1))
    for si, s in enumerate(sets):
        tr.write_scalar(eval_poly(s["poly"], x))
        tr.write_scalar(eval_poly(s["poly"], x_next))
        if si + 1 < len(sets):
            tr.write_scalar(eval_poly(s["poly"], x_last))
    for lk in lookups:
        tr.write_scalar(eval_poly(lk["z_poly"], x))
        tr.write_scalar(eval_poly(lk["z_poly"], x_next))
        tr.write_scalar(eval_poly(lk["pin_poly"], x))
        tr.write_scalar(eval_poly(lk["pin_poly"], x_prev))
        tr.write_scalar(eval_poly(lk["ptab_poly"], x))
    # -- multiopen queries: (poly object, point)
    queries = []
    for col, rot in cs.advice_queries:
        queries.append((advice_polys[col], rotate_omega(dom, x, rot)))
    for s in sets:
        queries.append((s["poly"], x)); queries.append((s["poly"], x_next))
    for s in list(reversed(sets))[1:]:
        queries.append((s["poly"], x_last))
    for lk in lookups:
        queries += [(lk["z_poly"], x), (lk["pin_poly"], x), (lk["ptab_poly"], x), (lk["pin_poly"], x_prev), (lk["z_poly"], x_next)]
    for col, rot in cs.fixed_queries:
        queries.append((pk.fixed_polys[col], rotate_omega(dom, x, rot)))
    for sp in pk.perm_polys:
        queries.append((sp, x))
    queries.append((h_poly, x)); queries.append((random_poly, x))
    shplonk_create_proof(params_g, n, queries, tr)
    trace.update(theta=theta, beta=beta, gamma=gamma, y=y, x=x)
    return bytes(tr.proof), trace


def construct_intermediate_sets(queries):
    """shplonk.rs: commitments identified by polynomial identity; point sets ordered as BTreeSet<Fr>."""
    comm_sets = []                                   # [(poly, set(points))] in first-seen order
    super_points = set()
    for poly, pt in queries:
        super_points.add(pt)
        for entry in comm_sets:
            if entry[0] is poly:
                entry[1].add(pt)
                break
        else:
            comm_sets.append([poly, {pt}])
    rot_sets = []                                    # [(frozenset(points), [polys])]
    for poly, pts in comm_sets:
        fs = frozenset(pts)
        for entry in rot_sets:
            if entry[0] == fs:
                if not any(p is poly for p in entry[1]):
                    entry[1].append(poly)
                break
        else:
            rot_sets.append((fs, [poly]))
    return [(sorted(fs), polys) for fs, polys in rot_sets], sorted(super_points)


def shplonk_create_proof(params_g, n, queries, tr):
    y = tr.squeeze_challenge()
    rot_sets, super_points = construct_intermediate_sets(queries)
    ext = []
    for points, polys in rot_sets:
        items = []
        for poly in polys:
            evals = [eval_poly(poly, pt) for pt in points]
            items.append((poly, lagrange_interpolate(points, evals)))
        ext.append((points, items))
    v = tr.squeeze_challenge()

    def div_by_vanishing(poly, roots):
        for r in roots:
            poly = B.kate_division(poly, M(r))
        return poly

    h_x = None
    v_pow = 1
    for points, items in ext:
        n_x, y_pow = None, 1
        for poly, low in items:
            num = np.array(poly, dtype=np.uint64)
            num[: len(low)] = vsub(num[: len(low)], B.ints_to_mont(low))
            num = vscale(num, y_pow)
            n_x = num if n_x is None else vadd(n_x, num)
            y_pow = y_pow * y % R
        qpoly = div_by_vanishing(n_x, points)
        qpoly = np.concatenate([qpoly, np.zeros((n - qpoly.shape[0], 4), dtype=np.uint64)])
        qpoly = vscale(qpoly, v_pow)
        h_x = qpoly if h_x is None else vadd(h_x, qpoly)
        v_pow = v_pow * v % R
    tr.write_point(commit(params_g, h_x))
    u = tr.squeeze_challenge()

    def vanish_eval(roots, z):
        acc = 1
        for r in roots:
            acc = acc * (z - r) % R
        return acc

    l_x, z_diffs, v_pow = None, [], 1
    for points, items in ext:
        diffs = [p for p in super_points if p not in points]
        z_i = vanish_eval(diffs, u)
        z_diffs.append(z_i)
        inner, y_pow = None, 1
        for poly, low in items:
            r_eval = sum(c * pow(u, i, R) for i, c in enumerate(low)) % R
            t = np.array(poly, dtype=np.uint64)
            t[0] = M(I(t[0]) - r_eval)
            t = vscale(t, y_pow)
            inner = t if inner is None else vadd(inner, t)
            y_pow = y_pow * y % R
        contrib = vscale(vscale(inner, z_i), v_pow)
        l_x = contrib if l_x is None else vadd(l_x, contrib)
        v_pow = v_pow * v % R
    zt_eval = vanish_eval(super_points, u)
    l_x = vsub(l_x, vscale(h_x, zt_eval))
    h2 = B.kate_division(l_x, M(u))
    h2 = vscale(h2, pow(z_diffs[0], -1, R))
    tr.write_point(commit(params_g, h2))


# ------------------------------------------------------------------ verifier (trapdoor form)
class _Reader:
    def __init__(self, proof):
        self.h = __import__("hashlib").blake2b(digest_size=64, person=b"Halo2-Transcript")
        self.buf, self.pos = proof, 0

    def common_scalar(self, s):
        self.h.update(b"\x02" + s.to_bytes(32, "little"))

    def read_point(self):
        b = self.buf[self.pos:self.pos + 32]; self.pos += 32
        pt = decompress(b)
        self.h.update(b"\x01")
        xx, yy = (0, 0) if pt is None else pt
        self.h.update(xx.to_bytes(32, "little") + yy.to_bytes(32, "little"))
        return pt

    def read_scalar(self):
        s = int.from_bytes(self.buf[self.pos:self.pos + 32], "little"); self.pos += 32
        assert s < R
        self.common_scalar(s)
        return s

    def squeeze(self):
        self.h.update(b"\x00")
        return int.from_bytes(self.h.copy().digest(), "little") % R


def decompress(b):
    b = bytearray(b)
    sign = b[31] >> 7
    b[31] &= 0x7F
    x = int.from_bytes(b, "little")
    if x == 0 and sign == 0:
        return None
    Q = P.Q_MOD
    y = pow((x * x * x + 3) % Q, (Q + 1) // 4, Q)
    assert y * y % Q == (x * x * x + 3) % Q, "not on curve"
    if (y & 1) != sign:
        y = Q - y
    return (x, y)


def verify_proof(s_secret, pk, instances, proof, transcript_repr):
    """plonk::verify_proof + VerifierSHPLONK restated.  The final pairing check
    e(h2, [s]_2) = e(u*h2 + L, [1]_2) is checked as  s*h2 == u*h2 + L  in G1 using the SRS secret."""
    cs, dom, n, k = pk.cs, pk.dom, pk.n, pk.k
    bf = cs.blinding_factors()
    rd = _Reader(proof)
    rd.common_scalar(transcript_repr)
    for col in instances:
        for v in col:
            rd.common_scalar(v % R)
    A, L = cs.num_advice, len(cs.lookups)
    S, q = cs.num_permutation_sets(), cs.degree() - 1
    advice_c = [rd.read_point() for _ in range(A)]
    theta = rd.squeeze()
    lk_perm = [(rd.read_point(), rd.read_point()) for _ in range(L)]
    beta, gamma = rd.squeeze(), rd.squeeze()
    perm_c = [rd.read_point() for _ in range(S)]
    lk_prod = [rd.read_point() for _ in range(L)]
    random_c = rd.read_point()
    y = rd.squeeze()
    h_c = [rd.read_point() for _ in range(q)]
    x = rd.squeeze()
    xn = pow(x, n, R)
    advice_evals = [rd.read_scalar() for _ in cs.advice_queries]
    fixed_evals = [rd.read_scalar() for _ in cs.fixed_queries]
    random_eval = rd.read_scalar()
    sigma_evals = [rd.read_scalar() for _ in cs.permutation]
    perm_evals = []
    for si in range(S):
        e = [rd.read_scalar(), rd.read_scalar()]
        e.append(rd.read_scalar() if si + 1 < S else None)
        perm_evals.append(e)
    lk_evals = [[rd.read_scalar() for _ in range(5)] for _ in range(L)]   # z, z_next, a, a_prev, s
    # instance evals (KZG: computed by the verifier from the public inputs via Lagrange basis)
    w = I(dom.omega)
    instance_evals = []
    for col, rot in cs.instance_queries:
        vals = instances[col]
        acc = 0
        for i, v in enumerate(vals):
            # l_i(x*w^rot) = (x'^n - 1)/n * w^i / (x' - w^i)
            xr = rotate_omega(dom, x, rot)
            wi = pow(w, i, R)
            li = (pow(xr, n, R) - 1) * pow(n, -1, R) % R * wi % R * pow(xr - wi, -1, R) % R
            acc = (acc + v * li) % R
        instance_evals.append(acc)
    # l_0, l_last, l_blind at x
    def l_i(i):
        wi = pow(w, i % n, R)
        return (xn - 1) * pow(n, -1, R) % R * wi % R * pow(x - wi, -1, R) % R
    l_0 = l_i(0)
    l_last = l_i(n - bf - 1)
    l_blind = sum(l_i(n - 1 - j) for j in range(bf)) % R
    l_active = (1 - l_last - l_blind) % R
    # expected h(x)
    terms = [eval_expr_at(g, fixed_evals, advice_evals, instance_evals) for g in cs.gates]
    def any_eval(ct, ci):
        if ct == 0:
            return advice_evals[cs.advice_queries.index((ci, 0))]
        if ct == 1:
            return fixed_evals[cs.fixed_queries.index((ci, 0))]
        return instance_evals[cs.instance_queries.index((ci, 0))]
    if S:
        chunk = cs.permutation_chunk_len()
        terms.append(l_0 * (1 - perm_evals[0][0]) % R)
        zl = perm_evals[-1][0]
        terms.append(l_last * (zl * zl - zl) % R)
        for si in range(1, S):
            terms.append(l_0 * (perm_evals[si][0] - perm_evals[si - 1][2]) % R)
        for si in range(S):
            cols = cs.permutation[si * chunk:(si + 1) * chunk]
            left = perm_evals[si][1]
            for j, (ct, ci) in enumerate(cols):
                left = left * (any_eval(ct, ci) + beta * sigma_evals[si * chunk + j] + gamma) % R
            right = perm_evals[si][0]
            cur = beta * x % R * pow(P.FR_DELTA, si * chunk, R) % R
            for ct, ci in cols:
                right = right * (any_eval(ct, ci) + cur + gamma) % R
                cur = cur * P.FR_DELTA % R
            terms.append(l_active * (left - right) % R)
    for (ins, tabs), ev in zip(cs.lookups, lk_evals):
        z, zn, a, ap, s = ev
        def comp(exprs):
            acc = 0
            for e in exprs:
                acc = (acc * theta + eval_expr_at(e, fixed_evals, advice_evals, instance_evals)) % R
            return acc
        terms.append(l_0 * (1 - z) % R)
        terms.append(l_last * (z * z - z) % R)
        terms.append(l_active * (zn * (a + beta) % R * (s + gamma) - z * (comp(ins) + beta) % R * (comp(tabs) + gamma)) % R)
        terms.append(l_0 * (a - s) % R)
        terms.append(l_active * ((a - s) * (a - ap) % R) % R)
    expected_h = 0
    for t in terms:
        expected_h = (expected_h * y + t) % R
    expected_h = expected_h * pow(xn - 1, -1, R) % R
    # h commitment folded in x^n
    def pmul(pt, s):
        return P.g1_mul(pt, s) if pt is not None else None
    h_commit = None
    for c in reversed(h_c):
        h_commit = P.g1_add(pmul(h_commit, xn), c)
    # queries: (commitment point, point, eval)
    x_next, x_prev, x_last = rotate_omega(dom, x, 1), rotate_omega(dom, x, -1), rotate_omega(dom, x, -(bf + 1))
    Q = []
    for (col, rot), e in zip(cs.advice_queries, advice_evals):
        Q.append((("adv", col), advice_c[col], rotate_omega(dom, x, rot), e))
    for si in range(S):
        Q.append((("pz", si), perm_c[si], x, perm_evals[si][0])); Q.append((("pz", si), perm_c[si], x_next, perm_evals[si][1]))
    for si in reversed(range(S - 1)):
        Q.append((("pz", si), perm_c[si], x_last, perm_evals[si][2]))
    for li, ev in enumerate(lk_evals):
        z, zn, a, ap, s = ev
        Q += [(("lz", li), lk_prod[li], x, z), (("la", li), lk_perm[li][0], x, a), (("ls", li), lk_perm[li][1], x, s),
              (("la", li), lk_perm[li][0], x_prev, ap), (("lz", li), lk_prod[li], x_next, zn)]
    # fixed and sigma commitments: recomputed from the pk polys (they belong to the vk)
    return Q, dict(expected_h=expected_h, h_commit=h_commit, random_c=random_c, random_eval=random_eval, x=x, rd=rd,
                   fixed_evals=fixed_evals, sigma_evals=sigma_evals)


def verifying_key(cs, k, fixed_commitments, sigma_commitments):
    """What verify_proof needs of a VerifyingKey: the constraint system, the domain and the
    commitments to the fixed and permutation polynomials (affine integer pairs, None = identity)."""
    vk = ProvingKey()
    vk.cs, vk.k, vk.n = cs, k, 1 << k
    vk.dom = B.Domain(cs.degree(), k)
    vk.fixed_commitments, vk.sigma_commitments = list(fixed_commitments), list(sigma_commitments)
    return vk


def verify_full(s_secret, params_g, pk, instances, proof, transcript_repr, s_g2=None):
    """Complete check, including the SHPLONK opening: against the known secret (s_secret), or — as
    halo2's verifier does it (shplonk/verifier.rs, DualMSM::check) — by pairing against the SRS's
    [s]_2 when `s_g2` is given and s_secret is None:  e(h2, [s]_2) = e(u h2 + L / z_0, [1]_2)."""
    cs, dom = pk.cs, pk.dom
    Q, st = verify_proof(s_secret, pk, instances, proof, transcript_repr)
    x, rd = st["x"], st["rd"]
    have_vk = hasattr(pk, "fixed_commitments")                   # verifying_key(): commitments given, no polynomials
    fixed_c = {}
    for (col, rot), e in zip(cs.fixed_queries, st["fixed_evals"]):
        if col not in fixed_c:
            fixed_c[col] = pk.fixed_commitments[col] if have_vk else commit(params_g, pk.fixed_polys[col])
        Q.append((("fix", col), fixed_c[col], rotate_omega(dom, x, rot), e))
    for j, e in enumerate(st["sigma_evals"]):
        Q.append((("sig", j), pk.sigma_commitments[j] if have_vk else commit(params_g, pk.perm_polys[j]), x, e))
    Q.append((("h",), st["h_commit"], x, st["expected_h"]))
    Q.append((("rand",), st["random_c"], x, st["random_eval"]))
    y = rd.squeeze()
    v = rd.squeeze()
    h1 = rd.read_point()
    u = rd.squeeze()
    h2 = rd.read_point()
    assert rd.pos == len(proof), "trailing bytes in proof"
    # intermediate sets on (id, commitment, point, eval)
    comm_sets, super_points = [], set()
    for cid, c, pt, e in Q:
        super_points.add(pt)
        for entry in comm_sets:
            if entry[0] == cid:
                entry[2][pt] = e
                break
        else:
            comm_sets.append([cid, c, {pt: e}])
    rot_sets = []
    for cid, c, pe in comm_sets:
        fs = frozenset(pe)
        for entry in rot_sets:
            if entry[0] == fs:
                entry[1].append((c, pe))
                break
        else:
            rot_sets.append((fs, [(c, pe)]))
    super_points = sorted(super_points)
    def vanish_eval(roots, z):
        acc = 1
        for r in roots:
            acc = acc * (z - r) % R
        return acc
    # L = sum_i v^i z_i sum_j y^j (C_ij - [r_ij(u)]G)  -  Z_T(u) h1 ;   check  s*h2 == u*h2 + L / z_0
    Lpt, v_pow, z0 = None, 1, None
    G = P.G1_GEN
    for fs, items in rot_sets:
        points = sorted(fs)
        z_i = vanish_eval([p for p in super_points if p not in points], u)
        if z0 is None:
            z0 = z_i
        inner, y_pow = None, 1
        for c, pe in items:
            low = lagrange_interpolate(points, [pe[p] for p in points])
            r_eval = sum(cf * pow(u, i, R) for i, cf in enumerate(low)) % R
            t = P.g1_add(c, P.g1_mul(G, (-r_eval) % R))
            inner = P.g1_add(inner, P.g1_mul(t, y_pow) if t is not None else None)
            y_pow = y_pow * y % R
        Lpt = P.g1_add(Lpt, P.g1_mul(inner, z_i * v_pow % R) if inner is not None else None)
        v_pow = v_pow * v % R
    zt = vanish_eval(super_points, u)
    Lpt = P.g1_add(Lpt, P.g1_mul(h1, (-zt) % R) if h1 is not None else None)
    Lpt = P.g1_mul(Lpt, pow(z0, -1, R)) if Lpt is not None else None
    rhs = P.g1_add(P.g1_mul(h2, u) if h2 is not None else None, Lpt)
    if s_secret is None:
        from oracle import pairing as PR
        neg_rhs = None if rhs is None else (rhs[0], (-rhs[1]) % PR.Q)
        return PR.pairing_product_is_one([(h2, s_g2), (neg_rhs, PR.G2_GEN)])
    lhs = P.g1_mul(h2, s_secret) if h2 is not None else None
    return lhs == rhs
