// ORACLE — TEST INFRASTRUCTURE ONLY (see ff.hpp header; PARITY UNPINNED vs Rust).
// Flat C entry points over the restatement so tests/ and bench.py's cpu_baseline
// leg can drive it through ctypes.  All field elements cross this boundary as
// 4 x u64 LE limbs in Montgomery form (the in-memory layout of halo2curves'
// Fr / Fq), points as G1Affine {x,y} (64 B) or G1 {x,y,z} (96 B).
#include "arith.hpp"
#include <cstdio>
#include <algorithm>
#include <map>

using namespace orc;

extern "C" {

void orc_init() { init_fields(); }
void orc_set_threads(int t) { set_num_threads(t); }
int orc_get_threads() { return num_threads(); }

// ---- field element helpers; which = 0 -> Fr, 1 -> Fq ----
#define FIELD_DISPATCH(EXPR_FR, EXPR_FQ) do { if (which == 0) { EXPR_FR; } else { EXPR_FQ; } } while (0)

void orc_f_from_raw(int which, const uint64_t* raw, uint64_t* out, size_t n) {
    init_fields();
    for (size_t i = 0; i < n; ++i)
        FIELD_DISPATCH(*(Fr*)(out + 4 * i) = Fr::from_raw(raw + 4 * i), *(Fq*)(out + 4 * i) = Fq::from_raw(raw + 4 * i));
}
void orc_f_to_raw(int which, const uint64_t* in, uint64_t* raw, size_t n) {
    init_fields();
    for (size_t i = 0; i < n; ++i)
        FIELD_DISPATCH(((const Fr*)(in + 4 * i))->to_raw(raw + 4 * i), ((const Fq*)(in + 4 * i))->to_raw(raw + 4 * i));
}
void orc_f_from_u512(int which, const uint64_t* wide, uint64_t* out, size_t n) {
    init_fields();
    for (size_t i = 0; i < n; ++i)
        FIELD_DISPATCH(*(Fr*)(out + 4 * i) = Fr::from_u512(wide + 8 * i), *(Fq*)(out + 4 * i) = Fq::from_u512(wide + 8 * i));
}
// op: 0 add, 1 sub, 2 mul ; elementwise over n
void orc_f_binop(int which, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    init_fields();
    auto run = [&](auto tag) {
        typedef decltype(tag) F;
        const F* A = (const F*)a; const F* B = (const F*)b; F* O = (F*)out;
        parallelize(n, [&](size_t s, size_t e) {
            for (size_t i = s; i < e; ++i) O[i] = op == 0 ? A[i] + B[i] : op == 1 ? A[i] - B[i] : A[i] * B[i];
        });
    };
    if (which == 0) run(Fr()); else run(Fq());
}
void orc_f_inv(int which, const uint64_t* a, uint64_t* out, size_t n) {
    init_fields();
    for (size_t i = 0; i < n; ++i)
        FIELD_DISPATCH(*(Fr*)(out + 4 * i) = ((const Fr*)(a + 4 * i))->inv(), *(Fq*)(out + 4 * i) = ((const Fq*)(a + 4 * i))->inv());
}
void orc_fr_batch_invert(uint64_t* a, size_t n) { init_fields(); batch_invert((Fr*)a, n); }
void orc_fr_pow(const uint64_t* a, const uint64_t* e4, uint64_t* out) { init_fields(); *(Fr*)out = ((const Fr*)a)->pow(e4); }
void orc_fr_constants(uint64_t* root_of_unity, uint64_t* delta, uint64_t* zeta) {
    init_fields();
    memcpy(root_of_unity, FR_ROOT_OF_UNITY.l, 32); memcpy(delta, FR_DELTA.l, 32); memcpy(zeta, FR_ZETA.l, 32);
}
void orc_field_params(int which, uint64_t* p, uint64_t* inv, uint64_t* r, uint64_t* r2, uint64_t* r3) {
    init_fields();
    const FieldParams& P = which == 0 ? Fr::P : Fq::P;
    memcpy(p, P.p, 32); *inv = P.inv; memcpy(r, P.r, 32); memcpy(r2, P.r2, 32); memcpy(r3, P.r3, 32);
}

// ---- G1 ----
void orc_g1_generator(uint64_t* out_affine) { init_fields(); G1Affine g = G1::generator().to_affine(); memcpy(out_affine, &g, 64); }
void orc_g1_add(const uint64_t* a, const uint64_t* b, uint64_t* out) { init_fields(); *(G1*)out = ((const G1*)a)->add(*(const G1*)b); }
void orc_g1_add_affine(const uint64_t* a, const uint64_t* b_aff, uint64_t* out) { init_fields(); *(G1*)out = ((const G1*)a)->add_affine(*(const G1Affine*)b_aff); }
void orc_g1_double(const uint64_t* a, uint64_t* out) { init_fields(); *(G1*)out = ((const G1*)a)->dbl(); }
void orc_g1_mul(const uint64_t* a, const uint64_t* fr_scalar, uint64_t* out) { init_fields(); *(G1*)out = ((const G1*)a)->mul(*(const Fr*)fr_scalar); }
void orc_g1_from_affine(const uint64_t* a, uint64_t* out, size_t n) {
    init_fields(); for (size_t i = 0; i < n; ++i) ((G1*)out)[i] = G1::from_affine(((const G1Affine*)a)[i]);
}
void orc_g1_batch_normalize(const uint64_t* jac, uint64_t* aff, size_t n) { init_fields(); batch_normalize((const G1*)jac, (G1Affine*)aff, n); }
void orc_g1_compress(const uint64_t* aff, uint8_t* out32, size_t n) {
    init_fields(); for (size_t i = 0; i < n; ++i) ((const G1Affine*)aff)[i].to_bytes(out32 + 32 * i);
}
int orc_g1_on_curve(const uint64_t* aff, size_t n) {
    init_fields(); for (size_t i = 0; i < n; ++i) if (!((const G1Affine*)aff)[i].on_curve()) return 0; return 1;
}

// Fixed-base table for the generator: tab[w][d-1] = d * 2^(8w) * G  (affine), w < 32, d in 1..255
static std::vector<G1Affine> g_gen_table;
static void build_gen_table() {
    if (!g_gen_table.empty()) return;
    std::vector<G1> jac(32 * 255);
    G1 base = G1::generator();
    for (int w = 0; w < 32; ++w) {
        G1 cur = base;
        for (int d = 1; d <= 255; ++d) { jac[w * 255 + d - 1] = cur; cur = cur.add(base); }
        base = cur;                                     // 256 * base
    }
    g_gen_table.resize(32 * 255);
    batch_normalize(jac.data(), g_gen_table.data(), jac.size());
}
static G1 gen_mul(const Fr& s) {
    uint64_t e[4]; s.to_raw(e);
    const uint8_t* b = (const uint8_t*)e;
    G1 acc = G1::identity();
    for (int w = 0; w < 32; ++w) if (b[w]) acc = acc.add_affine(g_gen_table[w * 255 + b[w] - 1]);
    return acc;
}
// out[i] = [scalars[i]] G, affine
void orc_g1_fixed_base_mul(const uint64_t* scalars, uint64_t* out_aff, size_t n) {
    init_fields(); build_gen_table();
    std::vector<G1> jac(n);
    parallelize(n, [&](size_t s, size_t e) { for (size_t i = s; i < e; ++i) jac[i] = gen_mul(((const Fr*)scalars)[i]); });
    batch_normalize(jac.data(), (G1Affine*)out_aff, n);
}

// ParamsKZG::setup (poly/kzg/commitment.rs): g[i] = [s^i]G, g_lagrange[i] = [l_i(s)]G
void orc_params_setup(unsigned k, const uint64_t* s_mont, uint64_t* g, uint64_t* g_lagrange) {
    init_fields(); build_gen_table();
    size_t n = (size_t)1 << k;
    Fr s = *(const Fr*)s_mont;
    std::vector<Fr> sc(n);
    Fr cur = Fr::one();
    for (size_t i = 0; i < n; ++i) { sc[i] = cur; cur *= s; }
    orc_g1_fixed_base_mul((const uint64_t*)sc.data(), g, n);
    if (!g_lagrange) return;
    Fr root = FR_ROOT_OF_UNITY;
    for (unsigned i = k; i < FR_S; ++i) root = root.sqr();
    Fr n_inv = Fr::from_u64(n).inv();
    Fr multiplier = (s.pow_u64(n) - Fr::one()) * n_inv;
    std::vector<Fr> den(n), rp(n);
    cur = Fr::one();
    for (size_t i = 0; i < n; ++i) { rp[i] = cur; den[i] = s - cur; cur *= root; }
    batch_invert(den.data(), n);
    for (size_t i = 0; i < n; ++i) sc[i] = multiplier * rp[i] * den[i];
    orc_g1_fixed_base_mul((const uint64_t*)sc.data(), g_lagrange, n);
}

// ---- arithmetic.rs ----
void orc_best_multiexp(const uint64_t* coeffs, const uint64_t* bases, size_t n, uint64_t* out_jac) {
    init_fields(); *(G1*)out_jac = best_multiexp((const Fr*)coeffs, (const G1Affine*)bases, n);
}
void orc_best_fft(uint64_t* a, const uint64_t* omega, unsigned log_n) { init_fields(); best_fft((Fr*)a, *(const Fr*)omega, log_n); }
void orc_eval_polynomial(const uint64_t* poly, size_t n, const uint64_t* x, uint64_t* out) {
    init_fields(); *(Fr*)out = eval_polynomial((const Fr*)poly, n, *(const Fr*)x);
}
void orc_kate_division(const uint64_t* a, size_t n, const uint64_t* b, uint64_t* q) {
    init_fields(); kate_division((const Fr*)a, n, *(const Fr*)b, (Fr*)q);
}

// ---- poly/domain.rs ----
void* orc_domain_new(unsigned j, unsigned k) { init_fields(); return new Domain(j, k); }
void orc_domain_free(void* d) { delete (Domain*)d; }
// fields: 0 omega 1 omega_inv 2 extended_omega 3 extended_omega_inv 4 g_coset 5 g_coset_inv
//         6 ifft_divisor 7 extended_ifft_divisor 8 barycentric_weight
void orc_domain_get(void* d_, int field, uint64_t* out) {
    Domain* d = (Domain*)d_;
    const Fr* f[] = {&d->omega, &d->omega_inv, &d->extended_omega, &d->extended_omega_inv, &d->g_coset,
                     &d->g_coset_inv, &d->ifft_divisor, &d->extended_ifft_divisor, &d->barycentric_weight};
    memcpy(out, f[field]->l, 32);
}
unsigned orc_domain_extended_k(void* d) { return ((Domain*)d)->extended_k; }
unsigned orc_domain_quotient_degree(void* d) { return (unsigned)((Domain*)d)->quotient_poly_degree; }
void orc_domain_t_evaluations(void* d_, uint64_t* out) {
    Domain* d = (Domain*)d_; memcpy(out, d->t_evaluations.data(), d->t_evaluations.size() * 32);
}
void orc_domain_lagrange_to_coeff(void* d, uint64_t* a) { ((Domain*)d)->lagrange_to_coeff((Fr*)a); }
void orc_domain_coeff_to_extended(void* d, const uint64_t* a, uint64_t* out) { ((Domain*)d)->coeff_to_extended((const Fr*)a, (Fr*)out); }
void orc_domain_extended_to_coeff(void* d, uint64_t* a) { ((Domain*)d)->extended_to_coeff((Fr*)a); }
void orc_domain_divide_by_vanishing_poly(void* d, uint64_t* a) { ((Domain*)d)->divide_by_vanishing_poly((Fr*)a); }
void orc_domain_rotate_omega(void* d, const uint64_t* x, int rot, uint64_t* out) {
    *(Fr*)out = ((Domain*)d)->rotate_omega(*(const Fr*)x, rot);
}

}  // extern "C"

// ---- extras used by oracle/prover.py (test infrastructure) ----
extern "C" {

// op: 0 a+s, 1 a-s, 2 a*s  with a scalar s broadcast over n elements (Fr)
void orc_fr_binop_scalar(int op, const uint64_t* a, const uint64_t* s, uint64_t* out, size_t n) {
    init_fields();
    const Fr* A = (const Fr*)a; Fr S = *(const Fr*)s; Fr* O = (Fr*)out;
    parallelize(n, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) O[i] = op == 0 ? A[i] + S : op == 1 ? A[i] - S : A[i] * S;
    });
}

// rand_xorshift::XorShiftRng: fills count x 8 u64 (the 512-bit inputs of Fr::random), advancing state[4]
void orc_xorshift_fill_wide(uint32_t* state, uint64_t* out, size_t count) {
    uint32_t x = state[0], y = state[1], z = state[2], w = state[3];
    for (size_t i = 0; i < count * 8; ++i) {
        uint32_t lo, hi;
        { uint32_t t = x ^ (x << 11); x = y; y = z; z = w; w = w ^ (w >> 19) ^ (t ^ (t >> 8)); lo = w; }
        { uint32_t t = x ^ (x << 11); x = y; y = z; z = w; w = w ^ (w >> 19) ^ (t ^ (t >> 8)); hi = w; }
        out[i] = (uint64_t)lo | ((uint64_t)hi << 32);
    }
    state[0] = x; state[1] = y; state[2] = z; state[3] = w;
}

// z[0] = z0, z[i] = z[i-1] * p[i-1]  (the serial grand-product loops of permutation/lookup provers)
void orc_prefix_product(const uint64_t* p, const uint64_t* z0, uint64_t* z, size_t n) {
    init_fields();
    const Fr* P = (const Fr*)p; Fr* Z = (Fr*)z;
    Fr run = *(const Fr*)z0;
    for (size_t i = 0; i < n; ++i) { Z[i] = run; run = run * P[i]; }
}

// out[i] = base^i
void orc_powers(const uint64_t* base, uint64_t* out, size_t n) {
    init_fields();
    Fr b = *(const Fr*)base, cur = Fr::one();
    for (size_t i = 0; i < n; ++i) { ((Fr*)out)[i] = cur; cur *= b; }
}

// lookup::prover::permute_expression_pair on the first `usable` rows (halo2_proofs v2023_02_02
// plonk/lookup/prover.rs): sort the input, BTreeMap multiset of the table, fill repeated rows
// from the leftover table values in ascending order, popping rows from the back.
// Returns 0 on success, 1 if an input value is missing from the table (ConstraintSystemFailure).
int orc_permute_expression_pair(const uint64_t* input, const uint64_t* table, size_t usable, uint64_t* a_out, uint64_t* s_out) {
    init_fields();
    struct Less { bool operator()(const Fr& x, const Fr& y) const { return Fr::less(x, y); } };
    std::vector<Fr> a((const Fr*)input, (const Fr*)input + usable);
    // sort by canonical value; compare on canonical limbs to avoid repeated conversions
    std::vector<std::array<uint64_t, 5>> keyed(usable);
    for (size_t i = 0; i < usable; ++i) { uint64_t r[4]; a[i].to_raw(r); keyed[i] = {r[3], r[2], r[1], r[0], (uint64_t)i}; }
    std::stable_sort(keyed.begin(), keyed.end(), [](const std::array<uint64_t, 5>& x, const std::array<uint64_t, 5>& y) {
        for (int i = 0; i < 4; ++i) if (x[i] != y[i]) return x[i] < y[i];
        return false;
    });
    std::vector<Fr> sorted(usable);
    for (size_t i = 0; i < usable; ++i) sorted[i] = a[keyed[i][4]];
    std::map<Fr, uint32_t, Less> leftover;
    for (size_t i = 0; i < usable; ++i) leftover[((const Fr*)table)[i]] += 1;
    std::vector<Fr> s(usable, Fr::zero());
    std::vector<size_t> repeated;
    for (size_t row = 0; row < usable; ++row) {
        if (row == 0 || sorted[row] != sorted[row - 1]) {
            s[row] = sorted[row];
            auto it = leftover.find(sorted[row]);
            if (it == leftover.end() || it->second == 0) return 1;
            it->second -= 1;
        } else repeated.push_back(row);
    }
    for (auto& kv : leftover) for (uint32_t c = 0; c < kv.second; ++c) { s[repeated.back()] = kv.first; repeated.pop_back(); }
    if (!repeated.empty()) return 1;
    memcpy(a_out, sorted.data(), usable * 32); memcpy(s_out, s.data(), usable * 32);
    return 0;
}

// sort rows of canonical 4xu64 values ascending by numeric value (halo2curves `Ord for Fr`)
void orc_sort_canonical(uint64_t* a, size_t n) {
    typedef std::array<uint64_t, 4> K;
    K* k = (K*)a;
    std::sort(k, k + n, [](const K& x, const K& y) {
        for (int i = 3; i >= 0; --i) if (x[i] != y[i]) return x[i] < y[i];
        return false;
    });
}

}  // extern "C"
