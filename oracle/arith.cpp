// ORACLE — TEST INFRASTRUCTURE ONLY (see ff.hpp header; PARITY UNPINNED vs Rust).
// Restates halo2_proofs v2023_02_02 src/arithmetic.rs + src/poly/domain.rs
// (third-party dependency pinned at /root/reference/Cargo.toml:10; source absent here).
#include "arith.hpp"
#include <cmath>
#include <cassert>
#include <atomic>

namespace orc {

// ---------------------------------------------------------------- fields ----
FieldParams make_params(const uint64_t p[4]) {
    FieldParams P;
    memcpy(P.p, p, 32);
    uint64_t inv = 1;                                  // Newton: inv = p^-1 mod 2^64
    for (int i = 0; i < 63; ++i) { inv *= inv; inv *= p[0]; }
    P.inv = (uint64_t)0 - inv;
    auto dbl_mod = [&](uint64_t v[4]) {
        uint64_t c = add4(v, v, v);
        if (c || geq4(v, p)) sub4(v, v, p);
    };
    uint64_t v[4] = {1, 0, 0, 0};
    for (int i = 0; i < 256; ++i) dbl_mod(v);
    memcpy(P.r, v, 32);
    for (int i = 0; i < 256; ++i) dbl_mod(v);
    memcpy(P.r2, v, 32);
    for (int i = 0; i < 256; ++i) dbl_mod(v);
    memcpy(P.r3, v, 32);
    return P;
}

Fr FR_ROOT_OF_UNITY, FR_DELTA, FR_ZETA;

void init_fields() {
    static std::atomic<bool> done{false};
    if (done.load()) return;
    static const uint64_t r[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL,
                                  0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    static const uint64_t q[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL,
                                  0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    Fr::P = make_params(r);
    Fq::P = make_params(q);
    // ROOT_OF_UNITY = 7^((r-1)/2^28), DELTA = 7^(2^28)
    uint64_t e[4] = {r[0] - 1, r[1], r[2], r[3]};
    for (int i = 0; i < 4; ++i) {                       // e >>= 28
        e[i] = (e[i] >> 28) | (i < 3 ? (e[i + 1] << 36) : 0);
    }
    Fr seven = Fr::from_u64(7);
    FR_ROOT_OF_UNITY = seven.pow(e);
    FR_DELTA = seven.pow_u64(1ULL << 28);
    static const uint64_t zeta[4] = {0xb8ca0b2d36636f23ULL, 0xcc37a73fec2bc5e9ULL,
                                     0x048b6e193fd84104ULL, 0x30644e72e131a029ULL};
    FR_ZETA = Fr::from_raw(zeta);
    done.store(true);
}

// --------------------------------------------------------------- threads ----
static int g_threads = 0;
int num_threads() {
    if (g_threads > 0) return g_threads;
    int t = (int)std::thread::hardware_concurrency();
    return t > 0 ? t : 1;
}
void set_num_threads(int t) { g_threads = t; }

void parallelize(size_t n, const std::function<void(size_t, size_t)>& f) {
    size_t T = (size_t)num_threads();
    if (n < T || T == 1) { f(0, n); return; }
    size_t chunk = n / T;                               // upstream: chunk = n / threads
    std::vector<std::thread> th;
    for (size_t s = 0; s < n; s += chunk) {
        size_t e = s + chunk < n ? s + chunk : n;
        th.emplace_back([=, &f] { f(s, e); });
    }
    for (auto& t : th) t.join();
}

// ------------------------------------------------------------------- MSM ----
namespace {
struct Bucket {                                         // enum Bucket { None, Affine, Projective }
    int kind = 0; G1Affine a; G1 p;
    void add_assign(const G1Affine& o) {
        if (kind == 0) { kind = 1; a = o; }
        else if (kind == 1) { p = G1::from_affine(a).add_affine(o); kind = 2; }
        else p = p.add_affine(o);
    }
    G1 add(const G1& o) const {
        if (kind == 0) return o;
        if (kind == 1) return o.add_affine(a);
        return o.add(p);
    }
};
inline size_t get_at(size_t segment, size_t c, const uint8_t bytes[32]) {
    size_t skip_bits = segment * c, skip_bytes = skip_bits / 8;
    if (skip_bytes >= 32) return 0;
    uint8_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t i = 0; i < 8 && skip_bytes + i < 32; ++i) v[i] = bytes[skip_bytes + i];
    uint64_t tmp; memcpy(&tmp, v, 8);
    tmp >>= skip_bits - skip_bytes * 8;
    tmp %= (1ULL << c);
    return (size_t)tmp;
}
}  // namespace

void multiexp_serial(const Fr* coeffs, const G1Affine* bases, size_t n, G1& acc) {
    std::vector<std::array<uint8_t, 32>> reprs(n);
    for (size_t i = 0; i < n; ++i) { uint64_t t[4]; coeffs[i].to_raw(t); memcpy(reprs[i].data(), t, 32); }
    size_t c;
    if (n < 4) c = 1; else if (n < 32) c = 3; else c = (size_t)std::ceil(std::log((double)(uint32_t)n));
    size_t segments = 256 / c + 1;
    std::vector<Bucket> buckets;
    for (size_t seg = segments; seg-- > 0;) {
        for (size_t i = 0; i < c; ++i) acc = acc.dbl();
        buckets.assign(((size_t)1 << c) - 1, Bucket());
        for (size_t i = 0; i < n; ++i) {
            size_t d = get_at(seg, c, reprs[i].data());
            if (d != 0) buckets[d - 1].add_assign(bases[i]);
        }
        G1 running = G1::identity();
        for (size_t b = buckets.size(); b-- > 0;) {
            running = buckets[b].add(running);
            acc = acc.add(running);
        }
    }
}

G1 best_multiexp(const Fr* coeffs, const G1Affine* bases, size_t n) {
    size_t T = (size_t)num_threads();
    if (n > T) {
        size_t chunk = n / T;
        size_t num_chunks = (n + chunk - 1) / chunk;
        std::vector<G1> results(num_chunks, G1::identity());
        std::vector<std::thread> th;
        for (size_t i = 0; i < num_chunks; ++i) {
            size_t s = i * chunk, e = s + chunk < n ? s + chunk : n;
            th.emplace_back([=, &results] { multiexp_serial(coeffs + s, bases + s, e - s, results[i]); });
        }
        for (auto& t : th) t.join();
        G1 acc = G1::identity();
        for (auto& r : results) acc = acc.add(r);
        return acc;
    }
    G1 acc = G1::identity();
    multiexp_serial(coeffs, bases, n, acc);
    return acc;
}

// ------------------------------------------------------------------- FFT ----
namespace {
inline size_t bitreverse(size_t n, unsigned l) {
    size_t r = 0;
    for (unsigned i = 0; i < l; ++i) { r = (r << 1) | (n & 1); n >>= 1; }
    return r;
}
inline void butterfly_layer(Fr* left, Fr* right, size_t half, size_t twiddle_chunk, const Fr* tw) {
    Fr t = right[0]; right[0] = left[0]; left[0] += t; right[0] -= t;   // twiddle == 1
    for (size_t i = 1; i < half; ++i) {
        Fr t2 = right[i] * tw[i * twiddle_chunk];
        right[i] = left[i]; left[i] += t2; right[i] -= t2;
    }
}
// The combining layer of a node split over `threads` workers (disjoint index ranges, same arithmetic).
void butterfly_layer_par(Fr* left, Fr* right, size_t half, size_t twiddle_chunk, const Fr* tw, int threads) {
    if (threads <= 1 || half < 4096) { butterfly_layer(left, right, half, twiddle_chunk, tw); return; }
    auto range = [=](size_t s, size_t e) {
        for (size_t i = s; i < e; ++i) {
            if (i == 0) { Fr t = right[0]; right[0] = left[0]; left[0] += t; right[0] -= t; continue; }
            Fr t2 = right[i] * tw[i * twiddle_chunk];
            right[i] = left[i]; left[i] += t2; right[i] -= t2;
        }
    };
    std::vector<std::thread> th;
    size_t chunk = (half + threads - 1) / threads;
    for (int t = 1; t < threads; ++t) { size_t s = t * chunk, e = std::min(half, s + chunk); if (s < e) th.emplace_back(range, s, e); }
    range(0, std::min(half, chunk));
    for (auto& t : th) t.join();
}
// arithmetic::recursive_butterfly_arithmetic: rayon::join on the halves down to par_depth levels.  Upstream's tag runs
// the combining butterflies of a node on the joining thread; here the 2^par_depth worker budget is also spent on
// that layer (as later halo2 releases do), which changes the schedule and nothing else.
void recursive_butterfly(Fr* a, size_t n, size_t twiddle_chunk, const Fr* tw, int par_depth) {
    if (n == 2) { Fr t = a[1]; a[1] = a[0]; a[0] += t; a[1] -= t; return; }
    if (par_depth > 0) {                                // rayon::join
        std::thread other([=] { recursive_butterfly(a, n / 2, twiddle_chunk * 2, tw, par_depth - 1); });
        recursive_butterfly(a + n / 2, n / 2, twiddle_chunk * 2, tw, par_depth - 1);
        other.join();
        butterfly_layer_par(a, a + n / 2, n / 2, twiddle_chunk, tw, 1 << par_depth);
    } else {
        recursive_butterfly(a, n / 2, twiddle_chunk * 2, tw, 0);
        recursive_butterfly(a + n / 2, n / 2, twiddle_chunk * 2, tw, 0);
        butterfly_layer(a, a + n / 2, n / 2, twiddle_chunk, tw);
    }
}
}  // namespace

void best_fft(Fr* a, const Fr& omega, unsigned log_n) {
    size_t n = (size_t)1 << log_n;
    int threads = num_threads();
    unsigned log_threads = 0; while ((2 << log_threads) <= threads) ++log_threads;
    // bit-reversal permutation and twiddle table omega^i, i < n/2 (chunked over the workers: same values)
    parallelize(n, [&](size_t s, size_t e) { for (size_t k = s; k < e; ++k) { size_t rk = bitreverse(k, log_n); if (k < rk) std::swap(a[rk], a[k]); } });
    std::vector<Fr> tw(n / 2 ? n / 2 : 1);
    parallelize(n / 2, [&](size_t s, size_t e) { Fr w = omega.pow_u64(s); for (size_t i = s; i < e; ++i) { tw[i] = w; w *= omega; } });
    if (n / 2 == 0) tw[0] = Fr::one();
    if (log_n <= log_threads) {
        size_t chunk = 2, twiddle_chunk = n / 2;
        for (unsigned l = 0; l < log_n; ++l) {
            for (size_t s = 0; s < n; s += chunk) butterfly_layer(a + s, a + s + chunk / 2, chunk / 2, twiddle_chunk, tw.data());
            chunk *= 2; twiddle_chunk /= 2;
        }
    } else {
        recursive_butterfly(a, n, 1, tw.data(), (int)log_threads);
    }
}

Fr eval_polynomial(const Fr* poly, size_t n, const Fr& x) {
    Fr acc = Fr::zero();
    for (size_t i = n; i-- > 0;) acc = acc * x + poly[i];
    return acc;
}

void kate_division(const Fr* a, size_t n, const Fr& b, Fr* q) {
    Fr tmp = Fr::zero();
    for (size_t i = n - 1; i-- > 0;) { q[i] = a[i + 1] + tmp; tmp = q[i] * b; }
}

// ---------------------------------------------------------------- domain ----
Domain::Domain(unsigned j, unsigned k_) {
    init_fields();
    k = k_;
    quotient_poly_degree = (uint64_t)j - 1;
    n = 1ULL << k;
    extended_k = k;
    while ((1ULL << extended_k) < n * quotient_poly_degree) ++extended_k;
    extended_omega = FR_ROOT_OF_UNITY;
    for (unsigned i = extended_k; i < FR_S; ++i) extended_omega = extended_omega.sqr();
    omega = extended_omega;
    for (unsigned i = k; i < extended_k; ++i) omega = omega.sqr();
    omega_inv = omega.inv();
    extended_omega_inv = extended_omega.inv();
    g_coset = FR_ZETA;
    g_coset_inv = g_coset.sqr();
    ifft_divisor = Fr::from_u64(1ULL << k).inv();
    extended_ifft_divisor = Fr::from_u64(1ULL << extended_k).inv();
    barycentric_weight = Fr::from_u64(n).inv();
    Fr orig = FR_ZETA.pow_u64(n), step = extended_omega.pow_u64(n), cur = orig;
    do { t_evaluations.push_back(cur); cur *= step; } while (cur != orig);
    assert(t_evaluations.size() == ((size_t)1 << (extended_k - k)));
    for (auto& t : t_evaluations) t -= Fr::one();
    batch_invert(t_evaluations.data(), t_evaluations.size());
}

void Domain::distribute_powers_zeta(Fr* a, size_t len, bool into_coset) const {
    Fr p1 = into_coset ? g_coset : g_coset_inv, p2 = into_coset ? g_coset_inv : g_coset;
    parallelize(len, [&](size_t s, size_t e) {
        for (size_t i = s; i < e; ++i) {
            size_t m = i % 3;
            if (m == 1) a[i] *= p1; else if (m == 2) a[i] *= p2;
        }
    });
}

void Domain::lagrange_to_coeff(Fr* a) const {
    best_fft(a, omega_inv, k);
    parallelize(n, [&](size_t s, size_t e) { for (size_t i = s; i < e; ++i) a[i] *= ifft_divisor; });
}

void Domain::coeff_to_extended(const Fr* a, Fr* out) const {
    parallelize(extended_len(), [&](size_t s, size_t e) { for (size_t i = s; i < e; ++i) out[i] = i < n ? a[i] : Fr::zero(); });
    distribute_powers_zeta(out, n, true);
    best_fft(out, extended_omega, extended_k);
}

void Domain::extended_to_coeff(Fr* a) const {
    best_fft(a, extended_omega_inv, extended_k);
    size_t len = extended_len();
    parallelize(len, [&](size_t s, size_t e) { for (size_t i = s; i < e; ++i) a[i] *= extended_ifft_divisor; });
    distribute_powers_zeta(a, len, false);
}

void Domain::divide_by_vanishing_poly(Fr* a) const {
    size_t m = t_evaluations.size();
    parallelize(extended_len(), [&](size_t s, size_t e) { for (size_t i = s; i < e; ++i) a[i] *= t_evaluations[i % m]; });
}

Fr Domain::rotate_omega(const Fr& x, int rot) const {
    if (rot >= 0) return x * omega.pow_u64((uint64_t)rot);
    return x * omega_inv.pow_u64((uint64_t)(-(int64_t)rot));
}

}  // namespace orc
