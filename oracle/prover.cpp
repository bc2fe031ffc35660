// ORACLE — TEST INFRASTRUCTURE ONLY (see ff.hpp header; PARITY UNPINNED vs Rust).
//
// keygen_pk + create_proof of halo2_proofs (PSE tag v2023_02_02) for one circuit instance with
// KZG / SHPLONK / Blake2b — the calls `full_prover` makes at /root/reference/src/circuits/utils.rs:35
// and :38-49 — restated in threaded C++ so that the CPU arm of the benchmark and the full-size
// parity checks run the headline size (k = 20) in tens of seconds instead of minutes.
//
// It is the same restatement as oracle/prover.py, statement for statement (that file is the
// readable specification and carries the upstream file names per step); tests/test_oracle_prover.py
// checks that both produce identical proof bytes.  Upstream's structure is kept: every column is
// extended to the full 2^extended_k domain (coeff_to_extended), evaluate_h walks the extended rows
// in parallel chunks with a GraphEvaluator-style stack machine, commitments are best_multiexp
// (chunk per thread, c = ceil(ln n), unsigned windows), FFTs are best_fft (recursive butterflies).
// Parallelism = arith.hpp's parallelize (one chunk per hardware thread), like upstream's rayon pool.
#include "arith.hpp"
#include <algorithm>
#include <map>
#include <memory>
#include <set>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace orc {

// ------------------------------------------------------------------ BLAKE2b-512, personal "Halo2-Transcript"
struct Blake2b {
    uint64_t h[8], t = 0;
    uint8_t buf[128];
    size_t buflen = 0;
    static uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
    static const uint64_t* iv() {
        static const uint64_t IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                       0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
        return IV;
    }
    explicit Blake2b(const char personal[16]) {
        for (int i = 0; i < 8; ++i) h[i] = iv()[i];
        h[0] ^= 0x01010000ULL ^ 64;                       // digest length 64, fanout 1, depth 1, no key
        uint64_t p[2]; memcpy(p, personal, 16);
        h[6] ^= p[0]; h[7] ^= p[1];
    }
    void compress(const uint8_t* block, bool last) {
        static const uint8_t S[12][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
        uint64_t m[16], v[16];
        memcpy(m, block, 128);
        for (int i = 0; i < 8; ++i) { v[i] = h[i]; v[i + 8] = iv()[i]; }
        v[12] ^= t;                                        // messages are far below 2^64 bytes
        if (last) v[14] = ~v[14];
        auto G = [&](int a, int b, int c, int d, uint64_t x, uint64_t y) {
            v[a] = v[a] + v[b] + x; v[d] = rotr(v[d] ^ v[a], 32);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 24);
            v[a] = v[a] + v[b] + y; v[d] = rotr(v[d] ^ v[a], 16);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 63);
        };
        for (int r = 0; r < 12; ++r) {
            const uint8_t* s = S[r];
            G(0, 4, 8, 12, m[s[0]], m[s[1]]);   G(1, 5, 9, 13, m[s[2]], m[s[3]]);
            G(2, 6, 10, 14, m[s[4]], m[s[5]]);  G(3, 7, 11, 15, m[s[6]], m[s[7]]);
            G(0, 5, 10, 15, m[s[8]], m[s[9]]);  G(1, 6, 11, 12, m[s[10]], m[s[11]]);
            G(2, 7, 8, 13, m[s[12]], m[s[13]]); G(3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
        for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
    }
    void update(const uint8_t* data, size_t len) {
        while (len) {
            if (buflen == 128) { t += 128; compress(buf, false); buflen = 0; }
            size_t take = std::min(len, 128 - buflen);
            memcpy(buf + buflen, data, take);
            buflen += take; data += take; len -= take;
        }
    }
    void digest(uint8_t out[64]) const {                   // of a copy: the running state is kept (squeeze_challenge clones)
        Blake2b c = *this;
        c.t += c.buflen;
        memset(c.buf + c.buflen, 0, 128 - c.buflen);
        c.compress(c.buf, true);
        memcpy(out, c.h, 64);
    }
};

// transcript.rs: Blake2bWrite<_, G1Affine, Challenge255<_>>
struct Transcript {
    Blake2b st{"Halo2-Transcript"};
    std::vector<uint8_t> proof;
    void common_scalar(const Fr& s) {
        uint8_t b[33]; b[0] = 2; uint64_t r[4]; s.to_raw(r); memcpy(b + 1, r, 32); st.update(b, 33);
    }
    void common_point(const G1Affine& p) {
        uint8_t b[65]; b[0] = 1; uint64_t r[4];
        p.x.to_raw(r); memcpy(b + 1, r, 32); p.y.to_raw(r); memcpy(b + 33, r, 32);
        st.update(b, 65);
    }
    void write_point(const G1Affine& p) { common_point(p); uint8_t c[32]; p.to_bytes(c); proof.insert(proof.end(), c, c + 32); }
    void write_scalar(const Fr& s) { common_scalar(s); uint64_t r[4]; s.to_raw(r); proof.insert(proof.end(), (uint8_t*)r, (uint8_t*)r + 32); }
    Fr squeeze_challenge() {
        uint8_t z = 0; st.update(&z, 1);
        uint8_t d[64]; st.digest(d);
        uint64_t w[8]; memcpy(w, d, 64);
        return Fr::from_u512(w);
    }
};

// ------------------------------------------------------------------ constraint system (blob of circuit.py::to_blob)
enum : uint32_t { OP_CONST = 0, OP_FIXED = 1, OP_ADVICE = 2, OP_INSTANCE = 3, OP_NEG = 4, OP_ADD = 5, OP_MUL = 6, OP_SCALE = 7 };
struct Expr { uint32_t off, len; };
struct Cs {
    uint32_t k = 0, A = 0, F = 0, I = 0, bf = 0, degree = 0;
    std::vector<std::pair<int32_t, int32_t>> adv_q, fix_q, inst_q;
    std::vector<std::pair<uint32_t, uint32_t>> perm;
    std::vector<Expr> gates;
    struct Lk { std::vector<Expr> ins, tabs; };
    std::vector<Lk> lookups;
    std::vector<Fr> consts;
    std::vector<uint32_t> prog;
};
static bool parse_blob(const uint32_t* w, size_t nw, Cs& cs) {
    if (nw < 16 || w[0] != 0x324B5A42u || w[1] != 1) return false;
    cs.k = w[2]; cs.A = w[3]; cs.F = w[4]; cs.I = w[5];
    uint32_t naq = w[6], nfq = w[7], niq = w[8], ng = w[9], nl = w[10], np = w[11], nc = w[12], nprog = w[13];
    cs.bf = w[14]; cs.degree = w[15];
    size_t p = 16;
    auto rdq = [&](std::vector<std::pair<int32_t, int32_t>>& q, uint32_t cnt) { for (uint32_t i = 0; i < cnt; ++i) { q.push_back({(int32_t)w[p], (int32_t)w[p + 1]}); p += 2; } };
    rdq(cs.adv_q, naq); rdq(cs.fix_q, nfq); rdq(cs.inst_q, niq);
    for (uint32_t i = 0; i < np; ++i) { cs.perm.push_back({w[p], w[p + 1]}); p += 2; }
    for (uint32_t i = 0; i < ng; ++i) { cs.gates.push_back({w[p], w[p + 1]}); p += 2; }
    for (uint32_t i = 0; i < nl; ++i) {
        uint32_t m = w[p++];
        Cs::Lk lk;
        for (uint32_t j = 0; j < m; ++j) { lk.ins.push_back({w[p], w[p + 1]}); p += 2; }
        for (uint32_t j = 0; j < m; ++j) { lk.tabs.push_back({w[p], w[p + 1]}); p += 2; }
        cs.lookups.push_back(lk);
    }
    for (uint32_t i = 0; i < nc; ++i) {
        uint64_t c[4];
        for (int j = 0; j < 4; ++j) c[j] = (uint64_t)w[p + 2 * j] | ((uint64_t)w[p + 2 * j + 1] << 32);
        cs.consts.push_back(Fr::from_raw(c));
        p += 8;
    }
    if (p + nprog > nw) return false;
    cs.prog.assign(w + p, w + p + nprog);
    return true;
}

typedef std::vector<Fr> Poly;

// plonk::evaluation: one expression on one row.  Rotations wrap in the whole array of `size` rows
// (np.roll in prover.py): row (idx + rot * rot_scale) mod size.
struct Cols { const std::vector<Poly>* fixed; const std::vector<Poly>* advice; const std::vector<Poly>* instance; };
static inline Fr eval_row(const Cs& cs, Expr e, const Cols& c, size_t idx, size_t size, int64_t rot_scale) {
    Fr st[32];
    int sp = 0;
    auto load = [&](const std::vector<Poly>& cols, const std::pair<int32_t, int32_t>& q) -> const Fr& {
        int64_t r = ((int64_t)idx + (int64_t)q.second * rot_scale) % (int64_t)size;
        if (r < 0) r += (int64_t)size;
        return cols[q.first][(size_t)r];
    };
    for (uint32_t i = e.off; i < e.off + e.len; ++i) {
        uint32_t op = cs.prog[i] & 0xff, arg = cs.prog[i] >> 8;
        switch (op) {
            case OP_CONST: st[sp++] = cs.consts[arg]; break;
            case OP_FIXED: st[sp++] = load(*c.fixed, cs.fix_q[arg]); break;
            case OP_ADVICE: st[sp++] = load(*c.advice, cs.adv_q[arg]); break;
            case OP_INSTANCE: st[sp++] = load(*c.instance, cs.inst_q[arg]); break;
            case OP_NEG: st[sp - 1] = st[sp - 1].neg(); break;
            case OP_ADD: st[sp - 2] = st[sp - 2] + st[sp - 1]; --sp; break;
            case OP_MUL: st[sp - 2] = st[sp - 2] * st[sp - 1]; --sp; break;
            case OP_SCALE: st[sp - 1] = st[sp - 1] * cs.consts[arg]; break;
        }
    }
    return st[0];
}

// ------------------------------------------------------------------ proving key (plonk/keygen.rs keygen_pk)
struct Pk {
    Cs cs;
    std::unique_ptr<Domain> dom;
    size_t n = 0, ext_n = 0;
    std::vector<Poly> fixed_values, fixed_polys, fixed_cosets, perm_values, perm_polys, perm_cosets;
    Poly l0, l_last, l_active;
};

static Poly to_coeff(const Domain& d, const Poly& v) { Poly p = v; d.lagrange_to_coeff(p.data()); return p; }
static Poly to_ext(const Domain& d, const Poly& p) { Poly e(d.extended_len()); d.coeff_to_extended(p.data(), e.data()); return e; }

static Fr eval_poly_par(const Poly& poly, const Fr& x) {       // arithmetic::eval_polynomial (chunks weighted by x^start)
    size_t n = poly.size(), T = (size_t)num_threads();
    if (n * 2 < T || T == 1) return eval_polynomial(poly.data(), n, x);
    size_t chunk = (n + T - 1) / T;
    std::vector<Fr> parts((n + chunk - 1) / chunk, Fr::zero());
    std::vector<std::thread> th;
    for (size_t i = 0; i < parts.size(); ++i)
        th.emplace_back([&, i] {
            size_t s = i * chunk, e = std::min(n, s + chunk);
            parts[i] = eval_polynomial(poly.data() + s, e - s, x) * x.pow_u64(s);
        });
    for (auto& t : th) t.join();
    Fr acc = Fr::zero();
    for (auto& p : parts) acc += p;
    return acc;
}
static void batch_invert_par(Poly& v) {                        // parallelize(.., |chunk| chunk.batch_invert())
    parallelize(v.size(), [&](size_t s, size_t e) { batch_invert(v.data() + s, e - s); });
}
static G1Affine commit(const G1Affine* bases, const Poly& poly) { return best_multiexp(poly.data(), bases, poly.size()).to_affine(); }

// arithmetic::lagrange_interpolate
static std::vector<Fr> lagrange_interpolate(const std::vector<Fr>& pts, const std::vector<Fr>& evals) {
    size_t m = pts.size();
    std::vector<Fr> poly(m, Fr::zero());
    for (size_t j = 0; j < m; ++j) {
        std::vector<Fr> num(1, Fr::one());
        Fr den = Fr::one();
        for (size_t kx = 0; kx < m; ++kx) {
            if (kx == j) continue;
            std::vector<Fr> nw(num.size() + 1, Fr::zero());
            for (size_t i = 0; i < num.size(); ++i) { nw[i] -= num[i] * pts[kx]; nw[i + 1] += num[i]; }
            num.swap(nw);
            den *= pts[j] - pts[kx];
        }
        Fr s = evals[j] * den.inv();
        for (size_t i = 0; i < num.size(); ++i) poly[i] += num[i] * s;
    }
    return poly;
}

struct FrLess { bool operator()(const Fr& a, const Fr& b) const { return Fr::less(a, b); } };

// lookup::prover::permute_expression_pair (capi.cpp)
extern "C" int orc_permute_expression_pair(const uint64_t* input, const uint64_t* table, size_t usable, uint64_t* a_out, uint64_t* s_out);
extern "C" size_t orc_pk_rng_draws(const void* pk);

struct Rng {                                                    // pre-drawn Fr::random inputs, consumed in upstream order
    const uint64_t* wide; size_t pos = 0, count;
    Fr one() { return Fr::from_u512(wide + 8 * pos++); }
    void take(Fr* out, size_t c) { for (size_t i = 0; i < c; ++i) out[i] = one(); }
};

static int create_proof(const Pk& pk, const G1Affine* g, const G1Affine* gl, const uint64_t* const* advice_in, const uint64_t* const* instances,
                        const uint32_t* inst_lens, const uint64_t* rng_wide, size_t rng_count, const Fr& transcript_repr, std::vector<uint8_t>& out) {
    const Cs& cs = pk.cs;
    const Domain& dom = *pk.dom;
    const size_t n = pk.n, ext_n = pk.ext_n, bf = cs.bf, usable = n - (bf + 1);
    const int64_t rot_scale = (int64_t)1 << (dom.extended_k - cs.k);
    Rng rng{rng_wide, 0, rng_count};
    if (rng_count < orc_pk_rng_draws(&pk)) return 3;
    const bool trace = getenv("ORC_TRACE") != nullptr;           // phase wall clock on stderr
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!trace) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[orc prover] %-28s %8.3f s\n", what, std::chrono::duration<double>(now - t_last).count());
        t_last = now;
    };
    Transcript tr;
    tr.common_scalar(transcript_repr);
    // -- instances
    std::vector<Poly> inst_values(cs.I, Poly(n, Fr::zero()));
    for (uint32_t c = 0; c < cs.I; ++c) {
        if (inst_lens[c] > usable) return 2;                    // InstanceTooLarge
        for (uint32_t i = 0; i < inst_lens[c]; ++i) { Fr v; memcpy(v.l, instances[c] + 4 * i, 32); tr.common_scalar(v); inst_values[c][i] = v; }
    }
    std::vector<Poly> inst_polys;
    for (auto& v : inst_values) inst_polys.push_back(to_coeff(dom, v));
    // -- advice: blinding rows, blinds, commitments
    std::vector<Poly> advice(cs.A, Poly(n));
    for (uint32_t c = 0; c < cs.A; ++c) memcpy(advice[c].data(), advice_in[c], n * 32);
    for (auto& a : advice) rng.take(a.data() + usable, bf + 1);
    for (uint32_t c = 0; c < cs.A; ++c) rng.one();
    for (auto& a : advice) tr.write_point(commit(gl, a));
    mark("advice commits");
    const Fr theta = tr.squeeze_challenge();
    // -- lookups: commit_permuted
    struct Lookup { Poly cin, ctab, pin, ptab, pin_poly, ptab_poly, z_poly; };
    std::vector<Lookup> lookups(cs.lookups.size());
    const Cols lag{&pk.fixed_values, &advice, &inst_values};
    for (size_t l = 0; l < cs.lookups.size(); ++l) {
        Lookup& lk = lookups[l];
        auto compress = [&](const std::vector<Expr>& exprs, Poly& acc) {
            acc.assign(n, Fr::zero());
            parallelize(n, [&](size_t s, size_t e) {
                for (size_t i = s; i < e; ++i) { Fr a = Fr::zero(); for (auto& ex : exprs) a = a * theta + eval_row(cs, ex, lag, i, n, 1); acc[i] = a; }
            });
        };
        compress(cs.lookups[l].ins, lk.cin); compress(cs.lookups[l].tabs, lk.ctab);
        lk.pin.assign(n, Fr::zero()); lk.ptab.assign(n, Fr::zero());
        if (orc_permute_expression_pair((const uint64_t*)lk.cin.data(), (const uint64_t*)lk.ctab.data(), usable, (uint64_t*)lk.pin.data(), (uint64_t*)lk.ptab.data())) return 1;
        rng.take(lk.pin.data() + usable, bf + 1);
        rng.take(lk.ptab.data() + usable, bf + 1);
        lk.pin_poly = to_coeff(dom, lk.pin); rng.one();
        G1Affine c_in = commit(gl, lk.pin);
        lk.ptab_poly = to_coeff(dom, lk.ptab); rng.one();
        G1Affine c_tab = commit(gl, lk.ptab);
        tr.write_point(c_in); tr.write_point(c_tab);
    }
    mark("lookups commit_permuted");
    const Fr beta = tr.squeeze_challenge(), gamma = tr.squeeze_challenge();
    // -- permutation commit
    auto perm_values = [&](uint32_t ct, uint32_t ci) -> const Poly& { return ct == 0 ? advice[ci] : ct == 1 ? pk.fixed_values[ci] : inst_values[ci]; };
    struct Set { Poly poly, coset; };
    std::vector<Set> sets;
    const size_t chunk = cs.degree - 2, P = cs.perm.size();
    Poly omega_pows;
    if (P) { omega_pows.resize(n); Fr c = Fr::one(); for (size_t i = 0; i < n; ++i) { omega_pows[i] = c; c *= dom.omega; } }
    Fr last_z = Fr::one(), deltaomega = Fr::one();
    for (size_t s0 = 0; s0 < P; s0 += chunk) {
        size_t s1 = std::min(P, s0 + chunk);
        Poly mod(n);
        parallelize(n, [&](size_t s, size_t e) {
            for (size_t i = s; i < e; ++i) {
                Fr m = Fr::one();
                for (size_t j = s0; j < s1; ++j) m *= pk.perm_values[j][i] * beta + gamma + perm_values(cs.perm[j].first, cs.perm[j].second)[i];
                mod[i] = m;
            }
        });
        batch_invert_par(mod);
        std::vector<Fr> coef;
        for (size_t j = s0; j < s1; ++j) { coef.push_back(deltaomega * beta); deltaomega *= FR_DELTA; }
        parallelize(n, [&](size_t s, size_t e) {
            for (size_t i = s; i < e; ++i) {
                Fr m = mod[i];
                for (size_t j = s0; j < s1; ++j) m *= omega_pows[i] * coef[j - s0] + gamma + perm_values(cs.perm[j].first, cs.perm[j].second)[i];
                mod[i] = m;
            }
        });
        Poly z(n);
        { Fr run = last_z; for (size_t i = 0; i < n; ++i) { z[i] = run; run *= mod[i]; } }
        rng.take(z.data() + (n - bf), bf);
        last_z = z[n - (bf + 1)];
        rng.one();
        tr.write_point(commit(gl, z));
        Set st;
        st.poly = to_coeff(dom, z);
        st.coset = to_ext(dom, st.poly);
        sets.push_back(std::move(st));
    }
    mark("permutation commit");
    // -- lookups: commit_product
    for (auto& lk : lookups) {
        Poly den(n);
        parallelize(n, [&](size_t s, size_t e) { for (size_t i = s; i < e; ++i) den[i] = (lk.pin[i] + beta) * (lk.ptab[i] + gamma); });
        batch_invert_par(den);
        parallelize(n, [&](size_t s, size_t e) { for (size_t i = s; i < e; ++i) den[i] = den[i] * (lk.cin[i] + beta) * (lk.ctab[i] + gamma); });
        Poly z(n);
        { Fr run = Fr::one(); for (size_t i = 0; i < n; ++i) { z[i] = run; run *= den[i]; } }
        rng.take(z.data() + (n - bf), bf);
        rng.one();
        tr.write_point(commit(gl, z));
        lk.z_poly = to_coeff(dom, z);
    }
    mark("lookups commit_product");
    // -- vanishing commit
    Poly random_poly(n);
    rng.take(random_poly.data(), n); rng.one();
    tr.write_point(commit(g, random_poly));
    const Fr y = tr.squeeze_challenge();
    // -- advice polys and extended cosets
    std::vector<Poly> advice_polys, adv_cos, inst_cos;
    for (auto& a : advice) advice_polys.push_back(to_coeff(dom, a));
    for (auto& p : advice_polys) adv_cos.push_back(to_ext(dom, p));
    for (auto& p : inst_polys) inst_cos.push_back(to_ext(dom, p));
    mark("random commit + advice cosets");
    // -- evaluate_h
    Poly h(ext_n, Fr::zero());
    const Cols ext{&pk.fixed_cosets, &adv_cos, &inst_cos};
    auto rot = [&](size_t idx, int64_t r) { int64_t v = ((int64_t)idx + r * rot_scale) % (int64_t)ext_n; if (v < 0) v += (int64_t)ext_n; return (size_t)v; };
    parallelize(ext_n, [&](size_t s, size_t e) {
        for (size_t i = s; i < e; ++i) { Fr acc = Fr::zero(); for (auto& gt : cs.gates) acc = acc * y + eval_row(cs, gt, ext, i, ext_n, rot_scale); h[i] = acc; }
    });
    mark("evaluate_h gates");
    if (!sets.empty()) {
        const int64_t last_rot = -(int64_t)(bf + 1);
        auto cos_col = [&](uint32_t ct, uint32_t ci) -> const Poly& { return ct == 0 ? adv_cos[ci] : ct == 1 ? pk.fixed_cosets[ci] : inst_cos[ci]; };
        Poly ext_pows(ext_n);                                      // extended_omega^idx
        parallelize(ext_n, [&](size_t s, size_t e) { Fr c = dom.extended_omega.pow_u64(s); for (size_t i = s; i < e; ++i) { ext_pows[i] = c; c *= dom.extended_omega; } });
        std::vector<Fr> cdelta;                                    // beta * zeta * delta^j
        { Fr d = Fr::one(); for (size_t j = 0; j < P; ++j) { cdelta.push_back(beta * FR_ZETA * d); d *= FR_DELTA; } }
        parallelize(ext_n, [&](size_t s, size_t e) {
            for (size_t i = s; i < e; ++i) {
                Fr acc = h[i];
                const Fr first = sets.front().coset[i], last = sets.back().coset[i];
                acc = acc * y + (Fr::one() - first) * pk.l0[i];
                acc = acc * y + (last * last - last) * pk.l_last[i];
                for (size_t si = 1; si < sets.size(); ++si) acc = acc * y + (sets[si].coset[i] - sets[si - 1].coset[rot(i, last_rot)]) * pk.l0[i];
                for (size_t si = 0, s0 = 0; s0 < P; ++si, s0 += chunk) {
                    size_t s1 = std::min(P, s0 + chunk);
                    Fr left = sets[si].coset[rot(i, 1)], right = sets[si].coset[i];
                    for (size_t j = s0; j < s1; ++j) {
                        const Fr& v = cos_col(cs.perm[j].first, cs.perm[j].second)[i];
                        left *= v + pk.perm_cosets[j][i] * beta + gamma;
                        right *= v + ext_pows[i] * cdelta[j] + gamma;
                    }
                    acc = acc * y + (left - right) * pk.l_active[i];
                }
                h[i] = acc;
            }
        });
    }
    mark("evaluate_h permutation");
    for (size_t l = 0; l < lookups.size(); ++l) {
        Lookup& lk = lookups[l];
        Poly zc = to_ext(dom, lk.z_poly), ac = to_ext(dom, lk.pin_poly), sc = to_ext(dom, lk.ptab_poly);
        parallelize(ext_n, [&](size_t s, size_t e) {
            for (size_t i = s; i < e; ++i) {
                Fr cin = Fr::zero(), ctab = Fr::zero();
                for (auto& ex : cs.lookups[l].ins) cin = cin * theta + eval_row(cs, ex, ext, i, ext_n, rot_scale);
                for (auto& ex : cs.lookups[l].tabs) ctab = ctab * theta + eval_row(cs, ex, ext, i, ext_n, rot_scale);
                const Fr table_value = (cin + beta) * (ctab + gamma);
                const Fr z = zc[i], zn = zc[rot(i, 1)], a = ac[i], ap = ac[rot(i, -1)], sv = sc[i], ams = a - sv;
                Fr acc = h[i];
                acc = acc * y + (Fr::one() - z) * pk.l0[i];
                acc = acc * y + (z * z - z) * pk.l_last[i];
                acc = acc * y + (zn * (a + beta) * (sv + gamma) - z * table_value) * pk.l_active[i];
                acc = acc * y + ams * pk.l0[i];
                acc = acc * y + ams * (a - ap) * pk.l_active[i];
                h[i] = acc;
            }
        });
    }
    mark("evaluate_h lookups");
    // -- vanishing construct
    dom.divide_by_vanishing_poly(h.data());
    dom.extended_to_coeff(h.data());
    const size_t q = dom.quotient_poly_degree;
    std::vector<Poly> h_pieces;
    for (size_t i = 0; i < q; ++i) h_pieces.emplace_back(h.begin() + i * n, h.begin() + (i + 1) * n);
    { Poly().swap(h); }
    for (size_t i = 0; i < q; ++i) rng.one();
    for (auto& hp : h_pieces) tr.write_point(commit(g, hp));
    if (rng.pos > rng.count) return 3;                          // stream shorter than orc_pk_rng_draws
    mark("vanishing construct + h commits");
    const Fr x = tr.squeeze_challenge();
    const Fr xn = x.pow_u64(n);
    // -- evaluations
    for (auto& qy : cs.adv_q) tr.write_scalar(eval_poly_par(advice_polys[qy.first], dom.rotate_omega(x, qy.second)));
    for (auto& qy : cs.fix_q) tr.write_scalar(eval_poly_par(pk.fixed_polys[qy.first], dom.rotate_omega(x, qy.second)));
    Poly h_poly(n, Fr::zero());
    for (size_t p = q; p-- > 0;) parallelize(n, [&](size_t s, size_t e) { for (size_t i = s; i < e; ++i) h_poly[i] = h_poly[i] * xn + h_pieces[p][i]; });
    tr.write_scalar(eval_poly_par(random_poly, x));
    for (auto& sp : pk.perm_polys) tr.write_scalar(eval_poly_par(sp, x));
    const Fr x_next = dom.rotate_omega(x, 1), x_prev = dom.rotate_omega(x, -1), x_last = dom.rotate_omega(x, -(int)(bf + 1));
    for (size_t si = 0; si < sets.size(); ++si) {
        tr.write_scalar(eval_poly_par(sets[si].poly, x));
        tr.write_scalar(eval_poly_par(sets[si].poly, x_next));
        if (si + 1 < sets.size()) tr.write_scalar(eval_poly_par(sets[si].poly, x_last));
    }
    for (auto& lk : lookups) {
        tr.write_scalar(eval_poly_par(lk.z_poly, x));
        tr.write_scalar(eval_poly_par(lk.z_poly, x_next));
        tr.write_scalar(eval_poly_par(lk.pin_poly, x));
        tr.write_scalar(eval_poly_par(lk.pin_poly, x_prev));
        tr.write_scalar(eval_poly_par(lk.ptab_poly, x));
    }
    // -- multiopen queries (poly identity = pointer)
    struct Query { const Poly* poly; Fr point; };
    std::vector<Query> queries;
    for (auto& qy : cs.adv_q) queries.push_back({&advice_polys[qy.first], dom.rotate_omega(x, qy.second)});
    for (auto& st : sets) { queries.push_back({&st.poly, x}); queries.push_back({&st.poly, x_next}); }
    for (size_t si = sets.size(); si-- > 0;) if (si + 1 < sets.size()) queries.push_back({&sets[si].poly, x_last});
    for (auto& lk : lookups) {
        queries.push_back({&lk.z_poly, x}); queries.push_back({&lk.pin_poly, x}); queries.push_back({&lk.ptab_poly, x});
        queries.push_back({&lk.pin_poly, x_prev}); queries.push_back({&lk.z_poly, x_next});
    }
    for (auto& qy : cs.fix_q) queries.push_back({&pk.fixed_polys[qy.first], dom.rotate_omega(x, qy.second)});
    for (auto& sp : pk.perm_polys) queries.push_back({&sp, x});
    queries.push_back({&h_poly, x}); queries.push_back({&random_poly, x});
    mark("evaluations");
    // -- ProverSHPLONK::create_proof
    const Fr sy = tr.squeeze_challenge();
    struct CommSet { const Poly* poly; std::set<Fr, FrLess> pts; };
    std::vector<CommSet> comm_sets;
    std::set<Fr, FrLess> super_points;
    for (auto& qy : queries) {
        super_points.insert(qy.point);
        auto it = std::find_if(comm_sets.begin(), comm_sets.end(), [&](const CommSet& c) { return c.poly == qy.poly; });
        if (it == comm_sets.end()) { comm_sets.push_back({qy.poly, {}}); it = comm_sets.end() - 1; }
        it->pts.insert(qy.point);
    }
    struct RotSet { std::vector<Fr> pts; std::vector<const Poly*> polys; std::vector<std::vector<Fr>> low; };
    std::vector<RotSet> rot_sets;
    for (auto& c : comm_sets) {
        std::vector<Fr> pts(c.pts.begin(), c.pts.end());
        auto it = std::find_if(rot_sets.begin(), rot_sets.end(), [&](const RotSet& r) { return r.pts == pts; });
        if (it == rot_sets.end()) { rot_sets.push_back({pts, {}, {}}); it = rot_sets.end() - 1; }
        if (std::find(it->polys.begin(), it->polys.end(), c.poly) == it->polys.end()) it->polys.push_back(c.poly);
    }
    for (auto& rs : rot_sets)
        for (const Poly* p : rs.polys) {
            std::vector<Fr> evals;
            for (auto& pt : rs.pts) evals.push_back(eval_poly_par(*p, pt));
            rs.low.push_back(lagrange_interpolate(rs.pts, evals));
        }
    const Fr sv = tr.squeeze_challenge();
    Poly h_x(n, Fr::zero());
    {
        Fr v_pow = Fr::one();
        for (auto& rs : rot_sets) {
            Poly n_x(n, Fr::zero());
            Fr y_pow = Fr::one();
            for (size_t j = 0; j < rs.polys.size(); ++j) {
                const Poly& p = *rs.polys[j];
                const std::vector<Fr>& low = rs.low[j];
                parallelize(n, [&](size_t s, size_t e) {
                    for (size_t i = s; i < e; ++i) { Fr c = p[i]; if (i < low.size()) c -= low[i]; n_x[i] += c * y_pow; }
                });
                y_pow *= sy;
            }
            size_t len = n;
            Poly tmp(n);
            for (auto& root : rs.pts) { kate_division(n_x.data(), len, root, tmp.data()); --len; std::copy(tmp.begin(), tmp.begin() + len, n_x.begin()); }
            parallelize(len, [&](size_t s, size_t e) { for (size_t i = s; i < e; ++i) h_x[i] += n_x[i] * v_pow; });
            v_pow *= sv;
        }
    }
    tr.write_point(commit(g, h_x));
    const Fr su = tr.squeeze_challenge();
    {
        auto vanish = [&](const std::vector<Fr>& roots) { Fr acc = Fr::one(); for (auto& r : roots) acc *= su - r; return acc; };
        Poly l_x(n, Fr::zero());
        Fr v_pow = Fr::one(), z0 = Fr::zero();
        for (size_t si = 0; si < rot_sets.size(); ++si) {
            RotSet& rs = rot_sets[si];
            std::vector<Fr> diffs;
            for (auto& p : super_points) if (std::find(rs.pts.begin(), rs.pts.end(), p) == rs.pts.end()) diffs.push_back(p);
            const Fr z_i = vanish(diffs);
            if (si == 0) z0 = z_i;
            Fr y_pow = Fr::one();
            for (size_t j = 0; j < rs.polys.size(); ++j) {
                const Poly& p = *rs.polys[j];
                Fr r_eval = Fr::zero();
                for (size_t c = rs.low[j].size(); c-- > 0;) r_eval = r_eval * su + rs.low[j][c];
                const Fr f = y_pow * z_i * v_pow;
                parallelize(n, [&](size_t s, size_t e) { for (size_t i = s; i < e; ++i) { Fr c = p[i]; if (i == 0) c -= r_eval; l_x[i] += c * f; } });
                y_pow *= sy;
            }
            v_pow *= sv;
        }
        const Fr zt = vanish(std::vector<Fr>(super_points.begin(), super_points.end()));
        parallelize(n, [&](size_t s, size_t e) { for (size_t i = s; i < e; ++i) l_x[i] -= h_x[i] * zt; });
        Poly h2(n - 1);
        kate_division(l_x.data(), n, su, h2.data());
        const Fr zi = z0.inv();
        parallelize(n - 1, [&](size_t s, size_t e) { for (size_t i = s; i < e; ++i) h2[i] *= zi; });
        tr.write_point(commit(g, h2));
    }
    mark("shplonk");
    out = tr.proof;
    return 0;
}

}  // namespace orc

using namespace orc;

extern "C" {

// keygen_pk: fixed = F columns of n Lagrange values (Montgomery), map_col / map_row = the permutation Assembly's
// mapping (P x n).  Returns an opaque handle (null on a malformed blob).
void* orc_pk_create(const uint32_t* blob, size_t nwords, const uint64_t* const* fixed, const uint32_t* map_col, const uint32_t* map_row) {
    init_fields();
    std::unique_ptr<Pk> pk(new Pk());
    if (!parse_blob(blob, nwords, pk->cs)) return nullptr;
    const Cs& cs = pk->cs;
    pk->dom.reset(new Domain(cs.degree, cs.k));
    const Domain& dom = *pk->dom;
    const size_t n = pk->n = (size_t)1 << cs.k;
    pk->ext_n = dom.extended_len();
    for (uint32_t c = 0; c < cs.F; ++c) {
        Poly v(n); memcpy(v.data(), fixed[c], n * 32);
        pk->fixed_values.push_back(v);
        pk->fixed_polys.push_back(to_coeff(dom, v));
        pk->fixed_cosets.push_back(to_ext(dom, pk->fixed_polys.back()));
    }
    const size_t P = cs.perm.size();
    if (P) {
        Poly om(n), dl(P);
        { Fr c = Fr::one(); for (size_t i = 0; i < n; ++i) { om[i] = c; c *= dom.omega; } }
        { Fr c = Fr::one(); for (size_t i = 0; i < P; ++i) { dl[i] = c; c *= FR_DELTA; } }
        for (size_t c = 0; c < P; ++c) {
            Poly v(n);
            for (size_t i = 0; i < n; ++i) v[i] = dl[map_col[c * n + i]] * om[map_row[c * n + i]];
            pk->perm_values.push_back(v);
            pk->perm_polys.push_back(to_coeff(dom, v));
            pk->perm_cosets.push_back(to_ext(dom, pk->perm_polys.back()));
        }
    }
    auto indicator = [&](size_t lo, size_t hi) { Poly v(n, Fr::zero()); for (size_t i = lo; i < hi; ++i) v[i] = Fr::one(); return to_ext(dom, to_coeff(dom, v)); };
    pk->l0 = indicator(0, 1);
    pk->l_last = indicator(n - cs.bf - 1, n - cs.bf);
    Poly lb = indicator(n - cs.bf, n);
    pk->l_active.resize(pk->ext_n);
    for (size_t i = 0; i < pk->ext_n; ++i) pk->l_active[i] = Fr::one() - (pk->l_last[i] + lb[i]);
    return pk.release();
}
void orc_pk_destroy(void* pk) { delete (Pk*)pk; }

size_t orc_pk_rng_draws(const void* pk_) {
    const Pk* pk = (const Pk*)pk_;
    const Cs& cs = pk->cs;
    size_t bf = cs.bf, L = cs.lookups.size(), chunk = cs.degree - 2, S = (cs.perm.size() + chunk - 1) / chunk;
    return cs.A * (bf + 1) + cs.A + L * (2 * (bf + 1) + 2) + S * (bf + 1) + L * (bf + 1) + pk->n + 1 + pk->dom->quotient_poly_degree;
}

// 0 ok; 1 ConstraintSystemFailure (lookup input not in table); 2 InstanceTooLarge; 3 rng draw count; 4 buffer too small
int orc_create_proof(const void* pk, const uint64_t* g, const uint64_t* g_lagrange, const uint64_t* const* advice, const uint64_t* const* instances,
                     const uint32_t* inst_lens, const uint64_t* rng_wide, size_t rng_count, const uint64_t* transcript_repr,
                     uint8_t* proof_out, size_t cap, size_t* proof_len) {
    init_fields();
    std::vector<uint8_t> proof;
    Fr repr; memcpy(repr.l, transcript_repr, 32);
    int rc = create_proof(*(const Pk*)pk, (const G1Affine*)g, (const G1Affine*)g_lagrange, advice, instances, inst_lens, rng_wide, rng_count, repr, proof);
    if (rc) return rc;
    *proof_len = proof.size();
    if (proof.size() > cap) return 4;
    memcpy(proof_out, proof.data(), proof.size());
    return 0;
}

}  // extern "C"
