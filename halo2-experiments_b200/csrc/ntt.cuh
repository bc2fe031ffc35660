// Radix-2 NTT over bn256 Fr, natural order in -> natural order out, replacing
// halo2_proofs `arithmetic::best_fft` (PSE v2023_02_02 src/arithmetic.rs; reached
// from keygen_pk / create_proof at /root/reference/src/circuits/utils.rs:35,40-48).
//
// A transform of size N = M_1 * M_2 * ... * M_P (P <= 4, each M_p <= 2^10) runs as
// P passes over HBM.  View the index as mixed-radix digits n = (n_1, ..., n_P),
// n_1 slowest.  Pass p < P is a batch of M_p-point DFTs along axis p (stride
// L_p = M_{p+1}...M_P) done in shared memory on a tile of TW adjacent columns (so
// every global access is a TW*32-byte contiguous run), followed by the inter-pass
// twiddle w_{M_p L_p}^{l k_p}.  The last pass transforms the contiguous axis of TW
// rows whose slowest digit k_1 is adjacent and stores digit-reversed, which makes
// its writes TW*32-byte runs as well and the final order natural without a
// separate bit-reversal or transpose pass.  Inside a tile the DFT is
// decimation-in-frequency (natural in, bit-reversed out); the bit reversal is
// folded into the shared-memory read of the store phase.
//
// Pre/post hooks fuse the work EvaluationDomain does around best_fft
// (poly/domain.rs): zero-padding + zeta^(i mod 3) coset distribution on load
// (coeff_to_extended), and the 1/n resp. 1/ext * zeta^-(i mod 3) scaling on store
// (lagrange_to_coeff / extended_to_coeff).
#pragma once
#include "field.cuh"
#include "blockexec.cuh"

namespace b200zk {

struct alignas(16) half_t { uint32_t v[4]; };

struct alignas(64) fe2_t { fe_t w, wq; };        // a constant and its precomputed quotient, one 64-byte record

struct NttPassArgs {
    const fe_t* in;
    fe_t* out;
    uint32_t log_n;         // full transform size
    uint32_t log_m;         // this pass: DFT length M
    uint32_t log_l;         // product of the faster axes (stride of axis m); 0 for the last pass
    uint32_t log_tw;        // tile width
    uint32_t is_last;
    // last pass only: digits of the row index.  log_m1 = slowest digit, log_mid = the rest.
    uint32_t log_m1, log_mid;
    uint32_t log_m3;        // four-pass plans of the warp-level kernel: the third digit's size (faster half of the middle bits), else 0
    uint32_t n_in;          // first pass: elements beyond n_in read as zero (n_in = N otherwise)
    uint32_t use_pre, use_post;
    fe_t pre[3], post[3];   // multiplied by index mod 3 on load (first pass) / store (last pass)
    const fe_t* pre_tab;    // optional, first pass: element g is multiplied by pre_tab[g] on load (coset powers)
    uint32_t in_mask;       // first pass of a batch that shares ONE input: read in[g & in_mask] (0 = read in[g])
    const fe_t* roots;      // w_R^j, j < R/2, R = 2^log_roots >= M  (w_R = omega^(N/R))
    uint32_t log_roots;
    const fe_t* tw_lo;      // omega^i,            i < 2^tw_lo_bits
    const fe_t* tw_hi;      // omega^(j << lo_bits), j < N >> tw_lo_bits
    const fe_t* tw_full;    // optional: omega^E for every E < N (saves the lo*hi multiplication)
    uint32_t tw_lo_bits;
    uint32_t tw_shift;      // inter-pass twiddle exponent = ((l + l_offset) * k) << tw_shift
    uint32_t l_offset;      // global column index of local column 0 (sharded four-step column step)
    uint32_t batch_tiles;   // last pass of a batch of independent transforms: tiles per transform (0 = single)
    // Sharded four-step column step fused with its all-to-all: output row k goes straight into the
    // memory of the rank that owns it (peer pointers over NVLink / NVSwitch, CUDA IPC), at
    // [k mod rows_per_rank][global column]; no staging buffer, no separate exchange.
    uint32_t scatter, log_rows_per_rank, log_c_total;
    uint32_t canonical_out; // warp-level kernel: this non-last pass is the end of a call (column step) — reduce its [0, 2p) values to [0, p)
    fe_t* peers[8];
    // Constant-operand (Shoup) form of the two twiddle tables — {plain value, floor(value * 2^256 / r)} records, see
    // Field::mul_shoup — used by the warp-level kernel when non-null: every multiplication of a pass is by a table entry.
    const struct fe2_t* roots_s;
    const struct fe2_t* tw_full_s;
};

ZK_D uint32_t bitrev32(uint32_t v, uint32_t bits) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; ++i) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
}

// shared memory holds the tile as two 16-byte planes so that 8 consecutive
// threads touch 128 contiguous bytes per LDS.128/STS.128 (no bank conflicts).
ZK_D fe_t tile_ld(const half_t* sm, uint32_t tile_elems, uint32_t idx) {
    fe_t r;
    half_t lo = sm[idx], hi = sm[tile_elems + idx];
    for (int i = 0; i < 4; ++i) { r.l[i] = lo.v[i]; r.l[4 + i] = hi.v[i]; }
    return r;
}
ZK_D void tile_st(half_t* sm, uint32_t tile_elems, uint32_t idx, const fe_t& v) {
    half_t lo, hi;
    for (int i = 0; i < 4; ++i) { lo.v[i] = v.l[i]; hi.v[i] = v.l[4 + i]; }
    sm[idx] = lo; sm[tile_elems + idx] = hi;
}

// omega^E for E < N via the two-level table
ZK_D fe_t ntt_twiddle(const NttPassArgs& a, uint32_t E) {
    uint32_t lo = E & ((1u << a.tw_lo_bits) - 1), hi = E >> a.tw_lo_bits;
    if (a.tw_full) return a.tw_full[E];
    if (hi == 0) return a.tw_lo[lo];
    if (lo == 0) return a.tw_hi[hi];
    return Fr::mul(a.tw_lo[lo], a.tw_hi[hi]);
}

// One tile of one pass.  nthreads must divide the butterfly count or not — any
// value works; each phase strides over its work items.
ZK_D void ntt_pass_block(const NttPassArgs& a, uint32_t bid, uint32_t nthreads, half_t* sm) {
    const uint32_t M = 1u << a.log_m, TW = 1u << a.log_tw, tile = M * TW;
    const uint32_t L = 1u << a.log_l;

    // ---- tile coordinates -------------------------------------------------
    // non-last: bid -> (h, l0); element (m, c) lives at  h*M*L + m*L + l0 + c
    // last:     bid -> (k1_0, rho'); row c is  rho = (k1_0 + c) << log_mid | rho'
    uint32_t h = 0, l0 = 0, k1_0 = 0, rho_mid = 0;
    size_t batch_base = 0;
    if (!a.is_last) {
        uint32_t tiles_per_h = L >> a.log_tw;
        h = bid / tiles_per_h; l0 = (bid % tiles_per_h) << a.log_tw;
    } else {
        uint32_t mid = 1u << a.log_mid;
        uint32_t bl = bid;
        if (a.batch_tiles) { batch_base = (size_t)(bid / a.batch_tiles) << a.log_n; bl = bid % a.batch_tiles; }
        rho_mid = bl % mid; k1_0 = (bl / mid) << a.log_tw;
    }

    // ---- load -------------------------------------------------------------
    ZK_PHASE_BEGIN(tid, nthreads)
    for (uint32_t idx = tid; idx < tile; idx += nthreads) {
        uint32_t m, c;
        size_t g;
        if (!a.is_last) {
            m = idx >> a.log_tw; c = idx & (TW - 1);
            g = ((size_t)h << (a.log_m + a.log_l)) + ((size_t)m << a.log_l) + l0 + c;
        } else {
            c = idx >> a.log_m; m = idx & (M - 1);
            size_t rho = ((size_t)(k1_0 + c) << a.log_mid) | rho_mid;
            g = batch_base + (rho << a.log_m) + m;
        }
        fe_t v;
        if (g < a.n_in) {
            v = a.in[a.in_mask ? (g & a.in_mask) : g];
            if (a.use_pre) { uint32_t r3 = (uint32_t)(g % 3); if (r3) v = Fr::mul(v, a.pre[r3]); }
            if (a.pre_tab) v = Fr::mul(v, a.pre_tab[g]);
        } else {
            v = Fr::zero();
        }
        tile_st(sm, tile, (m << a.log_tw) + c, v);
    }
    ZK_PHASE_END

    // ---- DIF butterflies along m -------------------------------------------
    for (uint32_t lh = a.log_m; lh-- > 0;) {           // half-size h = 2^lh
        ZK_PHASE_BEGIN(tid, nthreads)
        const uint32_t half = 1u << lh;
        // Butterfly p of this stage = (group, j), twiddle w^j.  With tiles of >= 8 columns a quarter-warp is one p
        // (8 contiguous 16-byte halves: conflict-free whatever p is), so p is decoded j-major: the threads of a warp
        // share j, and the warps whose j is 0 — 1 / 2^lh of the stage — skip the multiplication instead of idling
        // through it beside lanes that need it.  Narrower tiles keep the group-major order (contiguous rows).
        const uint32_t glog = a.log_m - 1 - lh;
        const bool jmajor = a.log_tw >= 3 && glog >= 2;
        for (uint32_t b = tid; b < tile / 2; b += nthreads) {
            uint32_t p = b >> a.log_tw, c = b & (TW - 1);
            uint32_t j = jmajor ? (p >> glog) : (p & (half - 1));
            uint32_t grp = jmajor ? (p & ((1u << glog) - 1)) : (p >> lh);
            uint32_t i0 = (grp << (lh + 1)) + j;
            uint32_t s0 = (i0 << a.log_tw) + c, s1 = ((i0 + half) << a.log_tw) + c;
            fe_t x = tile_ld(sm, tile, s0), y = tile_ld(sm, tile, s1);
            fe_t s = Fr::add(x, y), d = Fr::sub(x, y);
            if (j != 0) d = Fr::mul(d, a.roots[(size_t)j << (a.log_roots - lh - 1)]);
            tile_st(sm, tile, s0, s); tile_st(sm, tile, s1, d);
        }
        ZK_PHASE_END
    }

    // ---- store --------------------------------------------------------------
    ZK_PHASE_BEGIN(tid, nthreads)
    for (uint32_t idx = tid; idx < tile; idx += nthreads) {
        uint32_t k = idx >> a.log_tw, c = idx & (TW - 1);
        fe_t v = tile_ld(sm, tile, (bitrev32(k, a.log_m) << a.log_tw) + c);
        size_t g;
        if (!a.is_last) {
            uint32_t l = l0 + c;
            g = ((size_t)h << (a.log_m + a.log_l)) + ((size_t)k << a.log_l) + l;
            // w_{ML}^{l k} = omega^{l k N/(ML)}
            uint64_t E = ((uint64_t)(l + a.l_offset) * k) << a.tw_shift;
            if (E) v = Fr::mul(v, ntt_twiddle(a, (uint32_t)E));
            if (a.scatter) {
                const uint32_t dest = k >> a.log_rows_per_rank, row = k & ((1u << a.log_rows_per_rank) - 1);
                a.peers[dest][((size_t)row << a.log_c_total) + l + a.l_offset] = v;
                continue;
            }
        } else {
            // out index = k_1 + M_1 * rev(rho') + (M_1 * mid) * k ; mid digit order is preserved
            // (block plans have at most 3 passes: a single middle digit; the warp-level kernel handles two, ntt_warp.cuh)
            g = (size_t)(k1_0 + c) + ((size_t)rho_mid << a.log_m1) + ((size_t)k << (a.log_m1 + a.log_mid));
            if (a.use_post) v = Fr::mul(v, a.post[g % 3]);
            g += batch_base;
        }
        a.out[g] = v;
    }
    ZK_PHASE_END
}

// Table builders: out[i] = base^(i << shift)
ZK_D void ntt_pow_table_thread(fe_t* out, const fe_t& base, uint32_t i, uint32_t shift) {
    out[i] = Fr::pow_u64(base, (unsigned long long)i << shift);
}

}  // namespace b200zk
