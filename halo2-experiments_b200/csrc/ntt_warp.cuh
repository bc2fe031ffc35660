// Warp-level pass of the radix-2 NTT (same transform, pass structure, index maps and hooks as
// ntt_pass_block in ntt.cuh — see there for the reference mapping to halo2's best_fft).
//
// ntt_pass_block runs one butterfly stage per __syncthreads with every operand going through
// shared memory.  Here ONE WARP owns a tile of 128 elements = 2^lm points x 2^(7-lm) adjacent
// columns (4 <= lm <= 7), every lane keeps 4 elements in registers, and the <= 7 butterfly
// stages run in register rounds; a stage pairs the elements whose tile index
// u = point * TW + column differs in one bit:
//     round 1  lane holds u = e*32 + lane                        bits 6,5 lane-local
//     round 2  lane holds u = (lane>>3)*32 + e*8 + (lane&7)      bits 4,3 lane-local
//     round 3  lane holds u = (lane>>1)*8 + e*2 + (lane&1)       bits 2,1 lane-local
//     bit 0    partner lane = lane ^ 1: its twiddle is always 1, so the stage is a shuffle, an
//              addition on the even lane and a subtraction on the odd one.
// slot = u ^ ((u >> 2) & 6) keeps every quarter-warp access conflict-free in all three layouts.
//
// No block-level barrier anywhere; 80 registers per thread, 24 resident warps per SM.  Measured on
// B200 (profiles/r01_ntt_warp_vs_block.txt): 2^16 0.037 vs 0.064 ms, 2^20 0.235 vs 0.284 ms,
// 2^21 0.478 vs 0.505 ms against the block kernel.  Three passes of at most 7 bits cover 2^21, four
// cover 2^28 (round 2: the last pass then stores its two middle digits swapped, see the store phase): with
// the constant-operand multiplier and the [0, 2p) butterflies this kernel outruns the block kernel at every
// size — 2^22 0.81 vs 1.19 ms, 2^24 3.25 vs 3.85 ms, 2^26 15.7 vs 19.2 ms — so the block kernel is left with
// the sizes below 2^12 and the column step of the sharded four-step.  (A variant with 8 elements per lane — 128 registers, 16
// warps per SM, 288 KB of code — was slower than both from 2^21 up and is not kept.)
#pragma once
#include "ntt.cuh"

namespace b200zk {

static constexpr uint32_t NTT_WARP_TILE_LOG = 7;
static constexpr uint32_t NTT_WARPS_PER_BLOCK = 8;

__device__ __forceinline__ uint32_t ntt_warp_slot(uint32_t u) { return u ^ ((u >> 2) & 6u); }

template <int LAYOUT> __device__ __forceinline__ uint32_t ntt_warp_u(uint32_t lane, uint32_t e) {
    if (LAYOUT == 0) return (e << 5) | lane;
    if (LAYOUT == 1) return ((lane >> 3) << 5) | (e << 3) | (lane & 7u);
    return ((lane >> 1) << 3) | (e << 1) | (lane & 1u);
}
template <int LAYOUT> __device__ __forceinline__ void ntt_warp_put(half_t* sm, uint32_t lane, const fe_t (&x)[4]) {
#pragma unroll
    for (uint32_t e = 0; e < 4; ++e) tile_st(sm, 128, ntt_warp_slot(ntt_warp_u<LAYOUT>(lane, e)), x[e]);
}
template <int LAYOUT> __device__ __forceinline__ void ntt_warp_get(const half_t* sm, uint32_t lane, fe_t (&x)[4]) {
#pragma unroll
    for (uint32_t e = 0; e < 4; ++e) x[e] = tile_ld(sm, 128, ntt_warp_slot(ntt_warp_u<LAYOUT>(lane, e)));
}
// d * (table entry ix): CIOS against the Montgomery-form table, or the constant-operand product against the {w, wq} table.
// With the constant-operand multiplier the values of a pass live in [0, 2p) (Field::mul_shoup_lazy): LAZY butterflies.
template <bool SHOUP> __device__ __forceinline__ fe_t ntt_warp_mul_root(const NttPassArgs& a, const fe_t& d, size_t ix) {
    if (SHOUP) { const fe2_t t = a.roots_s[ix]; return Fr::mul_shoup_lazy(d, t.w, t.wq); }
    return Fr::mul(d, a.roots[ix]);
}
template <bool LAZY> __device__ __forceinline__ fe_t ntt_bf_add(const fe_t& x, const fe_t& y) { return LAZY ? Fr::add_lazy(x, y) : Fr::add(x, y); }
// difference that feeds a multiplication (LAZY: left in (0, 4p)) ...
template <bool LAZY> __device__ __forceinline__ fe_t ntt_bf_sub_raw(const fe_t& x, const fe_t& y) { return LAZY ? Fr::sub_raw(x, y) : Fr::sub(x, y); }
// ... and the same value when it is kept as it is
template <bool LAZY> __device__ __forceinline__ fe_t ntt_bf_keep(fe_t d) { if (LAZY) Fr::reduce_2p(d); return d; }
template <int LAYOUT, int B, int EB, bool SHOUP>
__device__ __forceinline__ void ntt_warp_stage(const NttPassArgs& a, uint32_t lane, uint32_t log_tw, fe_t (&x)[4]) {
    if ((uint32_t)B < log_tw) return;
    const uint32_t lh = (uint32_t)B - log_tw;
#pragma unroll
    for (uint32_t e0 = 0; e0 < 4; ++e0) {
        if (e0 & (1u << EB)) continue;
        const uint32_t e1 = e0 | (1u << EB);
        const uint32_t j = (ntt_warp_u<LAYOUT>(lane, e0) >> log_tw) & ((1u << lh) - 1u);
        fe_t s = ntt_bf_add<SHOUP>(x[e0], x[e1]), d = ntt_bf_sub_raw<SHOUP>(x[e0], x[e1]);
        if (j != 0) d = ntt_warp_mul_root<SHOUP>(a, d, (size_t)j << (a.log_roots - lh - 1));
        else d = ntt_bf_keep<SHOUP>(d);
        x[e0] = s; x[e1] = d;
    }
}

// SKIP: test the tile for all-zero inputs (first pass of an iNTT of Lagrange columns, see below); the dense
// instantiation carries none of it.
template <bool SKIP, bool SHOUP>
__device__ __forceinline__ void ntt_pass_warp(const NttPassArgs& a, uint32_t wid, uint32_t lane, half_t* sm) {
    const uint32_t log_tw = NTT_WARP_TILE_LOG - a.log_m, TW = 1u << log_tw;
    uint32_t h = 0, l0 = 0, k1_0 = 0, rho_mid = 0;
    size_t batch_base = 0;
    if (!a.is_last) {
        uint32_t tiles_per_h = (1u << a.log_l) >> log_tw;
        h = wid / tiles_per_h; l0 = (wid % tiles_per_h) << log_tw;
    } else {
        uint32_t mid = 1u << a.log_mid, wl = wid;
        if (a.batch_tiles) { batch_base = (size_t)(wid / a.batch_tiles) << a.log_n; wl = wid % a.batch_tiles; }
        rho_mid = wl % mid; k1_0 = (wl / mid) << log_tw;
    }
    fe_t x[4];
    bool nz = false;
#pragma unroll
    for (uint32_t e = 0; e < 4; ++e) {
        const uint32_t u = ntt_warp_u<0>(lane, e), m = u >> log_tw, c = u & (TW - 1);
        size_t g;
        if (!a.is_last) g = ((size_t)h << (a.log_m + a.log_l)) + ((size_t)m << a.log_l) + l0 + c;
        else g = batch_base + ((((size_t)(k1_0 + c) << a.log_mid) | rho_mid) << a.log_m) + m;
        if (g < a.n_in) {
            x[e] = a.in[a.in_mask ? (g & a.in_mask) : g];
            if (SKIP) nz |= !Fr::is_zero(x[e]);
        } else {
            x[e] = Fr::zero();
        }
    }
    // A tile whose 128 inputs are all zero transforms to zero: the warp skips the arithmetic and only writes its
    // outputs.  Witness columns of a padded circuit are zero outside the used rows and the blinding rows, so the
    // first pass of their iNTT (column tiles at stride n / 128) is almost entirely such tiles.
    const bool live = SKIP ? __any_sync(0xffffffffu, nz) : true;
    if (live) {
#pragma unroll
    for (uint32_t e = 0; e < 4; ++e) {
        const uint32_t u = ntt_warp_u<0>(lane, e), m = u >> log_tw, c = u & (TW - 1);
        size_t g;
        if (!a.is_last) g = ((size_t)h << (a.log_m + a.log_l)) + ((size_t)m << a.log_l) + l0 + c;
        else g = batch_base + ((((size_t)(k1_0 + c) << a.log_mid) | rho_mid) << a.log_m) + m;
        if (g < a.n_in) {
            if (a.use_pre) { uint32_t r3 = (uint32_t)(g % 3); if (r3) x[e] = Fr::mul(x[e], a.pre[r3]); }
            if (a.pre_tab) x[e] = Fr::mul(x[e], a.pre_tab[g]);
        }
    }
    ntt_warp_stage<0, 6, 1, SHOUP>(a, lane, log_tw, x);
    ntt_warp_stage<0, 5, 0, SHOUP>(a, lane, log_tw, x);
    ntt_warp_put<0>(sm, lane, x);
    __syncwarp();
    ntt_warp_get<1>(sm, lane, x);
    ntt_warp_stage<1, 4, 1, SHOUP>(a, lane, log_tw, x);
    ntt_warp_stage<1, 3, 0, SHOUP>(a, lane, log_tw, x);
    __syncwarp();
    ntt_warp_put<1>(sm, lane, x);
    __syncwarp();
    ntt_warp_get<2>(sm, lane, x);
    ntt_warp_stage<2, 2, 1, SHOUP>(a, lane, log_tw, x);
    if (log_tw == 0) {
        // Stage 1 of a 128-point tile: the twiddle index is the lane's parity, so even lanes have two trivial
        // butterflies and odd lanes two multiplications by w_4 — executed as written, the warp would spend two
        // multiplications with half its lanes idle.  The odd lane hands one of its differences to its even
        // neighbour instead: every lane multiplies exactly once.
        const bool odd = lane & 1u;
        fe_t s0 = ntt_bf_add<SHOUP>(x[0], x[1]), d0 = ntt_bf_sub_raw<SHOUP>(x[0], x[1]);
        fe_t s1 = ntt_bf_add<SHOUP>(x[2], x[3]), d1 = ntt_bf_sub_raw<SHOUP>(x[2], x[3]);
        fe_t y;
#pragma unroll
        for (int i = 0; i < 8; ++i) y.l[i] = __shfl_xor_sync(0xffffffffu, d1.l[i], 1);
        fe_t prod = ntt_warp_mul_root<SHOUP>(a, odd ? d0 : y, (size_t)1 << (a.log_roots - 2));
#pragma unroll
        for (int i = 0; i < 8; ++i) y.l[i] = __shfl_xor_sync(0xffffffffu, prod.l[i], 1);
        x[0] = s0; x[1] = odd ? prod : ntt_bf_keep<SHOUP>(d0); x[2] = s1; x[3] = odd ? y : ntt_bf_keep<SHOUP>(d1);
    } else {
        ntt_warp_stage<2, 1, 0, SHOUP>(a, lane, log_tw, x);
    }
    if (log_tw == 0) {                                         // bit 0 carries a point: twiddle-free stage across lane pairs
        const bool odd = lane & 1u;
#pragma unroll
        for (uint32_t e = 0; e < 4; ++e) {
            fe_t y;
#pragma unroll
            for (int i = 0; i < 8; ++i) y.l[i] = __shfl_xor_sync(0xffffffffu, x[e].l[i], 1);
            x[e] = odd ? ntt_bf_keep<SHOUP>(ntt_bf_sub_raw<SHOUP>(y, x[e])) : ntt_bf_add<SHOUP>(x[e], y);
        }
    }
    }   // live
#pragma unroll
    for (uint32_t e = 0; e < 4; ++e) {
        const uint32_t u = ntt_warp_u<2>(lane, e);
        const uint32_t p = u >> log_tw, c = u & (TW - 1);
        const uint32_t k = __brev(p) >> (32 - a.log_m);
        fe_t v = x[e];
        size_t g;
        if (!a.is_last) {
            uint32_t l = l0 + c;
            g = ((size_t)h << (a.log_m + a.log_l)) + ((size_t)k << a.log_l) + l;
            uint64_t E = ((uint64_t)(l + a.l_offset) * k) << a.tw_shift;
            if (E && live) {
                if (SHOUP) { const fe2_t t = a.tw_full_s[(uint32_t)E]; v = Fr::mul_shoup_lazy(v, t.w, t.wq); }
                else v = Fr::mul(v, ntt_twiddle(a, (uint32_t)E));
            }
            if (SHOUP && a.canonical_out) Fr::reduce_once(v);  // column step: what leaves the call is canonical
            if (a.scatter) {                                   // sharded four-step column step: row k to its owner (NttPassArgs::scatter)
                const uint32_t dest = k >> a.log_rows_per_rank, row = k & ((1u << a.log_rows_per_rank) - 1u);
                a.peers[dest][((size_t)row << a.log_c_total) + l + a.l_offset] = v;
                continue;
            }
        } else {
            // natural output index k1 + M1 k2 (+ M1 M2 k3) + M1 M_mid k: with four passes the row index carries k2 above k3
            g = (size_t)(k1_0 + c) + ((size_t)(rho_mid >> a.log_m3) << a.log_m1) +
                ((size_t)(rho_mid & ((1u << a.log_m3) - 1u)) << (a.log_m1 + a.log_mid - a.log_m3)) + ((size_t)k << (a.log_m1 + a.log_mid));
            if (SHOUP) Fr::reduce_once(v);                     // [0, 2p) -> the canonical value leaves the transform
            if (a.use_post && live) v = Fr::mul(v, a.post[g % 3]);
            g += batch_base;
        }
        a.out[g] = v;
    }
}

}  // namespace b200zk
