// Warp-level pass of the radix-2 NTT (same transform, pass structure, index maps and hooks as
// ntt_pass_block in ntt.cuh — see there for the reference mapping to halo2's best_fft).
//
// ntt_pass_block runs one butterfly stage per __syncthreads with every operand going through
// shared memory: the IMAD.WIDE pipe sat at ~70 % and issue slots at ~42 % while warps waited on
// barriers (profiles/r01_ncu_ntt_summary.txt).  Here ONE WARP owns a tile of 256 elements =
// 2^lm points x 2^(8-lm) adjacent columns (5 <= lm <= 8), every lane keeps 8 elements in
// registers, and the <= 8 butterfly stages run in three register rounds:
//     tile index u = point * TW + column  (8 bits);  a stage pairs the elements differing in one u bit
//     round 1  lane holds u = e*32 + lane                          bits 7,6,5 are lane-local
//     round 2  lane holds u = (lane>>2)*32 + e*4 + (lane&3)        bits 4,3,2 are lane-local
//     round 3  lane holds u = lane*8 + e                           bits 1,0   are lane-local
// Between rounds the 8 KB tile is transposed through warp-private shared memory (two 16-byte
// planes, slot = u ^ ((u >> 3) & 7), which makes every quarter-warp access cover the eight
// 16-byte bank groups once in all three layouts).  No block-level barrier anywhere; each lane has
// four independent field multiplications in flight per stage.
#pragma once
#include "ntt.cuh"

namespace b200zk {

static constexpr uint32_t NTT_WARP_TILE_LOG = 8;
static constexpr uint32_t NTT_WARPS_PER_BLOCK = 4;

__device__ __forceinline__ uint32_t ntt_warp_slot(uint32_t u) { return u ^ ((u >> 3) & 7u); }

template <int LAYOUT> __device__ __forceinline__ uint32_t ntt_warp_u(uint32_t lane, uint32_t e) {
    if (LAYOUT == 0) return (e << 5) | lane;
    if (LAYOUT == 1) return ((lane >> 2) << 5) | (e << 2) | (lane & 3u);
    return (lane << 3) | e;
}

template <int LAYOUT> __device__ __forceinline__ void ntt_warp_put(half_t* sm, uint32_t lane, const fe_t (&x)[8]) {
#pragma unroll
    for (uint32_t e = 0; e < 8; ++e) tile_st(sm, 256, ntt_warp_slot(ntt_warp_u<LAYOUT>(lane, e)), x[e]);
}
template <int LAYOUT> __device__ __forceinline__ void ntt_warp_get(const half_t* sm, uint32_t lane, fe_t (&x)[8]) {
#pragma unroll
    for (uint32_t e = 0; e < 8; ++e) x[e] = tile_ld(sm, 256, ntt_warp_slot(ntt_warp_u<LAYOUT>(lane, e)));
}

// One DIF stage on tile bit B, which is bit EB of the lane-local element index in this layout.
template <int LAYOUT, int B, int EB>
__device__ __forceinline__ void ntt_warp_stage(const NttPassArgs& a, uint32_t lane, uint32_t log_tw, fe_t (&x)[8]) {
    if ((uint32_t)B < log_tw) return;                          // column bit: not part of the transform
    const uint32_t lh = (uint32_t)B - log_tw;
#pragma unroll
    for (uint32_t e0 = 0; e0 < 8; ++e0) {
        if (e0 & (1u << EB)) continue;
        const uint32_t e1 = e0 | (1u << EB);
        const uint32_t j = (ntt_warp_u<LAYOUT>(lane, e0) >> log_tw) & ((1u << lh) - 1u);
        fe_t s = Fr::add(x[e0], x[e1]), d = Fr::sub(x[e0], x[e1]);
        if (j != 0) d = Fr::mul(d, a.roots[(size_t)j << (a.log_roots - lh - 1)]);
        x[e0] = s; x[e1] = d;
    }
}

// wid = global warp index = tile index (same tile numbering as ntt_pass_block with TW = 256 / M).
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
    // The inter-pass twiddles of this tile are scattered 32-byte reads from a table as large as the
    // transform: ask L2 for them now, they are needed after the last round.
    if (!a.is_last && a.tw_full) {
        const bool r3 = log_tw < 2;
#pragma unroll
        for (uint32_t e = 0; e < 8; ++e) {
            const uint32_t u = r3 ? ntt_warp_u<2>(lane, e) : ntt_warp_u<1>(lane, e);
            const uint32_t k = __brev(u >> log_tw) >> (32 - a.log_m);
            const uint64_t E = ((uint64_t)(l0 + (u & (TW - 1)) + a.l_offset) * k) << a.tw_shift;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.tw_full + (uint32_t)E));
        }
    }
    fe_t x[8];
    // ---- load (layout of round 1) ----
#pragma unroll
    for (uint32_t e = 0; e < 8; ++e) {
        const uint32_t u = ntt_warp_u<0>(lane, e), m = u >> log_tw, c = u & (TW - 1);
        size_t g;
        if (!a.is_last) g = ((size_t)h << (a.log_m + a.log_l)) + ((size_t)m << a.log_l) + l0 + c;
        else g = batch_base + ((((size_t)(k1_0 + c) << a.log_mid) | rho_mid) << a.log_m) + m;
        if (g < a.n_in) {
            x[e] = a.in[a.in_mask ? (g & a.in_mask) : g];
            if (a.use_pre) { uint32_t r3 = (uint32_t)(g % 3); if (r3) x[e] = Fr::mul(x[e], a.pre[r3]); }
            if (a.pre_tab) x[e] = Fr::mul(x[e], a.pre_tab[g]);
        } else {
            x[e] = Fr::zero();
        }
    }
    // ---- round 1: bits 7, 6, 5 ----
    ntt_warp_stage<0, 7, 2>(a, lane, log_tw, x);
    ntt_warp_stage<0, 6, 1>(a, lane, log_tw, x);
    ntt_warp_stage<0, 5, 0>(a, lane, log_tw, x);
    ntt_warp_put<0>(sm, lane, x);
    __syncwarp();
    ntt_warp_get<1>(sm, lane, x);
    // ---- round 2: bits 4, 3, 2 ----
    ntt_warp_stage<1, 4, 2>(a, lane, log_tw, x);
    ntt_warp_stage<1, 3, 1>(a, lane, log_tw, x);
    ntt_warp_stage<1, 2, 0>(a, lane, log_tw, x);
    const bool round3 = log_tw < 2;                            // uniform: bits 1 / 0 carry points only for M >= 128
    if (round3) {
        __syncwarp();
        ntt_warp_put<1>(sm, lane, x);
        __syncwarp();
        ntt_warp_get<2>(sm, lane, x);
        ntt_warp_stage<2, 1, 1>(a, lane, log_tw, x);
        ntt_warp_stage<2, 0, 0>(a, lane, log_tw, x);
    }
    // ---- store: position p holds X[bitrev(p)] ----
#pragma unroll
    for (uint32_t e = 0; e < 8; ++e) {
        const uint32_t u = round3 ? ntt_warp_u<2>(lane, e) : ntt_warp_u<1>(lane, e);
        const uint32_t p = u >> log_tw, c = u & (TW - 1);
        const uint32_t k = __brev(p) >> (32 - a.log_m);
        fe_t v = x[e];
        size_t g;
        if (!a.is_last) {
            uint32_t l = l0 + c;
            g = ((size_t)h << (a.log_m + a.log_l)) + ((size_t)k << a.log_l) + l;
            uint64_t E = ((uint64_t)(l + a.l_offset) * k) << a.tw_shift;
            if (E) v = Fr::mul(v, ntt_twiddle(a, (uint32_t)E));
        } else {
            g = (size_t)(k1_0 + c) + ((size_t)rho_mid << a.log_m1) + ((size_t)k << (a.log_m1 + a.log_mid));
            if (a.use_post) v = Fr::mul(v, a.post[g % 3]);
            g += batch_base;
        }
        a.out[g] = v;
    }
}

}  // namespace b200zk
