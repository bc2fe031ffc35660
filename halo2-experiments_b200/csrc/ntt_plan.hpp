// Host-side pass planner for ntt.cuh (pure C++, no CUDA types).
#pragma once
#include <cstdint>
#include <algorithm>

namespace b200zk {

struct NttPassShape {
    uint32_t log_m, log_l, log_tw, is_last, log_m1, log_mid, blocks;
    uint32_t log_m3;         // four passes: size of the third digit (the faster half of the log_mid middle bits); else 0
};

struct NttShape {
    uint32_t log_n, npass;
    NttPassShape pass[4];
    uint32_t log_roots;      // roots table = w_R^j, j < R/2
    uint32_t tw_lo_bits;     // two-level twiddle split
};

// max_log_m: largest in-shared-memory DFT (<= 10); max_log_tw: widest tile.
// tile_cap_log: log2 of the largest tile (M * TW elements) shared memory can hold.
inline NttShape ntt_plan_shape(uint32_t log_n, uint32_t max_log_m, uint32_t max_log_tw, uint32_t tile_cap_log) {
    NttShape s{};
    s.log_n = log_n;
    uint32_t P = log_n == 0 ? 1 : (log_n + max_log_m - 1) / max_log_m;
    if (P > 3) P = 3;                                   // caller guarantees log_n <= 3 * max_log_m
    s.npass = P;
    uint32_t base = log_n / P, rem = log_n % P, d[3] = {0, 0, 0};
    for (uint32_t p = 0; p < P; ++p) d[p] = base + (p < rem ? 1 : 0);
    uint32_t log_l = log_n;
    for (uint32_t p = 0; p < P; ++p) {
        NttPassShape& q = s.pass[p];
        log_l -= d[p];
        q.log_m = d[p];
        q.log_l = log_l;
        q.is_last = (p + 1 == P);
        q.log_m1 = P > 1 ? d[0] : 0;
        q.log_mid = P == 3 ? d[1] : 0;
        uint32_t cap = tile_cap_log > q.log_m ? tile_cap_log - q.log_m : 0;
        uint32_t tw = std::min(max_log_tw, cap);
        tw = std::min(tw, q.is_last ? q.log_m1 : q.log_l);
        q.log_tw = tw;
        q.blocks = 1u << (log_n - q.log_m - tw);
    }
    s.log_roots = d[0];
    for (uint32_t p = 1; p < P; ++p) s.log_roots = std::max(s.log_roots, d[p]);
    s.tw_lo_bits = (log_n + 1) / 2;
    return s;
}

// Plan for the warp-level kernel (ntt_warp.cuh): every pass is a 2^4..2^7-point transform on
// tiles of 128 elements (TW = 128 / M columns); three passes cover 2^21, four cover 2^28.
inline bool ntt_warp_eligible(uint32_t log_n, uint32_t max_log = 26) { return log_n >= 12 && log_n <= 28 && log_n <= max_log; }

inline NttShape ntt_plan_shape_warp(uint32_t log_n) {
    const uint32_t tile_log = 7;
    NttShape s{};
    s.log_n = log_n;
    uint32_t P = (log_n + tile_log - 1) / tile_log;
    s.npass = P;
    uint32_t base = log_n / P, rem = log_n % P, d[4] = {0, 0, 0, 0};
    for (uint32_t p = 0; p < P; ++p) d[p] = base + (p < rem ? 1 : 0);
    uint32_t log_l = log_n;
    for (uint32_t p = 0; p < P; ++p) {
        NttPassShape& q = s.pass[p];
        log_l -= d[p];
        q.log_m = d[p];
        q.log_l = log_l;
        q.is_last = (p + 1 == P);
        q.log_m1 = P > 1 ? d[0] : 0;
        // the last pass sees the middle digits as one index (second digit slower); it stores them digit-reversed
        q.log_mid = P == 3 ? d[1] : P == 4 ? d[1] + d[2] : 0;
        q.log_m3 = P == 4 ? d[2] : 0;
        q.log_tw = tile_log - d[p];
        q.blocks = 1u << (log_n - tile_log);            // warp tiles
    }
    s.log_roots = d[0];
    s.tw_lo_bits = (log_n + 1) / 2;
    return s;
}

}  // namespace b200zk
