// Host-side planning and finishing for msm.cuh (pure C++, no CUDA types).
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstdlib>
#include "host_field.hpp"

namespace b200zk {

struct MsmShape {
    uint32_t c, nwin, log_t;
    size_t nbuckets;         // nwin << (c-1)
};

// Window choice: accumulate does n*W mixed additions (10 muls each), reduce does
// 2 * 2^(c-1) * W full additions (14 muls) at much lower parallel efficiency.
inline MsmShape msm_plan_shape(size_t n, int force_c = 0) {
    MsmShape s{};
    uint32_t lg = 0; while (((size_t)1 << (lg + 1)) <= n) ++lg;
    uint32_t c = lg > 4 ? lg - 4 : 2;
    if (c < 4) c = 4;
    if (c > 16) c = 16;
    if (force_c > 0) c = (uint32_t)force_c;
    s.c = c;
    s.nwin = (255 + c - 1) / c;
    uint32_t log_b = c - 1;
    uint32_t log_m = 5;                                 // 32 buckets per reduce thread
    if (const char* e = getenv("B200ZK_MSM_REDUCE_M")) { long v = strtol(e, nullptr, 10); if (v >= 1 && v <= 8) log_m = (uint32_t)v; }
    s.log_t = log_b > log_m ? log_b - log_m : 0;
    s.nbuckets = (size_t)s.nwin << (c - 1);
    return s;
}

// Fixed-base mode (bases known ahead: the SRS): one bucket set shared by all windows, so the
// window can be wider — accumulate work is n * ceil(255 / c) additions, the reduce costs
// c * 2^(c-2) additions once.  c ~ log2(n) - 2.
inline MsmShape msm_pre_shape(size_t n_table) {
    MsmShape s{};
    uint32_t lg = 0; while (((size_t)1 << (lg + 1)) <= n_table) ++lg;
    // measured (profiles/r01_msm_fixed_base_window_sweep.txt): c = 16 is best up to 2^19 points,
    // 17 from 2^20; wider windows lose to the c * 2^(c-2) additions of the reduce
    uint32_t c = lg >= 20 ? 17 : 16;
    // small SRS (round 2): the bucket reduce costs c * 2^(c-2) additions whatever n is and runs far below the multiplication
    // rate (a latency-bound tree: 0.42 ms per commit for 2^15 buckets), so below 2^18 points the window shrinks with n —
    // Merkle tree v3 at k = 14: 6 reduce launches of 0.42 ms were 29 % of the 9.2 ms proof
    // (profiles/r02_launches_v3_k14_summary.txt); window sweep at k = 14: c = 10 / 11 / 12 / 13 -> 7.7 / 8.0 / 8.1 / 8.4 ms per proof
    if (lg < 18) c = lg >= 12 ? lg - 4 : 8;
    if (c < 8) c = 8;
    if (const char* e = getenv("B200ZK_MSM_PRE_C")) { long v = strtol(e, nullptr, 10); if (v >= 4 && v <= 24) c = (uint32_t)v; }
    s.c = c;
    s.nwin = (255 + c - 1) / c;
    uint32_t log_b = c - 1;
    uint32_t log_m = 5;                                 // 32 buckets per reduce thread ...
    if (log_b >= 7 && log_b < 12) log_m = log_b - 7;    // ... fewer when that would leave less than one 128-thread block per bit (the batched path needs log_t >= 7)
    if (const char* e = getenv("B200ZK_MSM_PRE_REDUCE_M")) { long v = strtol(e, nullptr, 10); if (v >= 1 && v <= 8) log_m = (uint32_t)v; }
    s.log_t = log_b > log_m ? log_b - log_m : 0;
    s.nbuckets = (size_t)1 << (c - 1);
    return s;
}

// sum_t 2^t * bit_sums[t]  (fixed-base mode; c XYZZ points)
inline host::HXyzz msm_finish_bits(const void* bit_sums, uint32_t c) {
    using namespace host;
    const HXyzz* s = (const HXyzz*)bit_sums;
    HXyzz acc = hx_identity();
    for (uint32_t t = c; t-- > 0;) { acc = hx_dbl(acc); acc = hx_add(acc, s[t]); }
    return acc;
}

// sum_j 2^(c j) * window_sums[j].  window_sums: nwin XYZZ points (device format).
inline host::HXyzz msm_finish(const void* window_sums, uint32_t nwin, uint32_t c) {
    using namespace host;
    const HXyzz* w = (const HXyzz*)window_sums;
    HXyzz acc = hx_identity();
    for (uint32_t j = nwin; j-- > 0;) {
        for (uint32_t i = 0; i < c; ++i) acc = hx_dbl(acc);
        acc = hx_add(acc, w[j]);
    }
    return acc;
}

}  // namespace b200zk
