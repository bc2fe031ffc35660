// Signed-digit Pippenger multi-scalar multiplication over bn256 G1, replacing
// halo2_proofs `arithmetic::best_multiexp` (PSE v2023_02_02 src/arithmetic.rs),
// which the reference reaches through ParamsKZG::commit / commit_lagrange inside
// keygen_vk / create_proof (/root/reference/src/circuits/utils.rs:31,40-48).
// The group element returned is independent of the bucket method, so parity with
// the CPU path is bit-exact after affine normalisation.
//
// Device pipeline (all state in HBM; launch side in msm.cu):
//   1. digits+count : scalar -> canonical (one Montgomery mul by 1), W signed c-bit digits,
//                     histogram of the bucket keys.  One warp walks 32 scalars in lockstep: all-zero
//                     warps leave at once, a window in which all lanes hit one bucket costs one atomic
//   2. scan         : exclusive prefix sum of the histogram
//   3. scatter      : counting sort of (point index, sign) entries by key
//   4. accumulate   : XYZZ mixed additions (the IMAD-bound kernel; N*W additions of 8M+2S), one
//                     thread per bucket when no bucket is long, else balanced levels of <= 16 items
//                     per task (entries -> partials -> ... -> bucket)
//   5. reduce       : sum_v v*B_v — per window by chunked running sums, or, with the fixed-base
//                     tables of an SRS (all windows share one bucket set), by bit decomposition with
//                     the first seven fold levels inside the block
//   6. fold         : tree sum of the remaining partials
// Several columns over the same SRS run as ONE such sequence (msm_run_multi: column b owns bucket
// set b), so the latency-bound tail is paid once per batch.  The per-window / per-bit sums go back to
// the host, which does the last doublings — a strictly serial chain that a CPU core finishes ~10x
// sooner than one GPU thread — and the affine normalisation for the transcript.
#pragma once
#include "curve.cuh"
#include "blockexec.cuh"

namespace b200zk {

#if defined(__CUDACC__)
ZK_D uint32_t zk_atomic_add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
#else
inline uint32_t zk_atomic_add(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
#endif

struct MsmArgs {
    const fe_t* scalars;        // n Montgomery-form Fr
    const affine_t* bases;      // n affine points
    uint32_t n;
    uint32_t c;                 // window bits
    uint32_t nwin;              // W = ceil(255 / c)
    uint32_t log_t;             // reduce: 2^log_t chunks per window
    // Fixed-base mode (SRS commits): `bases` is the table T[j*pre_stride + i] = 2^(c j) * P_i, so the
    // digit of window j selects a pre-shifted point and ALL windows share one set of 2^(c-1) buckets.
    uint32_t pre;               // 0 = per-window buckets, 1 = fixed-base table
    uint32_t pre_stride;        // points per window in the table (the params' n)
    uint32_t nbuckets;          // W << (c-1), or 1 << (c-1) in fixed-base mode
    uint32_t use_sub;           // scalars are taken as scalars[i] - sub (see params_commit_run: constant-run columns)
    fe_t sub;
    // Batched commits (fixed-base mode): `batch` columns of n scalars over the same table, column b
    // owning buckets [b * set_buckets, (b + 1) * set_buckets); nbuckets = batch * set_buckets.
    uint32_t batch, set_buckets;
    const fe_t* const* scalars_tab;     // device array of `batch` column pointers (batch > 1)
    const fe_t* subs;                   // device, per column; subs_on[b] != 0 selects it
    const uint32_t* subs_on;
    uint32_t* counts;           // [W << (c-1)]
    uint32_t* offsets;          // [(W << (c-1)) + 1]
    uint32_t* cursor;           // [W << (c-1)]
    uint32_t* entries;          // [n * W]   (index << 1) | sign
    xyzz_t* buckets;            // [W << (c-1)]
    xyzz_t* partials;           // [W << log_t]
    xyzz_t* window_sums;        // [W]
};

ZK_D uint32_t msm_window_bits(const fe_t& s, uint32_t bit, uint32_t c) {
    uint32_t idx = bit >> 5, sh = bit & 31;
    if (idx >= 8) return 0;
    uint32_t v = s.l[idx] >> sh;
    if (sh + c > 32 && idx + 1 < 8) v |= s.l[idx + 1] << (32 - sh);
    return v & ((1u << c) - 1);
}

// Calls f(key, sign) for every non-zero signed digit of scalar i.
// Digits d_j in [-2^(c-1), 2^(c-1)]; key = j * 2^(c-1) + |d_j| - 1.
template <class F> ZK_D void msm_for_each_digit(const MsmArgs& a, uint32_t i, F f) {
    fe_t s = a.scalars[i];
    if (a.use_sub) s = Fr::sub(s, a.sub);
    s = Fr::from_mont(s);
    uint32_t carry = 0, half = 1u << (a.c - 1);
    for (uint32_t j = 0; j < a.nwin; ++j) {
        uint32_t d = msm_window_bits(s, j * a.c, a.c) + carry;
        uint32_t sign = 0;
        if (d > half) { d = (1u << a.c) - d; sign = 1; carry = 1; } else carry = 0;
        if (d != 0) {
            if (a.pre) f(d - 1, sign, j * a.pre_stride + i);
            else f(j * half + d - 1, sign, i);
        }
    }
}

ZK_D void msm_count_thread(const MsmArgs& a, uint32_t i) {
    if (i >= a.n) return;
    msm_for_each_digit(a, i, [&](uint32_t key, uint32_t, uint32_t) { zk_atomic_add(&a.counts[key], 1u); });
}

ZK_D void msm_scatter_thread(const MsmArgs& a, uint32_t i) {
    if (i >= a.n) return;
    msm_for_each_digit(a, i, [&](uint32_t key, uint32_t sign, uint32_t point) {
        uint32_t pos = zk_atomic_add(&a.cursor[key], 1u);
        a.entries[pos] = (point << 1) | sign;
    });
}

// ---- exclusive scans (three launches: per-block sums, single-block scan of the sums, per-block
// scan + offset; each thread owns 8 consecutive items) --------------------------------------
// Scanned value of item i:  v = from_offsets ? in[i+1] - in[i] : in[i];  if div: v = max(1, ceil(v / div)).
static constexpr uint32_t MSM_SCAN_PER_THREAD = 8;

struct ScanArgs {
    const uint32_t* in;
    uint32_t* out;              // [total + 1]
    uint32_t* out2;             // optional copy of out[0..total) (scatter cursors)
    uint32_t* blocksums;
    uint32_t* maxv;             // optional: atomicMax of the raw (undivided) values
    uint32_t total, from_offsets, div;
};

ZK_D uint32_t scan_raw(const ScanArgs& s, uint32_t i) { return s.from_offsets ? s.in[i + 1] - s.in[i] : s.in[i]; }
ZK_D uint32_t scan_value(const ScanArgs& s, uint32_t i) {
    uint32_t v = scan_raw(s, i);
    if (s.div) { v = (v + s.div - 1) / s.div; if (v == 0) v = 1; }
    return v;
}

#if defined(__CUDACC__)
ZK_D void zk_atomic_max(uint32_t* p, uint32_t v) { atomicMax(p, v); }
#else
inline void zk_atomic_max(uint32_t* p, uint32_t v) { if (v > *p) *p = v; }
#endif

// Hillis-Steele inclusive scan of sm[0..T) (ping-pong in sm[0..2T)); returns the buffer index.
ZK_D uint32_t msm_block_inclusive_scan(uint32_t* sm, uint32_t T) {
    uint32_t cur = 0;
    for (uint32_t d = 1; d < T; d <<= 1) {
        ZK_PHASE_BEGIN(tid, T)
        uint32_t v = sm[cur * T + tid];
        if (tid >= d) v += sm[cur * T + tid - d];
        sm[(cur ^ 1) * T + tid] = v;
        ZK_PHASE_END
        cur ^= 1;
    }
    return cur;
}

// (1) blocksums[bid] = sum of the block's items (+ running max of the raw values).  sm: 2*T words.
ZK_D void scan_blocksum_block(const ScanArgs& s, uint32_t bid, uint32_t T, uint32_t* sm) {
    ZK_PHASE_BEGIN(tid, T)
    uint32_t base = (bid * T + tid) * MSM_SCAN_PER_THREAD, sum = 0, mx = 0;
    for (uint32_t i = 0; i < MSM_SCAN_PER_THREAD; ++i) if (base + i < s.total) {
        sum += scan_value(s, base + i);
        uint32_t r = scan_raw(s, base + i); if (r > mx) mx = r;
    }
    sm[tid] = sum;
    if (s.maxv && mx) zk_atomic_max(s.maxv, mx);
    ZK_PHASE_END
    uint32_t cur = msm_block_inclusive_scan(sm, T);
    ZK_PHASE_BEGIN(tid, T)
    if (tid == T - 1) s.blocksums[bid] = sm[cur * T + tid];
    ZK_PHASE_END
}

// (2) in-place exclusive scan of blocksums[0..nblocks) by one block; grand total to out[total].
ZK_D void scan_top_block(const ScanArgs& s, uint32_t nblocks, uint32_t T, uint32_t* sm) {
    const uint32_t per = (nblocks + T - 1) / T;
    ZK_PHASE_BEGIN(tid, T)
    uint32_t sum = 0, b = tid * per, e = b + per < nblocks ? b + per : nblocks;
    for (uint32_t k = b; k < e; ++k) sum += s.blocksums[k];
    sm[tid] = sum;
    ZK_PHASE_END
    uint32_t cur = msm_block_inclusive_scan(sm, T);
    ZK_PHASE_BEGIN(tid, T)
    uint32_t run = tid ? sm[cur * T + tid - 1] : 0, b = tid * per, e = b + per < nblocks ? b + per : nblocks;
    for (uint32_t k = b; k < e; ++k) { uint32_t v = s.blocksums[k]; s.blocksums[k] = run; run += v; }
    if (tid == T - 1) s.out[s.total] = sm[cur * T + tid];
    ZK_PHASE_END
}

// (3) out / out2 for the block's items.
ZK_D void scan_final_block(const ScanArgs& s, uint32_t bid, uint32_t T, uint32_t* sm) {
    ZK_PHASE_BEGIN(tid, T)
    uint32_t base = (bid * T + tid) * MSM_SCAN_PER_THREAD, sum = 0;
    for (uint32_t i = 0; i < MSM_SCAN_PER_THREAD; ++i) if (base + i < s.total) sum += scan_value(s, base + i);
    sm[tid] = sum;
    ZK_PHASE_END
    uint32_t cur = msm_block_inclusive_scan(sm, T);
    ZK_PHASE_BEGIN(tid, T)
    uint32_t base = (bid * T + tid) * MSM_SCAN_PER_THREAD;
    uint32_t run = s.blocksums[bid] + (tid ? sm[cur * T + tid - 1] : 0);
    for (uint32_t i = 0; i < MSM_SCAN_PER_THREAD; ++i) {
        if (base + i < s.total) {
            uint32_t v = scan_value(s, base + i);       // read before write: out may alias nothing, but keep order explicit
            s.out[base + i] = run; if (s.out2) s.out2[base + i] = run; run += v;
        }
    }
    ZK_PHASE_END
}

// ---- bucket accumulation ---------------------------------------------------------------------
// Fast path (every bucket has at most L entries): one thread per bucket.
ZK_D void msm_accumulate_thread(const MsmArgs& a, uint32_t key) {
    if (key >= a.nbuckets) return;
    uint32_t b = a.offsets[key], e = a.offsets[key + 1];
    xyzz_t acc = xyzz_identity();
    for (uint32_t k = b; k < e; ++k) {
        uint32_t en = a.entries[k];
        xyzz_madd(acc, a.bases[en >> 1], en & 1);
    }
    a.buckets[key] = acc;
}

// General path: witness columns put millions of points into a handful of buckets (bits, bytes,
// small integers), so buckets are cut into tasks of at most L entries and reduced in three
// balanced levels:  entries -> partial1 (per task) -> partial2 (per task of tasks) -> bucket.
// task_off* are exclusive scans of max(1, ceil(count / L)); the owning bucket of a task is found
// by binary search in the scan.
struct MsmTaskArgs {
    const uint32_t* seg_off;    // offsets of the items being reduced, per bucket   [nb + 1]
    const uint32_t* task_off;   // tasks per bucket, scanned                         [nb + 1]
    uint32_t nb;
    const uint32_t* ntasks_ptr; // device word holding the task count (= task_off[nb]); launches over-provision
    uint32_t L;
    const xyzz_t* in;           // level >= 2: partial sums of the level below
    xyzz_t* out;                // one sum per task
};

ZK_D uint32_t upper_bound_u32(const uint32_t* a, uint32_t n, uint32_t v) {   // first i with a[i] > v
    uint32_t lo = 0, hi = n;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (a[mid] <= v) lo = mid + 1; else hi = mid; }
    return lo;
}

ZK_D void msm_task_range(const MsmTaskArgs& t, uint32_t task, uint32_t& begin, uint32_t& end) {
    uint32_t b = upper_bound_u32(t.task_off, t.nb + 1, task) - 1;
    uint32_t seg = task - t.task_off[b];
    begin = t.seg_off[b] + seg * t.L;
    uint32_t stop = t.seg_off[b + 1];
    end = begin + t.L < stop ? begin + t.L : stop;
    if (begin > stop) begin = end = stop;                 // the single task of an empty bucket
}

ZK_D void msm_accumulate_task_thread(const MsmArgs& a, const MsmTaskArgs& t, uint32_t task) {
    if (task >= *t.ntasks_ptr) return;
    uint32_t b, e;
    msm_task_range(t, task, b, e);
    xyzz_t acc = xyzz_identity();
    for (uint32_t k = b; k < e; ++k) {
        uint32_t en = a.entries[k];
        xyzz_madd(acc, a.bases[en >> 1], en & 1);
    }
    t.out[task] = acc;
}

ZK_D void msm_combine_task_thread(const MsmTaskArgs& t, uint32_t task) {
    if (task >= *t.ntasks_ptr) return;
    uint32_t b, e;
    msm_task_range(t, task, b, e);
    xyzz_t acc = xyzz_identity();
    for (uint32_t k = b; k < e; ++k) xyzz_add(acc, t.in[k]);
    t.out[task] = acc;
}

// last level: bucket b = sum of its (few) level-2 partials
ZK_D void msm_combine_bucket_thread(const MsmTaskArgs& t, uint32_t b) {
    if (b >= t.nb) return;
    xyzz_t acc = xyzz_identity();
    for (uint32_t k = t.seg_off[b]; k < t.seg_off[b + 1]; ++k) xyzz_add(acc, t.in[k]);
    t.out[b] = acc;
}

// gid -> (window j, chunk t).  Chunk covers bucket indices [b0, b0 + m), bucket b has
// multiplier b + 1:  sum (b+1) B_b = sum (b - b0 + 1) B_b  +  b0 * sum B_b.
ZK_D void msm_reduce_thread(const MsmArgs& a, uint32_t gid) {
    if (gid >= (a.nwin << a.log_t)) return;
    uint32_t j = gid >> a.log_t, t = gid & ((1u << a.log_t) - 1);
    uint32_t m = (1u << (a.c - 1)) >> a.log_t, b0 = t * m;
    const xyzz_t* B = a.buckets + ((size_t)j << (a.c - 1));
    xyzz_t running = xyzz_identity(), acc = xyzz_identity();
    for (uint32_t b = b0 + m; b-- > b0;) {
        xyzz_add(running, B[b]);
        xyzz_add(acc, running);
    }
    if (b0 != 0) { xyzz_t s = xyzz_mul_small(running, b0); xyzz_add(acc, s); }
    a.partials[gid] = acc;
}

// Fixed-base mode: sum_v v * B_v over ONE bucket set (v = b + 1 <= 2^(c-1)) by bit decomposition:
//   sum_v v B_v = sum_{t < c} 2^t * S_t,   S_t = sum of the buckets whose multiplier has bit t set.
// gid -> (bit t, chunk); every S_t is a plain parallel sum (no serial running-sum chain), the c
// values go back to the host which applies the powers of two (c - 1 doublings).
// The multipliers with bit t set, in increasing order: the s-th is v = (s >> t) * 2^(t+1) + 2^t + (s mod 2^t); there are
// 2^(c-2) of them below 2^(c-1) for t < c - 1, and the single v = 2^(c-1) for t = c - 1.  Chunk `chunk` of 2^log_t sums an equal
// share of that list, so every lane of a warp has the same number of additions (walking all buckets and testing the bit left
// half the lanes of a warp idle for every t above the chunk length).
// `counts` (the histogram of the same bucket set) lets an empty bucket cost a 4-byte read instead of a 128-byte one: the witness
// columns of a padded circuit fill a few hundred of the 2^(c-1) buckets.
ZK_D xyzz_t msm_reduce_bits_chunk(const xyzz_t* B, const uint32_t* counts, uint32_t c, uint32_t log_t, uint32_t t, uint32_t chunk) {
    const uint32_t nset = t == c - 1 ? 1u : (1u << (c - 2));
    uint32_t per = c >= 2 ? ((1u << (c - 2)) >> log_t) : 1u;
    if (per == 0) per = 1;
    const uint32_t s0 = chunk * per;
    xyzz_t acc = xyzz_identity();
    for (uint32_t s = s0; s < s0 + per && s < nset; ++s) {
        const uint32_t v = ((s >> t) << (t + 1)) + (1u << t) + (s & ((1u << t) - 1u));
        if (counts[v - 1] == 0) continue;
        xyzz_add(acc, B[v - 1]);
    }
    return acc;
}
ZK_D void msm_reduce_bits_thread(const MsmArgs& a, uint32_t gid) {
    if (gid >= (a.c << a.log_t)) return;
    uint32_t t = gid >> a.log_t, chunk = gid & ((1u << a.log_t) - 1);
    a.partials[gid] = msm_reduce_bits_chunk(a.buckets, a.counts, a.c, a.log_t, t, chunk);
}

// One block per window: sum the 2^log_t chunk results.  sm: nthreads xyzz_t.
ZK_D void msm_fold_block(const MsmArgs& a, uint32_t j, uint32_t nthreads, xyzz_t* sm) {
    const uint32_t T = 1u << a.log_t;
    ZK_PHASE_BEGIN(tid, nthreads)
    xyzz_t acc = xyzz_identity();
    for (uint32_t t = tid; t < T; t += nthreads) xyzz_add(acc, a.partials[((size_t)j << a.log_t) + t]);
    sm[tid] = acc;
    ZK_PHASE_END
    for (uint32_t s = nthreads >> 1; s > 0; s >>= 1) {
        ZK_PHASE_BEGIN(tid, nthreads)
        if (tid < s) { xyzz_t v = sm[tid]; xyzz_add(v, sm[tid + s]); sm[tid] = v; }
        ZK_PHASE_END
    }
    ZK_PHASE_BEGIN(tid, nthreads)
    if (tid == 0) a.window_sums[j] = sm[0];
    ZK_PHASE_END
}

}  // namespace b200zk
