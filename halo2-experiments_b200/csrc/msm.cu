// Launch side of the Pippenger MSM (see msm.cuh for the pipeline and the reference mapping).
#include "context.hpp"
#include "msm.cuh"
#include <cmath>
#include <cstdlib>
#include <vector>

namespace b200zk {

static constexpr uint32_t MSM_DIGIT_THREADS = 256;
static constexpr uint32_t MSM_ACC_THREADS = 128;
static constexpr uint32_t MSM_SCAN_THREADS = 256;
static constexpr uint32_t MSM_FOLD_THREADS = 128;
// thread-per-bucket only while no bucket exceeds this many entries; beyond that the task-balanced
// path wins even for uniform scalars (Poisson bucket sizes leave warps waiting for their longest
// bucket: 5.09 ms -> 4.6 ms for a dense 2^20 MSM, profiles/r01_sweep_tunables.jsonl)
static constexpr uint32_t MSM_FAST_MAX = 16;

// Digit pass, one scalar per lane, the W windows walked in lockstep by the warp.  Same keys and
// entries as msm_count_thread / msm_scatter_thread (msm.cuh), with two warp-level shortcuts that
// matter for real witness columns:
//   * a warp whose 32 scalars are all zero leaves at once (advice columns of a padded circuit);
//   * when all 32 lanes hit the same bucket in a window — grand-product columns are 1 on every
//     unused row, sorted lookup columns are long runs of one value — one lane adds 32 to the
//     counter / claims 32 slots instead of 32 atomics serialising on one address
//     (1.1 ms -> count and 1.8 ms -> scatter for a z column of the Merkle Sum Tree circuit at k = 20).
template <bool SCATTER> __device__ __forceinline__ void msm_digit_pass_warp(const MsmArgs& a, uint32_t col, uint32_t i) {
    const unsigned FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31;
    const fe_t* scalars = a.batch > 1 ? a.scalars_tab[col] : a.scalars;
    fe_t s = i < a.n ? scalars[i] : Fr::zero();
    if (a.batch > 1) { if (a.subs_on[col] && i < a.n) s = Fr::sub(s, a.subs[col]); }
    else if (a.use_sub && i < a.n) s = Fr::sub(s, a.sub);
    if (!__any_sync(FULL, !Fr::is_zero(s))) return;
    const uint32_t key_base = col * a.set_buckets;
    s = Fr::from_mont(s);
    uint32_t carry = 0;
    const uint32_t half = 1u << (a.c - 1);
    for (uint32_t j = 0; j < a.nwin; ++j) {
        uint32_t d = msm_window_bits(s, j * a.c, a.c) + carry;
        uint32_t sign = 0;
        if (d > half) { d = (1u << a.c) - d; sign = 1; carry = 1; } else carry = 0;
        const uint32_t key = d == 0 ? 0xffffffffu : key_base + (a.pre ? d - 1 : j * half + d - 1);
        const uint32_t entry = ((a.pre ? j * a.pre_stride + i : i) << 1) | sign;
        const uint32_t k0 = __shfl_sync(FULL, key, 0);
        if (__all_sync(FULL, key == k0)) {
            if (k0 == 0xffffffffu) continue;
            if (SCATTER) {
                uint32_t pos = 0;
                if (lane == 0) pos = atomicAdd(&a.cursor[k0], 32u);
                pos = __shfl_sync(FULL, pos, 0);
                a.entries[pos + lane] = entry;
            } else if (lane == 0) {
                atomicAdd(&a.counts[k0], 32u);
            }
        } else if (key != 0xffffffffu) {
            if (SCATTER) a.entries[atomicAdd(&a.cursor[key], 1u)] = entry;
            else atomicAdd(&a.counts[key], 1u);
        }
    }
}
// grid = batch * blocks_per_col blocks: a block never straddles two columns
__global__ void __launch_bounds__(MSM_DIGIT_THREADS) msm_count_kernel(const MsmArgs a, uint32_t blocks_per_col) {
    msm_digit_pass_warp<false>(a, blockIdx.x / blocks_per_col, (blockIdx.x % blocks_per_col) * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(MSM_DIGIT_THREADS) msm_scatter_kernel(const MsmArgs a, uint32_t blocks_per_col) {
    msm_digit_pass_warp<true>(a, blockIdx.x / blocks_per_col, (blockIdx.x % blocks_per_col) * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(MSM_SCAN_THREADS) scan_blocksum_kernel(const ScanArgs s) {
    __shared__ uint32_t sm[2 * MSM_SCAN_THREADS];
    scan_blocksum_block(s, blockIdx.x, blockDim.x, sm);
}
__global__ void __launch_bounds__(MSM_SCAN_THREADS) scan_top_kernel(const ScanArgs s, uint32_t nblocks) {
    __shared__ uint32_t sm[2 * MSM_SCAN_THREADS];
    scan_top_block(s, nblocks, blockDim.x, sm);
}
__global__ void __launch_bounds__(MSM_SCAN_THREADS) scan_final_kernel(const ScanArgs s) {
    __shared__ uint32_t sm[2 * MSM_SCAN_THREADS];
    scan_final_block(s, blockIdx.x, blockDim.x, sm);
}
__global__ void __launch_bounds__(MSM_ACC_THREADS) msm_accumulate_kernel(const MsmArgs a) {
    msm_accumulate_thread(a, blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(MSM_ACC_THREADS) msm_accumulate_task_kernel(const MsmArgs a, const MsmTaskArgs t) {
    msm_accumulate_task_thread(a, t, blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(MSM_ACC_THREADS) msm_combine_task_kernel(const MsmTaskArgs t) {
    msm_combine_task_thread(t, blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(MSM_ACC_THREADS) msm_combine_bucket_kernel(const MsmTaskArgs t) {
    msm_combine_bucket_thread(t, blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(MSM_ACC_THREADS) msm_reduce_kernel(const MsmArgs a) {
    msm_reduce_thread(a, blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(MSM_ACC_THREADS) msm_reduce_bits_kernel(const MsmArgs a) {
    msm_reduce_bits_thread(a, blockIdx.x * blockDim.x + threadIdx.x);
}
// reduce_bits with the first 7 levels of the fold inside the block: the 128 chunk sums of a block
// are tree-added in shared memory and ONE partial per block is written, so the fold that follows
// (c blocks, latency bound) adds 2^(log_t - 7) values per bit instead of 2^log_t.
__global__ void __launch_bounds__(MSM_ACC_THREADS) msm_reduce_bits_tree_kernel(const MsmArgs a) {
    __shared__ xyzz_t sm[MSM_ACC_THREADS];
    const uint32_t blocks_per_set = (a.c << a.log_t) / MSM_ACC_THREADS;
    const uint32_t col = blockIdx.x / blocks_per_set;
    const uint32_t gid = (blockIdx.x % blocks_per_set) * blockDim.x + threadIdx.x;     // (bit t, chunk) inside the column
    const uint32_t t = gid >> a.log_t, chunk = gid & ((1u << a.log_t) - 1);
    sm[threadIdx.x] = msm_reduce_bits_chunk(a.buckets + (size_t)col * a.set_buckets, a.counts + (size_t)col * a.set_buckets, a.c, a.log_t, t, chunk);
    __syncthreads();
    for (uint32_t s = MSM_ACC_THREADS >> 1; s > 0; s >>= 1) {
        if (threadIdx.x < s) { xyzz_t v = sm[threadIdx.x]; xyzz_add(v, sm[threadIdx.x + s]); sm[threadIdx.x] = v; }
        __syncthreads();
    }
    if (threadIdx.x == 0) a.partials[blockIdx.x] = sm[0];          // index = (col * c + t) * 2^(log_t - 7) + block within the bit
}
__global__ void __launch_bounds__(MSM_FOLD_THREADS) msm_fold_kernel(const MsmArgs a) {
    __shared__ xyzz_t sm[MSM_FOLD_THREADS];
    msm_fold_block(a, blockIdx.x, blockDim.x, sm);
}

// fixed-base table: next window = 2^c * previous window (c doublings), normalised to affine
__global__ void __launch_bounds__(MSM_ACC_THREADS) msm_pre_shift_kernel(const affine_t* prev, size_t n, uint32_t c, xyzz_t* out, fe_t* zzz) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    affine_t p = prev[i];
    xyzz_t acc = xyzz_from_affine(p);
    for (uint32_t k = 0; k < c; ++k) acc = xyzz_dbl(acc);
    out[i] = acc;
    zzz[i] = acc.zzz;
}
__global__ void __launch_bounds__(MSM_ACC_THREADS) msm_pre_normalize_kernel(const xyzz_t* in, const fe_t* zzz_inv, size_t n, affine_t* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    xyzz_t p = in[i];
    fe_t iz = zzz_inv[i];
    affine_t r;
    if (Fq::is_zero(p.zz)) { r.x = Fq::zero(); r.y = Fq::zero(); }
    else { fe_t t = Fq::mul(p.zz, iz); r.x = Fq::mul(p.x, Fq::sqr(t)); r.y = Fq::mul(p.y, iz); }
    out[i] = r;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static unsigned nb(size_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

static void run_scan(b200zk_ctx* ctx, ScanArgs s) {
    const uint32_t items = MSM_SCAN_THREADS * MSM_SCAN_PER_THREAD;
    const uint32_t blocks = (s.total + items - 1) / items;
    scan_blocksum_kernel<<<blocks, MSM_SCAN_THREADS, 0, ctx->stream>>>(s);
    scan_top_kernel<<<1, MSM_SCAN_THREADS, 0, ctx->stream>>>(s, blocks);
    scan_final_kernel<<<blocks, MSM_SCAN_THREADS, 0, ctx->stream>>>(s);
    ctx->launches += 3;
}

// table[j*n + i] = 2^(c j) * bases[i] for j < nwin; table[0..n) is a copy of the bases.
int32_t msm_precompute_run(b200zk_ctx* ctx, const affine_t* d_bases, size_t n, uint32_t c, uint32_t nwin, affine_t* d_table) {
    cudaStream_t st = ctx->stream;
    ZK_CUDA(ctx, cudaMemcpyAsync(d_table, d_bases, n * sizeof(affine_t), cudaMemcpyDeviceToDevice, st));
    ZK_TRY(ws_reserve(ctx, ctx->setup_ws, n * (sizeof(xyzz_t) + sizeof(fe_t))));
    xyzz_t* xyzz = (xyzz_t*)ctx->setup_ws.p;
    fe_t* zzz = (fe_t*)(xyzz + n);
    for (uint32_t j = 1; j < nwin; ++j) {
        msm_pre_shift_kernel<<<nb(n, MSM_ACC_THREADS), MSM_ACC_THREADS, 0, st>>>(d_table + (size_t)(j - 1) * n, n, c, xyzz, zzz);
        ctx->launches++;
        ZK_TRY(batch_invert_run(ctx, zzz, n, 1));
        msm_pre_normalize_kernel<<<nb(n, MSM_ACC_THREADS), MSM_ACC_THREADS, 0, st>>>(xyzz, zzz, n, d_table + (size_t)j * n);
        ctx->launches++;
    }
    ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

// `ncols` columns of n scalars each against the same bases: d_bases = the n points (pre == null,
// ncols = 1), or the fixed-base table of a params object with `pre->stride` points per window
// (pre != null; then n <= stride).  One launch sequence for all columns: the latency-bound tail
// of a commit (scans, reduce, fold: ~0.25 ms for an almost empty column) is paid once per batch,
// and its kernels get ncols times the parallelism.  subs[b] (may be null) = constant subtracted
// from column b's scalars (params_commit_run).
int32_t msm_run_multi(b200zk_ctx* ctx, const fe_t* const* d_cols, uint32_t ncols, const affine_t* d_bases, size_t n, const MsmPre* pre,
                      host::HAffine* outs, const fe_t* const* subs) {
    if (ncols == 0) return B200ZK_OK;
    if (n == 0) { for (uint32_t b = 0; b < ncols; ++b) outs[b] = {host::HFq::zero(), host::HFq::zero()}; return B200ZK_OK; }
    if (n >= ((size_t)1 << 31)) return fail(ctx, B200ZK_EINVAL, "msm_run", "len must be < 2^31");
    MsmShape s = pre ? pre->shape : msm_plan_shape(n, ctx->msm_force_c);
    const uint32_t nsums = pre ? s.c : s.nwin;              // results per column handed to the host
    const bool tree = pre && s.log_t >= 7;
    if (ncols > 1 && (!tree || (size_t)ncols * nsums * sizeof(xyzz_t) > ((size_t)60 << 10) || (uint64_t)ncols * n * s.nwin >= ((uint64_t)1 << 32))) {
        for (uint32_t b = 0; b < ncols; ++b) ZK_TRY(msm_run_multi(ctx, d_cols + b, 1, d_bases, n, pre, outs + b, subs ? subs + b : nullptr));
        return B200ZK_OK;
    }
    const uint32_t NB = (uint32_t)s.nbuckets, B = NB * ncols;
    // workspace layout
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    size_t o_counts = take((size_t)B * 4 + 4), o_offsets = take(((size_t)B + 1) * 4), o_cursor = take((size_t)B * 4);
    size_t o_entries = take((size_t)ncols * n * s.nwin * 4);
    size_t o_buckets = take((size_t)B * sizeof(xyzz_t));
    size_t o_partials = take((((size_t)nsums * ncols) << s.log_t) * sizeof(xyzz_t));
    size_t o_wsum = take((size_t)nsums * ncols * sizeof(xyzz_t));
    const uint32_t scan_items = MSM_SCAN_THREADS * MSM_SCAN_PER_THREAD;
    size_t o_bsums = take((size_t)((B + scan_items - 1) / scan_items) * 4);
    size_t o_toff1 = take(((size_t)B + 1) * 4), o_toff2 = take(((size_t)B + 1) * 4);
    size_t o_tab = take((size_t)ncols * (sizeof(void*) + sizeof(fe_t) + 4));
    ZK_TRY(ws_reserve(ctx, ctx->msm_ws, off));
    char* base = (char*)ctx->msm_ws.p;
    MsmArgs a{};
    a.scalars = d_cols[0]; a.bases = d_bases; a.n = (uint32_t)n;
    a.c = s.c; a.nwin = s.nwin; a.log_t = s.log_t;
    a.pre = pre ? 1 : 0; a.pre_stride = pre ? pre->stride : 0; a.nbuckets = B;
    a.batch = ncols; a.set_buckets = NB;
    a.use_sub = (ncols == 1 && subs && subs[0]) ? 1 : 0;
    if (a.use_sub) a.sub = *subs[0];
    a.counts = (uint32_t*)(base + o_counts); a.offsets = (uint32_t*)(base + o_offsets); a.cursor = (uint32_t*)(base + o_cursor);
    a.entries = (uint32_t*)(base + o_entries);
    a.buckets = (xyzz_t*)(base + o_buckets); a.partials = (xyzz_t*)(base + o_partials); a.window_sums = (xyzz_t*)(base + o_wsum);
    uint32_t* d_max = a.counts + B;                         // one extra word after the histogram
    uint32_t* bsums = (uint32_t*)(base + o_bsums);
    cudaStream_t st = ctx->stream;
    if (ncols > 1) {                                        // column pointers, constants and flags, staged through pinned memory
        // layout (32-byte aligned first): constants [ncols] | column pointers [ncols] | flags [ncols]
        char* h = (char*)ctx->pinned + ((size_t)60 << 10);      // last 4 KB of the 64 KB staging buffer
        fe_t* hs = (fe_t*)h;
        const fe_t** hp = (const fe_t**)(h + ncols * sizeof(fe_t));
        uint32_t* hf = (uint32_t*)(h + ncols * (sizeof(fe_t) + sizeof(void*)));
        for (uint32_t b = 0; b < ncols; ++b) {
            hp[b] = d_cols[b];
            hf[b] = (subs && subs[b]) ? 1u : 0u;
            if (hf[b]) memcpy(&hs[b], subs[b], sizeof(fe_t)); else memset(&hs[b], 0, sizeof(fe_t));
        }
        ZK_CUDA(ctx, cudaMemcpyAsync(base + o_tab, h, (size_t)ncols * (sizeof(void*) + sizeof(fe_t) + 4), cudaMemcpyHostToDevice, st));
        a.subs = (const fe_t*)(base + o_tab);
        a.scalars_tab = (const fe_t* const*)(base + o_tab + ncols * sizeof(fe_t));
        a.subs_on = (const uint32_t*)(base + o_tab + ncols * (sizeof(fe_t) + sizeof(void*)));
    }

    const uint32_t blocks_per_col = nb(n, MSM_DIGIT_THREADS);
    ZK_CUDA(ctx, cudaMemsetAsync(a.counts, 0, (size_t)B * 4 + 4, st));
    msm_count_kernel<<<blocks_per_col * ncols, MSM_DIGIT_THREADS, 0, st>>>(a, blocks_per_col);
    ctx->launches++;
    run_scan(ctx, ScanArgs{a.counts, a.offsets, a.cursor, bsums, d_max, B, 0, 0});
    ZK_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, d_max, 4, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(ctx, cudaMemcpyAsync((char*)ctx->pinned + 4, a.offsets + B, 4, cudaMemcpyDeviceToHost, st));
    msm_scatter_kernel<<<blocks_per_col * ncols, MSM_DIGIT_THREADS, 0, st>>>(a, blocks_per_col);
    ctx->launches++;
    ZK_CUDA(ctx, cudaStreamSynchronize(st));
    const uint32_t maxcnt = ((const uint32_t*)ctx->pinned)[0];
    const size_t total_entries = ((const uint32_t*)ctx->pinned)[1];

    uint32_t fast_max = MSM_FAST_MAX;
    if (const char* e = getenv("B200ZK_MSM_FAST_MAX")) fast_max = (uint32_t)strtoul(e, nullptr, 10);
    // short columns are latency bound: shorter chains, one more level; long ones: 24 leaves a uniform 2^20 column (240 entries per
    // bucket) with 10 partials per bucket, summed directly (16: 129.2 ms per k = 20 proof, 24: 128.7, 32: 129.0, 48: 130.0).
    // (A warp per task of 256 entries with a shuffle tree was tried for the short columns: no gain, k = 14 proof 6.8 vs 6.4 ms.)
    uint32_t seg_min = n < ((size_t)1 << 18) ? 8 : 24;
    if (const char* e = getenv("B200ZK_MSM_SEG_MIN")) seg_min = (uint32_t)strtoul(e, nullptr, 10);
    if (maxcnt <= fast_max) {
        msm_accumulate_kernel<<<nb(B, MSM_ACC_THREADS), MSM_ACC_THREADS, 0, st>>>(a);
        ctx->launches++;
    } else {
        // Balanced levels of at most L items per task: entries -> partials (accumulate_task), then
        // partials -> partials (combine_task) until no bucket has more than L of them, then one
        // thread per bucket sums what is left.  A uniform 2^20 column takes two levels; the single
        // bucket holding the ~2^20 ones of a grand-product column takes five, every one of them
        // with >= maxcnt / L^level independent tasks (a cube-root split left 10^4 threads with
        // 100-long serial chains: 2.85 ms for one such bucket).
        uint32_t L = seg_min < 2 ? 2 : seg_min;
        // level k has at most entries / L^k + B * (1 + 1/L + ...) tasks: both ping-pong buffers hold entries/L + 2B
        const size_t t1_bound = total_entries / L + B;
        const size_t pcap = t1_bound + B;
        ZK_TRY(ws_reserve(ctx, ctx->msm_ws2, 2 * pcap * sizeof(xyzz_t) + (size_t)(B + 1) * 4 + 1024));
        xyzz_t* pbuf[2] = {(xyzz_t*)ctx->msm_ws2.p, (xyzz_t*)ctx->msm_ws2.p + pcap};
        uint32_t* toff[3] = {(uint32_t*)(base + o_toff1), (uint32_t*)(base + o_toff2), (uint32_t*)(pbuf[1] + pcap)};
        run_scan(ctx, ScanArgs{a.counts, toff[0], nullptr, bsums, nullptr, B, 0, L});
        MsmTaskArgs t1{a.offsets, toff[0], B, toff[0] + B, L, nullptr, pbuf[0]};
        msm_accumulate_task_kernel<<<nb(t1_bound, MSM_ACC_THREADS), MSM_ACC_THREADS, 0, st>>>(a, t1);
        ctx->launches++;
        uint32_t cur = 0, cur_t = 0;                            // partials in pbuf[cur], their offsets in toff[cur_t]
        uint64_t m = ((uint64_t)maxcnt + L - 1) / L;            // most partials any bucket holds
        size_t bound = t1_bound;
        while (m > L) {
            const uint32_t nxt_t = (cur_t + 1) % 3, nxt = cur ^ 1;
            bound = bound / L + B;                              // <= pcap, see above
            run_scan(ctx, ScanArgs{toff[cur_t], toff[nxt_t], nullptr, bsums, nullptr, B, 1, L});
            MsmTaskArgs t2{toff[cur_t], toff[nxt_t], B, toff[nxt_t] + B, L, pbuf[cur], pbuf[nxt]};
            msm_combine_task_kernel<<<nb(bound, MSM_ACC_THREADS), MSM_ACC_THREADS, 0, st>>>(t2);
            ctx->launches++;
            cur = nxt; cur_t = nxt_t;
            m = (m + L - 1) / L;
        }
        MsmTaskArgs t3{toff[cur_t], nullptr, B, nullptr, 0, pbuf[cur], a.buckets};
        msm_combine_bucket_kernel<<<nb(B, MSM_ACC_THREADS), MSM_ACC_THREADS, 0, st>>>(t3);
        ctx->launches++;
    }
    size_t nred = ((size_t)nsums << s.log_t) * ncols;
    if (tree) {
        msm_reduce_bits_tree_kernel<<<(unsigned)(nred / MSM_ACC_THREADS), MSM_ACC_THREADS, 0, st>>>(a);
        MsmArgs af = a;
        af.log_t = s.log_t - 7;                                  // partials per bit after the in-block tree
        msm_fold_kernel<<<nsums * ncols, af.log_t >= 5 ? 32 : (1u << af.log_t), 0, st>>>(af);
    } else {
        if (pre) msm_reduce_bits_kernel<<<nb(nred, MSM_ACC_THREADS), MSM_ACC_THREADS, 0, st>>>(a);
        else msm_reduce_kernel<<<nb(nred, MSM_ACC_THREADS), MSM_ACC_THREADS, 0, st>>>(a);
        msm_fold_kernel<<<nsums, MSM_FOLD_THREADS, 0, st>>>(a);
    }
    ctx->launches += 2;
    ZK_CUDA(ctx, cudaGetLastError());
    ZK_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, a.window_sums, (size_t)nsums * ncols * sizeof(xyzz_t), cudaMemcpyDeviceToHost, st));
    ZK_CUDA(ctx, cudaStreamSynchronize(st));
    std::vector<host::HXyzz> sums(ncols);
    for (uint32_t b = 0; b < ncols; ++b) {
        const char* ws = (const char*)ctx->pinned + (size_t)b * nsums * sizeof(xyzz_t);
        sums[b] = pre ? msm_finish_bits(ws, s.c) : msm_finish(ws, s.nwin, s.c);
    }
    host::hx_batch_to_affine(sums.data(), ncols, outs);          // one inversion per batch of commitments
    return B200ZK_OK;
}

int32_t msm_run_ex(b200zk_ctx* ctx, const fe_t* d_scalars, const affine_t* d_bases, size_t n, const MsmPre* pre, host::HAffine* out,
                   const fe_t* sub) {
    return msm_run_multi(ctx, &d_scalars, 1, d_bases, n, pre, out, sub ? &sub : nullptr);
}

int32_t msm_run(b200zk_ctx* ctx, const fe_t* d_scalars, const affine_t* d_bases, size_t n, host::HAffine* out) {
    return msm_run_ex(ctx, d_scalars, d_bases, n, nullptr, out, nullptr);
}

}  // namespace b200zk
