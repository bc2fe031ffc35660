// Lookup argument, permutation step: halo2_proofs v2023_02_02 plonk/lookup/prover.rs
// `permute_expression_pair` (reached from create_proof, /root/reference/src/circuits/utils.rs:40-48).
//
//   a' = sort(input[0..u])                      (numeric order of canonical values = `Ord for Fr`)
//   s' : first occurrence of a value in a' takes that value and removes one copy from the table
//        multiset; repeated rows take the leftover table values in ascending order, the highest
//        repeated row first (upstream: BTreeMap iteration + Vec::pop).
//
// On the device: 256-bit keys are sorted as four stable 64-bit radix passes (least significant
// limb first) carrying a row index; CUB's DeviceRadixSort / DeviceScan (shipped with the CUDA
// toolkit) do the 64-bit sort and the prefix sums — library calls on a step that is not on the
// critical path of the proof; the surrounding kernels are ours.  The result is a pure function of
// the two multisets, so it is bit-identical to the CPU path.
#include "context.hpp"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace b200zk {

static constexpr uint32_t LK_THREADS = 256;

// canonical values, plus the OR of every 32-bit limb over the column (limb_or[8]): lookup columns
// of real circuits hold bytes / small integers, and a radix pass over bits that are zero in every
// key is the identity
__global__ void lk_canonical_kernel(const fe_t* in, fe_t* out, uint32_t n, uint32_t* limb_or) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    fe_t v = Fr::zero();
    if (i < n) { v = Fr::from_mont(in[i]); out[i] = v; }
#pragma unroll
    for (int l = 0; l < 8; ++l) {
        uint32_t w = __reduce_or_sync(0xffffffffu, v.l[l]);
        if ((threadIdx.x & 31) == 0 && w) atomicOr(&limb_or[l], w);
    }
}
__global__ void lk_iota_kernel(uint32_t* idx, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = i;
}
__global__ void lk_limb_kernel(const fe_t* canon, const uint32_t* idx, uint32_t limb, unsigned long long* keys, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fe_t& v = canon[idx[i]];
    keys[i] = (unsigned long long)v.l[2 * limb] | ((unsigned long long)v.l[2 * limb + 1] << 32);
}
__global__ void lk_gather_kernel(const fe_t* src, const uint32_t* idx, fe_t* dst, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { fe_t v = src[idx[i]]; dst[i] = v; }
}

__device__ __forceinline__ int cmp256(const fe_t& a, const fe_t& b) {
    for (int i = 7; i >= 0; --i) { if (a.l[i] != b.l[i]) return a.l[i] < b.l[i] ? -1 : 1; }
    return 0;
}

// first[i] = 1 if a'[i] starts a run; for run starts, binary-search the sorted table and mark the
// first copy as taken; repeated[i] = 1 - first[i].
__global__ void lk_mark_kernel(const fe_t* a_sorted, const fe_t* t_sorted, uint32_t u, uint32_t* repeated, uint32_t* keep, uint32_t* err) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= u) return;
    fe_t v = a_sorted[i];
    bool first = (i == 0) || cmp256(v, a_sorted[i - 1]) != 0;
    repeated[i] = first ? 0u : 1u;
    if (!first) return;
    uint32_t lo = 0, hi = u;                        // lower_bound
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (cmp256(t_sorted[mid], v) < 0) lo = mid + 1; else hi = mid;
    }
    if (lo >= u || cmp256(t_sorted[lo], v) != 0) { atomicExch(err, 1u); return; }   // input value not in table
    keep[lo] = 0u;                                  // this table copy is consumed by the run start
}
__global__ void lk_fill_kernel(uint32_t* p, uint32_t v, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// leftover[rank] = table row index for kept rows
__global__ void lk_compact_kernel(const uint32_t* keep, const uint32_t* keep_scan, uint32_t* leftover, uint32_t u) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < u && keep[i]) leftover[keep_scan[i]] = i;
}
// s'[i] = a'[i] at run starts, else leftover[m - 1 - rank(i)]  (Montgomery values gathered through the sort index)
__global__ void lk_build_table_kernel(const fe_t* a_mont_sorted, const fe_t* t_mont_sorted, const uint32_t* repeated,
                                      const uint32_t* rep_scan, const uint32_t* leftover, uint32_t m, fe_t* s_out, uint32_t u) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= u) return;
    fe_t v;
    if (!repeated[i]) v = a_mont_sorted[i];
    else v = t_mont_sorted[leftover[m - 1 - rep_scan[i]]];
    s_out[i] = v;
}

static unsigned nb(size_t n) { return (unsigned)((n + LK_THREADS - 1) / LK_THREADS); }

struct SortScratch {
    fe_t* canon; fe_t* canon_sorted; uint32_t *idx_a, *idx_b; unsigned long long *key_a, *key_b; void* cub_tmp; size_t cub_bytes;
    uint32_t* limb_or;          // 8 device words
};

// sorts `in` (Montgomery, u rows) by canonical value; outputs Montgomery rows in sorted order and the
// canonical sorted keys.
static int32_t sort256(b200zk_ctx* ctx, const fe_t* in, uint32_t u, fe_t* out_mont, const SortScratch& s) {
    cudaStream_t st = ctx->stream;
    ZK_CUDA(ctx, cudaMemsetAsync(s.limb_or, 0, 32, st));
    lk_canonical_kernel<<<nb(u), LK_THREADS, 0, st>>>(in, s.canon, u, s.limb_or);
    lk_iota_kernel<<<nb(u), LK_THREADS, 0, st>>>(s.idx_a, u);
    ctx->launches += 2;
    uint32_t ors[8];
    ZK_CUDA(ctx, cudaMemcpyAsync(ors, s.limb_or, 32, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(ctx, cudaStreamSynchronize(st));
    uint32_t *ia = s.idx_a, *ib = s.idx_b;
    for (uint32_t limb = 0; limb < 4; ++limb) {
        const unsigned long long bits = (unsigned long long)ors[2 * limb] | ((unsigned long long)ors[2 * limb + 1] << 32);
        if (bits == 0) continue;                                   // every key has this limb zero: stable sort = identity
        int end_bit = 64;
        while (end_bit > 1 && !((bits >> (end_bit - 1)) & 1)) --end_bit;
        lk_limb_kernel<<<nb(u), LK_THREADS, 0, st>>>(s.canon, ia, limb, s.key_a, u);
        size_t bytes = s.cub_bytes;
        ZK_CUDA(ctx, cub::DeviceRadixSort::SortPairs(s.cub_tmp, bytes, s.key_a, s.key_b, ia, ib, (int)u, 0, end_bit, st));
        ctx->launches += 2;
        uint32_t* t = ia; ia = ib; ib = t;
    }
    lk_gather_kernel<<<nb(u), LK_THREADS, 0, st>>>(s.canon, ia, s.canon_sorted, u);
    lk_gather_kernel<<<nb(u), LK_THREADS, 0, st>>>(in, ia, out_mont, u);
    ctx->launches += 2;
    ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

// d_in, d_tab: compressed input / table expressions (n rows, Montgomery); the first `u` rows take part.
// d_pin, d_ptab receive a'[0..u) and s'[0..u); the caller appends the blinding rows.
// *err_flag (device word) is set to 1 if an input value is missing from the table.
int32_t lookup_permute_run(b200zk_ctx* ctx, const fe_t* d_in, const fe_t* d_tab, uint32_t u, fe_t* d_pin, fe_t* d_ptab, uint32_t* d_err) {
    if (u == 0) return B200ZK_OK;
    size_t cub_sort = 0, cub_scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_sort, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (uint32_t*)nullptr, (uint32_t*)nullptr, (int)u, 0, 64, ctx->stream);
    cub::DeviceScan::ExclusiveSum(nullptr, cub_scan, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)u, ctx->stream);
    size_t cub_bytes = std::max(cub_sort, cub_scan);
    // workspace
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
    size_t o_canon = take((size_t)u * 32), o_a_canon = take((size_t)u * 32), o_t_canon = take((size_t)u * 32), o_t_mont = take((size_t)u * 32);
    size_t o_idx_a = take((size_t)u * 4), o_idx_b = take((size_t)u * 4), o_key_a = take((size_t)u * 8), o_key_b = take((size_t)u * 8);
    size_t o_rep = take((size_t)u * 4), o_rep_scan = take((size_t)u * 4), o_keep = take((size_t)u * 4), o_keep_scan = take((size_t)u * 4);
    size_t o_left = take((size_t)u * 4), o_cub = take(cub_bytes), o_cnt = take(64);   // o_cnt: limb_or
    ZK_TRY(ws_reserve(ctx, ctx->lookup_ws, off));
    char* base = (char*)ctx->lookup_ws.p;
    SortScratch s;
    s.canon = (fe_t*)(base + o_canon);
    s.idx_a = (uint32_t*)(base + o_idx_a); s.idx_b = (uint32_t*)(base + o_idx_b);
    s.key_a = (unsigned long long*)(base + o_key_a); s.key_b = (unsigned long long*)(base + o_key_b);
    s.cub_tmp = base + o_cub; s.cub_bytes = cub_bytes; s.limb_or = (uint32_t*)(base + o_cnt);
    fe_t* a_canon = (fe_t*)(base + o_a_canon);
    fe_t* t_canon = (fe_t*)(base + o_t_canon);
    fe_t* t_mont = (fe_t*)(base + o_t_mont);
    uint32_t *rep = (uint32_t*)(base + o_rep), *rep_scan = (uint32_t*)(base + o_rep_scan);
    uint32_t *keep = (uint32_t*)(base + o_keep), *keep_scan = (uint32_t*)(base + o_keep_scan), *leftover = (uint32_t*)(base + o_left);
    cudaStream_t st = ctx->stream;

    s.canon_sorted = a_canon;
    ZK_TRY(sort256(ctx, d_in, u, d_pin, s));                       // a' (Montgomery) + canonical keys
    s.canon_sorted = t_canon;
    ZK_TRY(sort256(ctx, d_tab, u, t_mont, s));                     // sorted table
    lk_fill_kernel<<<nb(u), LK_THREADS, 0, st>>>(keep, 1u, u);
    lk_mark_kernel<<<nb(u), LK_THREADS, 0, st>>>(a_canon, t_canon, u, rep, keep, d_err);
    size_t bytes = cub_bytes;
    ZK_CUDA(ctx, cub::DeviceScan::ExclusiveSum(s.cub_tmp, bytes, rep, rep_scan, (int)u, st));
    bytes = cub_bytes;
    ZK_CUDA(ctx, cub::DeviceScan::ExclusiveSum(s.cub_tmp, bytes, keep, keep_scan, (int)u, st));
    lk_compact_kernel<<<nb(u), LK_THREADS, 0, st>>>(keep, keep_scan, leftover, u);
    ctx->launches += 5;
    // m = number of repeated rows = rep_scan[u-1] + rep[u-1]
    uint32_t tail[2];
    ZK_CUDA(ctx, cudaMemcpyAsync(&tail[0], rep_scan + (u - 1), 4, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(ctx, cudaMemcpyAsync(&tail[1], rep + (u - 1), 4, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(ctx, cudaStreamSynchronize(st));
    uint32_t m = tail[0] + tail[1];
    lk_build_table_kernel<<<nb(u), LK_THREADS, 0, st>>>(d_pin, t_mont, rep, rep_scan, leftover, m, d_ptab, u);
    ctx->launches++;
    ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

}  // namespace b200zk
