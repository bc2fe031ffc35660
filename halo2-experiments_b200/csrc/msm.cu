// Launch side of the Pippenger MSM (see msm.cuh for the pipeline and the reference mapping).
#include "context.hpp"
#include "msm.cuh"

namespace b200zk {

static constexpr uint32_t MSM_DIGIT_THREADS = 256;
static constexpr uint32_t MSM_ACC_THREADS = 128;
static constexpr uint32_t MSM_SCAN_THREADS = 256;
static constexpr uint32_t MSM_FOLD_THREADS = 128;

__global__ void __launch_bounds__(MSM_DIGIT_THREADS) msm_count_kernel(const MsmArgs a) {
    msm_count_thread(a, blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(MSM_DIGIT_THREADS) msm_scatter_kernel(const MsmArgs a) {
    msm_scatter_thread(a, blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(MSM_SCAN_THREADS) msm_scan_blocksum_kernel(const MsmArgs a, uint32_t* blocksums) {
    __shared__ uint32_t sm[2 * MSM_SCAN_THREADS];
    msm_scan_blocksum_block(a, blocksums, blockIdx.x, blockDim.x, sm);
}
__global__ void __launch_bounds__(MSM_SCAN_THREADS) msm_scan_top_kernel(const MsmArgs a, uint32_t* blocksums, uint32_t nblocks) {
    __shared__ uint32_t sm[2 * MSM_SCAN_THREADS];
    msm_scan_top_block(a, blocksums, nblocks, blockDim.x, sm);
}
__global__ void __launch_bounds__(MSM_SCAN_THREADS) msm_scan_final_kernel(const MsmArgs a, const uint32_t* blocksums) {
    __shared__ uint32_t sm[2 * MSM_SCAN_THREADS];
    msm_scan_final_block(a, blocksums, blockIdx.x, blockDim.x, sm);
}
__global__ void __launch_bounds__(MSM_ACC_THREADS) msm_accumulate_kernel(const MsmArgs a) {
    msm_accumulate_thread(a, blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(MSM_ACC_THREADS) msm_reduce_kernel(const MsmArgs a) {
    msm_reduce_thread(a, blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(MSM_FOLD_THREADS) msm_fold_kernel(const MsmArgs a) {
    __shared__ xyzz_t sm[MSM_FOLD_THREADS];
    msm_fold_block(a, blockIdx.x, blockDim.x, sm);
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int32_t msm_run(b200zk_ctx* ctx, const fe_t* d_scalars, const affine_t* d_bases, size_t n, host::HAffine* out) {
    if (n == 0) { *out = {host::HFq::zero(), host::HFq::zero()}; return B200ZK_OK; }
    if (n >= ((size_t)1 << 31)) return fail(ctx, B200ZK_EINVAL, "msm_run", "len must be < 2^31");
    MsmShape s = msm_plan_shape(n, ctx->msm_force_c);
    // workspace layout
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    size_t o_counts = take(s.nbuckets * 4), o_offsets = take((s.nbuckets + 1) * 4), o_cursor = take(s.nbuckets * 4);
    size_t o_entries = take(n * s.nwin * 4);
    size_t o_buckets = take(s.nbuckets * sizeof(xyzz_t));
    size_t o_partials = take(((size_t)s.nwin << s.log_t) * sizeof(xyzz_t));
    size_t o_wsum = take(s.nwin * sizeof(xyzz_t));
    const uint32_t scan_items = MSM_SCAN_THREADS * MSM_SCAN_PER_THREAD;
    const uint32_t scan_blocks = (uint32_t)((s.nbuckets + scan_items - 1) / scan_items);
    size_t o_bsums = take((size_t)scan_blocks * 4);
    ZK_TRY(ws_reserve(ctx, ctx->msm_ws, off));
    char* base = (char*)ctx->msm_ws.p;
    MsmArgs a{};
    a.scalars = d_scalars; a.bases = d_bases; a.n = (uint32_t)n;
    a.c = s.c; a.nwin = s.nwin; a.log_t = s.log_t;
    a.counts = (uint32_t*)(base + o_counts); a.offsets = (uint32_t*)(base + o_offsets); a.cursor = (uint32_t*)(base + o_cursor);
    a.entries = (uint32_t*)(base + o_entries);
    a.buckets = (xyzz_t*)(base + o_buckets); a.partials = (xyzz_t*)(base + o_partials); a.window_sums = (xyzz_t*)(base + o_wsum);

    cudaStream_t st = ctx->stream;
    ZK_CUDA(ctx, cudaMemsetAsync(a.counts, 0, s.nbuckets * 4, st));
    unsigned nb = (unsigned)((n + MSM_DIGIT_THREADS - 1) / MSM_DIGIT_THREADS);
    msm_count_kernel<<<nb, MSM_DIGIT_THREADS, 0, st>>>(a);
    uint32_t* bsums = (uint32_t*)(base + o_bsums);
    msm_scan_blocksum_kernel<<<scan_blocks, MSM_SCAN_THREADS, 0, st>>>(a, bsums);
    msm_scan_top_kernel<<<1, MSM_SCAN_THREADS, 0, st>>>(a, bsums, scan_blocks);
    msm_scan_final_kernel<<<scan_blocks, MSM_SCAN_THREADS, 0, st>>>(a, bsums);
    msm_scatter_kernel<<<nb, MSM_DIGIT_THREADS, 0, st>>>(a);
    msm_accumulate_kernel<<<(unsigned)((s.nbuckets + MSM_ACC_THREADS - 1) / MSM_ACC_THREADS), MSM_ACC_THREADS, 0, st>>>(a);
    size_t nred = (size_t)s.nwin << s.log_t;
    msm_reduce_kernel<<<(unsigned)((nred + MSM_ACC_THREADS - 1) / MSM_ACC_THREADS), MSM_ACC_THREADS, 0, st>>>(a);
    msm_fold_kernel<<<s.nwin, MSM_FOLD_THREADS, 0, st>>>(a);
    ctx->launches += 8;
    ZK_CUDA(ctx, cudaGetLastError());
    ZK_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, a.window_sums, s.nwin * sizeof(xyzz_t), cudaMemcpyDeviceToHost, st));
    ZK_CUDA(ctx, cudaStreamSynchronize(st));
    *out = msm_finish(ctx->pinned, s.nwin, s.c);
    return B200ZK_OK;
}

}  // namespace b200zk
