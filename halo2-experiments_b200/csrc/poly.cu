// Launch side of the O(n) field-vector primitives (poly.cuh).
#include "context.hpp"
#include "poly.cuh"

namespace b200zk {

static constexpr uint32_t POLY_THREADS = 128;
static constexpr uint32_t SCAN_THREADS = 512;
// elements per thread of the chunked scans: 64 keeps the single-block carry scan short on long columns; a short column is one
// thread's chain of dependent multiplications, so it gets 16
static size_t chunk_for(size_t n) { return n <= ((size_t)1 << 16) ? 16 : 64; }

template <class F> __global__ void __launch_bounds__(POLY_THREADS) batch_invert_kernel(fe_t* a, fe_t* scratch, size_t n, size_t lanes) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < lanes) batch_invert_lane<F>(a, scratch, n, t, lanes);
}

__global__ void __launch_bounds__(POLY_THREADS) recur_local_kernel(const fe_t* a, fe_t* y, size_t n, size_t m, const fe_t b, fe_t* heads) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    recur_local_chunk(a, y, n, m, c, b, heads);
}
__global__ void __launch_bounds__(SCAN_THREADS) recur_carries_kernel(fe_t* heads, fe_t* carries, size_t n, size_t m, const fe_t b) {
    __shared__ fe_t sm[2 * SCAN_THREADS];
    recur_carries_block(heads, carries, n, m, b, blockDim.x, sm);
}
__global__ void __launch_bounds__(POLY_THREADS) recur_apply_kernel(fe_t* y, size_t n, size_t m, const fe_t b, const fe_t* carries) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    recur_apply_chunk(y, n, m, c, b, carries);
}

__global__ void __launch_bounds__(POLY_THREADS) prodscan_product_kernel(const fe_t* p, size_t n, size_t m, fe_t* prods) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    prodscan_chunk_product(p, n, m, c, prods);
}
__global__ void __launch_bounds__(SCAN_THREADS) prodscan_carries_kernel(fe_t* prods, size_t C, const fe_t z0) {
    __shared__ fe_t sm[2 * SCAN_THREADS];
    prodscan_carries_block(prods, C, z0, blockDim.x, sm);
}
__global__ void __launch_bounds__(POLY_THREADS) prodscan_write_kernel(const fe_t* p, fe_t* z, size_t n, size_t m, const fe_t* prods) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    prodscan_write_chunk(p, z, n, m, c, prods);
}

static fe_t to_dev(const host::HFr& x) { fe_t r; memcpy(r.l, x.v, 32); return r; }
static unsigned nblocks(size_t work, unsigned threads) { return (unsigned)((work + threads - 1) / threads); }

int32_t batch_invert_run(b200zk_ctx* ctx, fe_t* d_a, size_t n, int field) {
    if (n == 0) return B200ZK_OK;
    ZK_TRY(ws_reserve(ctx, ctx->poly_ws, n * sizeof(fe_t)));
    // one inversion (Field::inv_gcd, about 40 multiplications' worth of instructions) per lane of 16 elements — 8 for a
    // short column, where the kernel is one lane's latency; at most what fills the machine
    const size_t per_lane = n <= ((size_t)1 << 16) ? 8 : 16;
    size_t lanes = std::min<size_t>((n + per_lane - 1) / per_lane, (size_t)ctx->sm_count * 1024);
    if (lanes == 0) lanes = 1;
    if (field == 0) batch_invert_kernel<Fr><<<nblocks(lanes, POLY_THREADS), POLY_THREADS, 0, ctx->stream>>>(d_a, (fe_t*)ctx->poly_ws.p, n, lanes);
    else batch_invert_kernel<Fq><<<nblocks(lanes, POLY_THREADS), POLY_THREADS, 0, ctx->stream>>>(d_a, (fe_t*)ctx->poly_ws.p, n, lanes);
    ctx->launches++;
    ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

// y[i] = a[i] + b*y[i+1]; d_y may be null (only the head y[0] is wanted) or alias d_a.
// The head is left in ctx->poly_heads[0] on the device; *head_out (host) is filled if non-null.
int32_t recurrence_run(b200zk_ctx* ctx, const fe_t* d_a, fe_t* d_y, size_t n, const host::HFr& b, host::HFr* head_out) {
    if (n == 0) { if (head_out) *head_out = host::HFr::zero(); return B200ZK_OK; }
    const size_t CHUNK = chunk_for(n);
    size_t C = (n + CHUNK - 1) / CHUNK;
    ZK_TRY(ws_reserve(ctx, ctx->poly_heads, 2 * C * sizeof(fe_t)));
    fe_t* heads = (fe_t*)ctx->poly_heads.p;
    fe_t* carries = heads + C;
    fe_t bd = to_dev(b);
    recur_local_kernel<<<nblocks(C, POLY_THREADS), POLY_THREADS, 0, ctx->stream>>>(d_a, d_y, n, CHUNK, bd, heads);
    recur_carries_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(heads, carries, n, CHUNK, bd);
    ctx->launches += 2;
    if (d_y) {
        recur_apply_kernel<<<nblocks(C, POLY_THREADS), POLY_THREADS, 0, ctx->stream>>>(d_y, n, CHUNK, bd, carries);
        ctx->launches++;
    }
    ZK_CUDA(ctx, cudaGetLastError());
    if (head_out) {
        ZK_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, heads, sizeof(fe_t), cudaMemcpyDeviceToHost, ctx->stream));
        ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *head_out = host::HFr::from_limbs(ctx->pinned);
    }
    return B200ZK_OK;
}

// ---- batched Horner: Q independent (polynomial, point) evaluations in two launches -------------
__global__ void __launch_bounds__(POLY_THREADS) recur_local_batch_kernel(const fe_t* const* polys, const fe_t* points, size_t n, size_t m,
                                                                         size_t C, fe_t* heads) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t q = blockIdx.y;
    if (c >= C) return;
    fe_t b = points[q];
    recur_local_chunk(polys[q], nullptr, n, m, c, b, heads + q * C);
}
__global__ void __launch_bounds__(SCAN_THREADS) recur_carries_batch_kernel(fe_t* heads, fe_t* carries, size_t n, size_t m, size_t C, const fe_t* points) {
    __shared__ fe_t sm[2 * SCAN_THREADS];
    size_t q = blockIdx.x;
    fe_t b = points[q];
    recur_carries_block(heads + q * C, carries + q * C, n, m, b, blockDim.x, sm);
}

// out[q] = polys[q](points[q]) for Q polynomials of n coefficients (device pointers in a host array)
int32_t eval_batch_run(b200zk_ctx* ctx, const fe_t* const* h_polys, const host::HFr* h_points, size_t Q, size_t n, host::HFr* out) {
    if (Q == 0) return B200ZK_OK;
    if (n == 0) { for (size_t q = 0; q < Q; ++q) out[q] = host::HFr::zero(); return B200ZK_OK; }
    const size_t CHUNK = chunk_for(n);
    const size_t C = (n + CHUNK - 1) / CHUNK;
    size_t bytes = 2 * Q * C * sizeof(fe_t) + Q * (sizeof(fe_t) + sizeof(void*)) + 512;
    ZK_TRY(ws_reserve(ctx, ctx->poly_batch, bytes));
    fe_t* heads = (fe_t*)ctx->poly_batch.p;
    fe_t* carries = heads + Q * C;
    fe_t* d_points = carries + Q * C;
    const fe_t** d_polys = (const fe_t**)(d_points + Q);
    cudaStream_t st = ctx->stream;
    ZK_CUDA(ctx, cudaMemcpyAsync(d_points, h_points, Q * sizeof(fe_t), cudaMemcpyHostToDevice, st));
    ZK_CUDA(ctx, cudaMemcpyAsync(d_polys, h_polys, Q * sizeof(void*), cudaMemcpyHostToDevice, st));
    dim3 grid(nblocks(C, POLY_THREADS), (unsigned)Q);
    recur_local_batch_kernel<<<grid, POLY_THREADS, 0, st>>>(d_polys, d_points, n, CHUNK, C, heads);
    recur_carries_batch_kernel<<<(unsigned)Q, SCAN_THREADS, 0, st>>>(heads, carries, n, CHUNK, C, d_points);
    ctx->launches += 2;
    ZK_CUDA(ctx, cudaGetLastError());
    std::vector<fe_t> tmp(Q);
    ZK_CUDA(ctx, cudaMemcpy2DAsync(tmp.data(), sizeof(fe_t), heads, C * sizeof(fe_t), sizeof(fe_t), Q, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(ctx, cudaStreamSynchronize(st));
    for (size_t q = 0; q < Q; ++q) out[q] = host::HFr::from_limbs(tmp[q].l);
    return B200ZK_OK;
}

__global__ void __launch_bounds__(POLY_THREADS) pow_table_kernel(fe_t* out, const fe_t base, uint32_t count, uint32_t shift) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = Fr::pow_u64(base, (unsigned long long)i << shift);
}
__global__ void __launch_bounds__(POLY_THREADS) pow_combine_kernel(fe_t* out, size_t n, const fe_t* lo, const fe_t* hi, uint32_t bits) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe_t a = lo[i & ((1u << bits) - 1)], b = hi[i >> bits];
    out[i] = Fr::mul(a, b);
}

// out[i] = base^i: base^i = lo[i mod 2^b] * hi[i >> b]
int32_t powers_run(b200zk_ctx* ctx, const host::HFr& base, size_t n, fe_t* d_out) {
    if (n == 0) return B200ZK_OK;
    uint32_t lg = 0; while (((size_t)1 << lg) < n) ++lg;
    uint32_t bits = (lg + 1) / 2;
    size_t n_lo = (size_t)1 << bits, n_hi = ((n - 1) >> bits) + 1;
    ZK_TRY(ws_reserve(ctx, ctx->poly_heads, (n_lo + n_hi) * sizeof(fe_t)));
    fe_t* lo = (fe_t*)ctx->poly_heads.p;
    fe_t* hi = lo + n_lo;
    pow_table_kernel<<<nblocks(n_lo, POLY_THREADS), POLY_THREADS, 0, ctx->stream>>>(lo, to_dev(base), (uint32_t)n_lo, 0);
    pow_table_kernel<<<nblocks(n_hi, POLY_THREADS), POLY_THREADS, 0, ctx->stream>>>(hi, to_dev(base), (uint32_t)n_hi, bits);
    pow_combine_kernel<<<nblocks(n, POLY_THREADS), POLY_THREADS, 0, ctx->stream>>>(d_out, n, lo, hi, bits);
    ctx->launches += 3;
    ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

// z[0] = z0, z[i] = z[i-1] * p[i-1] for i < n  (z may alias p)
int32_t prefix_product_run(b200zk_ctx* ctx, const fe_t* d_p, fe_t* d_z, size_t n, const host::HFr& z0) {
    if (n == 0) return B200ZK_OK;
    const size_t CHUNK = chunk_for(n);
    size_t C = (n + CHUNK - 1) / CHUNK;
    ZK_TRY(ws_reserve(ctx, ctx->poly_heads, 2 * C * sizeof(fe_t)));
    fe_t* prods = (fe_t*)ctx->poly_heads.p;
    prodscan_product_kernel<<<nblocks(C, POLY_THREADS), POLY_THREADS, 0, ctx->stream>>>(d_p, n, CHUNK, prods);
    prodscan_carries_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(prods, C, to_dev(z0));
    prodscan_write_kernel<<<nblocks(C, POLY_THREADS), POLY_THREADS, 0, ctx->stream>>>(d_p, d_z, n, CHUNK, prods);
    ctx->launches += 3;
    ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

}  // namespace b200zk
