// Launch side of the NTT (see ntt.cuh for the algorithm and the reference mapping).
#include "context.hpp"
#include "ntt.cuh"
#include "ntt_warp.cuh"

namespace b200zk {

static constexpr uint32_t NTT_THREADS = 512;
static constexpr uint32_t NTT_MAX_LOG_M = 10;      // <= 1024-point DFT per tile
static constexpr uint32_t NTT_MAX_LOG_TW = 3;      // <= 8 columns = 256 B contiguous
static constexpr uint32_t NTT_TILE_CAP_LOG = 12;   // 4096 elements = 128 KiB of shared memory
static constexpr int NTT_WARP_CFG_DEFAULT = 0;     // launch shape of the warp-level kernel, see ntt_warp_launch

__global__ void __launch_bounds__(NTT_THREADS, 1) ntt_pass_kernel(const NttPassArgs a) {
    extern __shared__ half_t ntt_sm[];
    ntt_pass_block(a, blockIdx.x, blockDim.x, ntt_sm);
}

// one warp per 128-element tile, 4 KB of warp-private shared memory each.  Launch shapes (warps per block x
// resident blocks per SM), selected per launch by ntt_warp_launch below:
//   8 x 3  80 registers, 24 warps / SM, no spills
//   4 x 7  72 registers (two spilled), 28 warps / SM: 148 x 28 = 4144 resident warps, so the 8192 tiles of one
//          2^20 pass are 1.98 waves instead of the 2.31 (= 3 rounds, a quarter of the last one idle) of 8 x 3
// SPARSE: first pass of lagrange_to_coeff (ctx->ntt_sparse_hint): all-zero tiles skip the arithmetic.  The witness
// columns of a padded circuit are zero outside the used rows and the blinding rows, and a first-pass tile gathers
// rows at stride n / 128: 262 of 8192 tiles are live for the Merkle Sum Tree circuit at k = 20.
// SHOUP: butterfly and inter-pass twiddle multiplications through Field::mul_shoup against the {w, wq} tables of the
// plan (a.roots_s / a.tw_full_s): 99 wide + 16 low multiply-adds instead of 128 + 8 per multiplication.
template <int WPB, int MINB, bool SPARSE, bool SHOUP>
__global__ void __launch_bounds__(32 * WPB, MINB) ntt_warp_pass_kernel(const NttPassArgs a, uint32_t ntiles) {
    __shared__ half_t sm[WPB][256];
    const uint32_t w = threadIdx.x >> 5, wid = blockIdx.x * WPB + w;
    if (wid < ntiles) ntt_pass_warp<SPARSE, SHOUP>(a, wid, threadIdx.x & 31, sm[w]);
}
template <int WPB, int MINB> static void ntt_warp_launch_t(const NttPassArgs& a, uint32_t ntiles, bool sparse, cudaStream_t st) {
    const uint32_t grid = (ntiles + WPB - 1) / WPB;
    const bool shoup = a.roots_s != nullptr && (a.is_last || a.tw_full_s != nullptr);
    if (shoup) {
        if (sparse) ntt_warp_pass_kernel<WPB, MINB, true, true><<<grid, 32 * WPB, 0, st>>>(a, ntiles);
        else ntt_warp_pass_kernel<WPB, MINB, false, true><<<grid, 32 * WPB, 0, st>>>(a, ntiles);
    } else {
        if (sparse) ntt_warp_pass_kernel<WPB, MINB, true, false><<<grid, 32 * WPB, 0, st>>>(a, ntiles);
        else ntt_warp_pass_kernel<WPB, MINB, false, false><<<grid, 32 * WPB, 0, st>>>(a, ntiles);
    }
}
static void ntt_warp_launch(b200zk_ctx* ctx, const NttPassArgs& a, uint32_t ntiles, bool sparse) {
    static const int cfg = [] { const char* e = getenv("B200ZK_NTT_WARP_CFG"); return e ? atoi(e) : NTT_WARP_CFG_DEFAULT; }();
    if (cfg == 1) ntt_warp_launch_t<4, 7>(a, ntiles, sparse, ctx->stream);
    else if (cfg == 2) ntt_warp_launch_t<7, 4>(a, ntiles, sparse, ctx->stream);
    else if (cfg == 3) ntt_warp_launch_t<4, 8>(a, ntiles, sparse, ctx->stream);
    else ntt_warp_launch_t<8, 3>(a, ntiles, sparse, ctx->stream);
}

__global__ void ntt_pow_table_kernel(fe_t* out, const fe_t base, uint32_t count, uint32_t shift) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) ntt_pow_table_thread(out, base, i, shift);
}

__global__ void fr_scale_periodic_kernel(fe_t* a, size_t n, const fe_t* m, uint32_t period) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) { fe_t x = a[i], y = m[i & (period - 1)]; a[i] = Fr::mul(x, y); }
}

// {w, wq} records from a Montgomery-form table: w = the plain value, wq = floor(w * 2^256 / r) by 256 steps of
// shift-and-subtract long division (plan build time only)
__global__ void ntt_shoup_table_kernel(const fe_t* mont, fe2_t* out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe_t w = Fr::from_mont(mont[i]);
    fe_t rem = w, q = Fr::zero();
    for (int b = 0; b < 256; ++b) {
        // rem <- 2 rem (rem < r < 2^254: no overflow); q <- 2 q + [rem >= r]
        uint32_t carry = 0;
#pragma unroll
        for (int l = 0; l < 8; ++l) { uint32_t v = rem.l[l]; rem.l[l] = (v << 1) | carry; carry = v >> 31; }
        carry = 0;
#pragma unroll
        for (int l = 0; l < 8; ++l) { uint32_t v = q.l[l]; q.l[l] = (v << 1) | carry; carry = v >> 31; }
        bool ge = true;
        for (int l = 7; l >= 0; --l) { uint32_t pl = FrCfg::p(l); if (rem.l[l] != pl) { ge = rem.l[l] > pl; break; } }
        if (ge) {
            uint32_t borrow = 0;
#pragma unroll
            for (int l = 0; l < 8; ++l) { uint64_t d = (uint64_t)rem.l[l] - FrCfg::p(l) - borrow; rem.l[l] = (uint32_t)d; borrow = (uint32_t)(d >> 63); }
            q.l[0] |= 1u;
        }
    }
    fe2_t r; r.w = w; r.wq = q;
    out[i] = r;
}

__global__ void ntt_tw_full_kernel(fe_t* out, size_t n, const fe_t* lo, const fe_t* hi, uint32_t bits) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe_t a = lo[i & ((1u << bits) - 1)], b = hi[i >> bits];
    out[i] = Fr::mul(a, b);
}

static fe_t to_dev(const host::HFr& x) { fe_t r; memcpy(r.l, x.v, 32); return r; }

static int32_t build_plan(b200zk_ctx* ctx, uint32_t log_n, const host::HFr& omega, NttPlan& plan) {
    // tuning overrides (profiling sweeps): B200ZK_NTT_MAX_M / _MAX_TW / _TILE_CAP, log2 values
    auto tune = [](const char* name, uint32_t dflt, uint32_t lo, uint32_t hi) {
        const char* e = getenv(name);
        if (!e) return dflt;
        long v = strtol(e, nullptr, 10);
        return (v < (long)lo || v > (long)hi) ? dflt : (uint32_t)v;
    };
    plan.warp = ntt_warp_eligible(log_n, tune("B200ZK_NTT_WARP_MAX", 26, 0, 28));
    plan.shape = plan.warp ? ntt_plan_shape_warp(log_n)
                           : ntt_plan_shape(log_n, tune("B200ZK_NTT_MAX_M", NTT_MAX_LOG_M, 2, 12),
                                            tune("B200ZK_NTT_MAX_TW", NTT_MAX_LOG_TW, 0, 5), tune("B200ZK_NTT_TILE_CAP", NTT_TILE_CAP_LOG, 6, 12));
    const NttShape& s = plan.shape;
    size_t n_roots = (size_t)1 << (s.log_roots ? s.log_roots - 1 : 0);
    size_t n_lo = (size_t)1 << s.tw_lo_bits, n_hi = ((size_t)1 << log_n) >> s.tw_lo_bits;
    if (n_hi == 0) n_hi = 1;
    ZK_CUDA(ctx, cudaMalloc(&plan.roots, n_roots * sizeof(fe_t)));
    ZK_CUDA(ctx, cudaMalloc(&plan.tw_lo, n_lo * sizeof(fe_t)));
    ZK_CUDA(ctx, cudaMalloc(&plan.tw_hi, n_hi * sizeof(fe_t)));
    host::HFr w_r = omega.pow_u64(1ull << (log_n - s.log_roots));
    auto launch = [&](fe_t* out, const host::HFr& base, size_t count, uint32_t shift) {
        ntt_pow_table_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(out, to_dev(base), (uint32_t)count, shift);
        ctx->launches++;
    };
    launch(plan.roots, w_r, n_roots, 0);
    launch(plan.tw_lo, omega, n_lo, 0);
    launch(plan.tw_hi, omega, n_hi, s.tw_lo_bits);
    // full inter-pass twiddle table: one scattered 32-byte read instead of two reads and a field
    // multiplication per element and pass (the kernel is IMAD-bound, not bandwidth-bound)
    const char* e = getenv("B200ZK_NTT_FULL_TW");
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (s.npass > 1 && log_n <= 26 && (sizeof(fe_t) << log_n) * 8 < free_b && !(e && e[0] == '0')) {
        size_t N = (size_t)1 << log_n;
        if (cudaMalloc(&plan.tw_full, N * sizeof(fe_t)) == cudaSuccess) {
            ntt_tw_full_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(plan.tw_full, N, plan.tw_lo, plan.tw_hi, s.tw_lo_bits);
            ctx->launches++;
        } else { plan.tw_full = nullptr; cudaGetLastError(); }
    }
    // constant-operand tables for the warp-level kernel (B200ZK_NTT_SHOUP=0 keeps the CIOS multiplications)
    const char* es = getenv("B200ZK_NTT_SHOUP");
    // (the block kernel was tried with the same multiplier while it still ran the sizes above 2^21, and lost: 2^24 3.85 -> 4.77 ms)
    if (plan.warp && !(es && es[0] == '0') && (s.npass == 1 || plan.tw_full)) {
        size_t N = (size_t)1 << log_n;
        if (cudaMalloc(&plan.roots_s, n_roots * sizeof(fe2_t)) == cudaSuccess) {
            ntt_shoup_table_kernel<<<(unsigned)((n_roots + 127) / 128), 128, 0, ctx->stream>>>(plan.roots, plan.roots_s, n_roots);
            ctx->launches++;
            if (s.npass > 1) {
                if (cudaMalloc(&plan.tw_full_s, N * sizeof(fe2_t)) == cudaSuccess) {
                    ntt_shoup_table_kernel<<<(unsigned)((N + 127) / 128), 128, 0, ctx->stream>>>(plan.tw_full, plan.tw_full_s, N);
                    ctx->launches++;
                } else { cudaGetLastError(); cudaFree(plan.roots_s); plan.roots_s = nullptr; plan.tw_full_s = nullptr; }
            }
        } else { plan.roots_s = nullptr; cudaGetLastError(); }
    }
    ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

static int32_t ntt_run_batch(b200zk_ctx* ctx, const fe_t* d_in, uint32_t n_in, fe_t* d_out, uint32_t log_n,
                             const host::HFr& omega, const host::HFr* pre, const host::HFr* post, uint32_t batch,
                             const fe_t* d_pre_tab = nullptr, bool shared_input = false);

int32_t ntt_run(b200zk_ctx* ctx, const fe_t* d_in, uint32_t n_in, fe_t* d_out, uint32_t log_n,
                const host::HFr& omega, const host::HFr* pre, const host::HFr* post, const fe_t* d_pre_tab) {
    return ntt_run_batch(ctx, d_in, n_in, d_out, log_n, omega, pre, post, 1, d_pre_tab);
}

// `batch` independent transforms of size 2^log_n on contiguous arrays (batch > 1: no pre/post hooks,
// n_in = 2^log_n): every pass is one launch over all of them.
static int32_t ntt_run_batch(b200zk_ctx* ctx, const fe_t* d_in, uint32_t n_in, fe_t* d_out, uint32_t log_n,
                             const host::HFr& omega, const host::HFr* pre, const host::HFr* post, uint32_t batch,
                             const fe_t* d_pre_tab, bool shared_input) {
    if (log_n > 3 * NTT_MAX_LOG_M) return fail(ctx, B200ZK_EINVAL, "ntt_run", "log_n too large");
    if (batch == 0) return B200ZK_OK;
    // (the post factors are indexed by the position inside a transform, so they work for a batch; the pre factors are not)
    if (batch > 1 && (pre || (d_pre_tab && !shared_input) || ((uint64_t)batch << log_n) > 0xFFFFFFFFull)) return fail(ctx, B200ZK_EINVAL, "ntt_run", "bad batch");
    std::array<uint64_t, 5> key = {log_n, omega.v[0], omega.v[1], omega.v[2], omega.v[3]};
    auto it = ctx->ntt_plans.find(key);
    if (it == ctx->ntt_plans.end()) {
        NttPlan plan;
        ZK_TRY(build_plan(ctx, log_n, omega, plan));
        it = ctx->ntt_plans.emplace(key, plan).first;
    }
    const NttPlan& plan = it->second;
    const NttShape& s = plan.shape;
    size_t N = (size_t)1 << log_n;
    fe_t* scratch = nullptr;
    if (s.npass > 1) {
        ZK_TRY(ws_reserve(ctx, ctx->ntt_scratch, N * batch * sizeof(fe_t)));
        scratch = (fe_t*)ctx->ntt_scratch.p;
    }
    if (!ctx->ntt_attr_set) {                                   // function attributes are per device (and per ctx, cheaply)
        ZK_CUDA(ctx, cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(sizeof(fe_t) << NTT_TILE_CAP_LOG)));
        ctx->ntt_attr_set = true;
    }
    for (uint32_t p = 0; p < s.npass; ++p) {
        const NttPassShape& q = s.pass[p];
        NttPassArgs a{};
        a.in = p == 0 ? d_in : scratch;
        a.out = q.is_last ? d_out : scratch;
        a.log_n = log_n; a.log_m = q.log_m; a.log_l = q.log_l; a.log_tw = q.log_tw; a.is_last = q.is_last;
        a.log_m1 = q.log_m1; a.log_mid = q.log_mid; a.log_m3 = q.log_m3;
        a.n_in = batch > 1 ? (uint32_t)(N * batch) : (p == 0 ? n_in : (uint32_t)N);
        a.batch_tiles = (batch > 1 && q.is_last) ? q.blocks : 0;
        a.use_pre = (p == 0 && pre) ? 1 : 0;
        a.pre_tab = p == 0 ? d_pre_tab : nullptr;
        a.in_mask = (p == 0 && shared_input) ? (uint32_t)(N - 1) : 0;
        a.use_post = (q.is_last && post) ? 1 : 0;
        for (int i = 0; i < 3; ++i) {
            if (pre) a.pre[i] = to_dev(pre[i]);
            if (post) a.post[i] = to_dev(post[i]);
        }
        a.roots = plan.roots; a.log_roots = s.log_roots;
        a.tw_lo = plan.tw_lo; a.tw_hi = plan.tw_hi; a.tw_lo_bits = s.tw_lo_bits;
        a.tw_shift = log_n - q.log_m - q.log_l; a.l_offset = 0;
        a.tw_full = plan.tw_full;
        a.roots_s = plan.roots_s; a.tw_full_s = plan.tw_full_s;
        size_t smem = sizeof(fe_t) << (q.log_m + q.log_tw);
        uint32_t tile = 1u << (q.log_m + q.log_tw);
        uint32_t threads = tile / 2 < NTT_THREADS ? (tile / 2 < 32 ? 32 : tile / 2) : NTT_THREADS;
        if (plan.warp) {
            ntt_warp_launch(ctx, a, q.blocks * batch, ctx->ntt_sparse_hint && p == 0);
        } else {
            ntt_pass_kernel<<<q.blocks * batch, threads, smem, ctx->stream>>>(a);
        }
        ctx->launches++;
    }
    ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

// `batch` transforms of ONE input of 2^log_n elements, transform b of  in[r] * d_pre_tab[b * N + r]
// (the quotient cosets of coeff_to_extended: pre_tab = powers of the coset generators); outputs
// contiguous.  One launch per pass over all of them.
int32_t ntt_run_cosets(b200zk_ctx* ctx, const fe_t* d_in, fe_t* d_out, uint32_t log_n, const host::HFr& omega,
                       const fe_t* d_pre_tab, uint32_t batch) {
    if (batch == 1) return ntt_run_batch(ctx, d_in, 1u << log_n, d_out, log_n, omega, nullptr, nullptr, 1, d_pre_tab, false);
    if (d_in == d_out) return fail(ctx, B200ZK_EINVAL, "ntt_run_cosets", "in-place not supported");
    return ntt_run_batch(ctx, d_in, 1u << log_n, d_out, log_n, omega, nullptr, nullptr, batch, d_pre_tab, true);
}

// Column step of a four-step NTT of size N = R * C sharded by column blocks:
// d_block is this rank's [R][Cg] block (row-major, Cg = 2^log_cg columns starting at global column
// col0).  In place: R-point transform down every column, then the twiddle omega_n^((col0 + c) * k_r).
int32_t ntt_colstep_run(b200zk_ctx* ctx, fe_t* d_block, uint32_t log_r, uint32_t log_cg, uint32_t col0,
                        const host::HFr& omega_n, uint32_t log_n, fe_t* const* peer_rows, uint32_t world) {
    if (log_r == 0 || log_r > NTT_MAX_LOG_M || log_r > log_n) return fail(ctx, B200ZK_EINVAL, "ntt_colstep", "row count must be 2^1..2^10");
    // dedicated plan: roots of order R, two-level table of omega_n
    std::array<uint64_t, 5> key = {((uint64_t)log_r << 32) | ((uint64_t)1 << 62) | log_n, omega_n.v[0], omega_n.v[1], omega_n.v[2], omega_n.v[3]};
    auto it = ctx->ntt_plans.find(key);
    if (it == ctx->ntt_plans.end()) {
        NttPlan plan;
        plan.shape = NttShape{};
        plan.shape.log_n = log_n; plan.shape.log_roots = log_r; plan.shape.tw_lo_bits = (log_n + 1) / 2;
        size_t n_roots = (size_t)1 << (log_r - 1), n_lo = (size_t)1 << plan.shape.tw_lo_bits, n_hi = ((size_t)1 << log_n) >> plan.shape.tw_lo_bits;
        if (n_hi == 0) n_hi = 1;
        ZK_CUDA(ctx, cudaMalloc(&plan.roots, n_roots * sizeof(fe_t)));
        ZK_CUDA(ctx, cudaMalloc(&plan.tw_lo, n_lo * sizeof(fe_t)));
        ZK_CUDA(ctx, cudaMalloc(&plan.tw_hi, n_hi * sizeof(fe_t)));
        host::HFr w_r = omega_n.pow_u64(1ull << (log_n - log_r));
        ntt_pow_table_kernel<<<(unsigned)((n_roots + 127) / 128), 128, 0, ctx->stream>>>(plan.roots, to_dev(w_r), (uint32_t)n_roots, 0);
        ntt_pow_table_kernel<<<(unsigned)((n_lo + 127) / 128), 128, 0, ctx->stream>>>(plan.tw_lo, to_dev(omega_n), (uint32_t)n_lo, 0);
        ntt_pow_table_kernel<<<(unsigned)((n_hi + 127) / 128), 128, 0, ctx->stream>>>(plan.tw_hi, to_dev(omega_n), (uint32_t)n_hi, plan.shape.tw_lo_bits);
        ctx->launches += 3;
        // R <= 128: the column step is ONE pass of the warp-level kernel (tiles of R points x 128 / R columns) and gets its
        // constant-operand tables — the twiddle records of the whole transform, omega_n^E for E < N — when they fit
        plan.warp = log_r >= 4 && log_r <= NTT_WARP_TILE_LOG;
        const char* es = getenv("B200ZK_NTT_SHOUP");
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if (plan.warp && !(es && es[0] == '0') && log_n <= 26 && (sizeof(fe_t) << log_n) * 8 < free_b) {
            const size_t N = (size_t)1 << log_n;
            if (cudaMalloc(&plan.tw_full, N * sizeof(fe_t)) == cudaSuccess && cudaMalloc(&plan.tw_full_s, N * sizeof(fe2_t)) == cudaSuccess &&
                cudaMalloc(&plan.roots_s, n_roots * sizeof(fe2_t)) == cudaSuccess) {
                ntt_tw_full_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(plan.tw_full, N, plan.tw_lo, plan.tw_hi, plan.shape.tw_lo_bits);
                ntt_shoup_table_kernel<<<(unsigned)((N + 127) / 128), 128, 0, ctx->stream>>>(plan.tw_full, plan.tw_full_s, N);
                ntt_shoup_table_kernel<<<(unsigned)((n_roots + 127) / 128), 128, 0, ctx->stream>>>(plan.roots, plan.roots_s, n_roots);
                ctx->launches += 3;
            } else {
                cudaGetLastError();
                cudaFree(plan.tw_full); cudaFree(plan.tw_full_s); cudaFree(plan.roots_s);
                plan.tw_full = nullptr; plan.tw_full_s = nullptr; plan.roots_s = nullptr;
            }
        }
        it = ctx->ntt_plans.emplace(key, plan).first;
    }
    const NttPlan& plan = it->second;
    if (!ctx->ntt_attr_set) {
        ZK_CUDA(ctx, cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(fe_t) << NTT_TILE_CAP_LOG)));
        ctx->ntt_attr_set = true;
    }
    NttPassArgs a{};
    a.in = d_block; a.out = d_block;
    a.log_n = log_r + log_cg;                       // local array size (addressing only)
    a.log_m = log_r; a.log_l = log_cg; a.is_last = 0;
    uint32_t cap = NTT_TILE_CAP_LOG - log_r;
    a.log_tw = std::min(std::min(NTT_MAX_LOG_TW, cap), log_cg);
    a.n_in = 1u << (log_r + log_cg);
    a.roots = plan.roots; a.log_roots = log_r;
    a.tw_lo = plan.tw_lo; a.tw_hi = plan.tw_hi; a.tw_lo_bits = plan.shape.tw_lo_bits;
    a.tw_full = plan.tw_full; a.roots_s = plan.roots_s; a.tw_full_s = plan.tw_full_s;
    a.tw_shift = 0; a.l_offset = col0;              // exponent (col0 + c) * k_r < C * R = N
    a.canonical_out = 1;
    if (peer_rows) {                                // fused all-to-all: row k -> rank k / (R / world), see NttPassArgs
        if (world == 0 || world > 8 || (world & (world - 1)) || ((1u << log_r) % world)) return fail(ctx, B200ZK_EINVAL, "ntt_colstep", "world must be a power of two <= 8 dividing R");
        uint32_t lw = 0; while ((1u << lw) < world) ++lw;
        a.scatter = 1; a.log_rows_per_rank = log_r - lw; a.log_c_total = log_n - log_r;
        for (uint32_t j = 0; j < world; ++j) a.peers[j] = peer_rows[j];
    }
    if (plan.warp && log_cg >= NTT_WARP_TILE_LOG - log_r) {
        a.log_tw = NTT_WARP_TILE_LOG - log_r;
        ntt_warp_launch(ctx, a, 1u << (log_r + log_cg - NTT_WARP_TILE_LOG), false);
    } else {
        uint32_t tile = 1u << (a.log_m + a.log_tw);
        uint32_t threads = tile / 2 < NTT_THREADS ? (tile / 2 < 32 ? 32 : tile / 2) : NTT_THREADS;
        ntt_pass_kernel<<<1u << (log_cg - a.log_tw), threads, sizeof(fe_t) << (a.log_m + a.log_tw), ctx->stream>>>(a);
    }
    ctx->launches++;
    ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

// Row step: `nrows` independent natural-order transforms of size 2^log_c on contiguous rows, in place; `post` (may be null):
// three factors applied to the outputs by index mod 3 (the 1/n of an inverse transform: three equal factors).
int32_t ntt_rows_run(b200zk_ctx* ctx, fe_t* d_rows, uint32_t nrows, const host::HFr& omega_c, uint32_t log_c, const host::HFr* post) {
    return ntt_run_batch(ctx, d_rows, 1u << log_c, d_rows, log_c, omega_c, nullptr, post, nrows);
}

int32_t fr_scale_periodic(b200zk_ctx* ctx, fe_t* d_a, size_t n, const fe_t* d_m, uint32_t period) {
    unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 8);
    fr_scale_periodic_kernel<<<blocks, 256, 0, ctx->stream>>>(d_a, n, d_m, period);
    ctx->launches++;
    ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

}  // namespace b200zk
