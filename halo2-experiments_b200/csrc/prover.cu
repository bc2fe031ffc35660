// keygen_pk + create_proof on the device: halo2_proofs v2023_02_02 src/plonk/keygen.rs and
// src/plonk/prover.rs (+ lookup/permutation/vanishing provers, evaluation.rs, shplonk prover),
// the calls `full_prover` makes at /root/reference/src/circuits/utils.rs:35 and :40-48.
// The host side of this file only sequences launches, runs the Blake2b transcript between phases
// and does O(#queries) scalar work; every O(n) step is a kernel.  Step numbers refer to
// SURVEY.md §3.2.
#include "context.hpp"
#include "comm.hpp"
#include "cs_desc.hpp"
#include "expr.cuh"
#include "prover_kernels.cuh"
#include "transcript.hpp"
#include <algorithm>
#include <array>
#include <map>
#include <memory>
#include <new>
#include <set>
#include <chrono>
#include <string>
#include <thread>

using namespace b200zk;
using host::HAffine;
using host::HFr;

namespace b200zk {

static constexpr uint32_t PK_THREADS = 128;
static unsigned nb(size_t n, unsigned t = PK_THREADS) { return (unsigned)((n + t - 1) / t); }
static fe_t to_dev(const HFr& x) { fe_t r; memcpy(r.l, x.v, 32); return r; }

// ------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(PK_THREADS) expr_kernel(const ExprArgs a) {
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < a.rows) expr_eval_row(a, a.row0 + idx);
}
__global__ void __launch_bounds__(PK_THREADS) from_u512_kernel(const uint32_t* wide, fe_t* out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = from_u512_row(wide + 16 * i);
}
__global__ void __launch_bounds__(PK_THREADS) perm_den_kernel(const PermLagArgs a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) perm_denominator_row(a, i);
}
__global__ void __launch_bounds__(PK_THREADS) perm_num_kernel(const PermLagArgs a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) perm_numerator_row(a, i);
}
__global__ void __launch_bounds__(PK_THREADS) lookup_den_kernel(const LookupProdArgs a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) lookup_den_row(a, i);
}
__global__ void __launch_bounds__(PK_THREADS) lookup_num_kernel(const LookupProdArgs a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) lookup_num_row(a, i);
}
__global__ void __launch_bounds__(PK_THREADS) quot_perm_a_kernel(const QuotPermAArgs a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.rows) quot_perm_a_row(a, a.row0 + i);
}
__global__ void __launch_bounds__(PK_THREADS) quot_perm_b_kernel(const QuotPermBArgs a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.rows) quot_perm_b_row(a, a.row0 + i);
}
__global__ void __launch_bounds__(PK_THREADS) quot_lookup_kernel(const QuotLookupArgs a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.rows) quot_lookup_row(a, a.row0 + i);
}
__global__ void __launch_bounds__(PK_THREADS) coset_interpolate_kernel(const fe_t* g, const fe_t* inv_pow, const fe_t* vinv, uint32_t C, size_t n, fe_t* out, const fe_t* extra) {
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) coset_interpolate_row(g, inv_pow, vinv, C, n, out, r, extra);
}
__global__ void __launch_bounds__(PK_THREADS) lookup_extrapolate_kernel(fe_t* g, const fe_t* inv_pow, const fe_t* lambda, const fe_t* tinv, uint32_t CL, uint32_t C, size_t n) {
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) lookup_extrapolate_row(g, inv_pow, lambda, tinv, CL, C, n, r);
}
__global__ void __launch_bounds__(PK_THREADS) fold_pieces_kernel(const fe_t* pieces, uint32_t npieces, size_t n, const fe_t xn, fe_t* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fold_pieces_row(pieces, npieces, n, xn, out, i);
}
__global__ void __launch_bounds__(PK_THREADS) axpy_kernel(fe_t* acc, const fe_t* p, const fe_t s, size_t n, int init) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) axpy_row(acc, p, s, i, init != 0);
}
__global__ void __launch_bounds__(PK_THREADS) scale_kernel(fe_t* a, const fe_t s, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) scale_row(a, s, i);
}
// acc[i] = (init ? 0 : acc[i]) + sum_j s[j] * p[j][i] for up to 16 polynomials per launch: the linear combinations of
// SHPLONK (A_i = sum_j y^j P_ij over the ~50 polynomials opened at x alone) read every polynomial once and touch the
// accumulator once per 16 of them, instead of one axpy launch (and one accumulator round trip) per polynomial.
struct LinCombArgs { const fe_t* p[16]; fe_t s[16]; uint32_t m; };
__global__ void __launch_bounds__(PK_THREADS) lincomb_kernel(fe_t* acc, const LinCombArgs a, size_t n, int init) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe_t t = init ? Fr::zero() : acc[i];
    for (uint32_t j = 0; j < a.m; ++j) { fe_t v = a.p[j][i]; t = Fr::add(t, Fr::mul(v, a.s[j])); }
    acc[i] = t;
}
struct LowCoeffs { fe_t c[8]; uint32_t m; };
__global__ void sub_low_kernel(fe_t* a, const LowCoeffs lc) {
    uint32_t i = threadIdx.x;
    if (i < lc.m) { fe_t v = a[i]; a[i] = Fr::sub(v, lc.c[i]); }
}
__global__ void __launch_bounds__(PK_THREADS) one_minus_sum_kernel(const fe_t* a, const fe_t* b, fe_t* out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) one_minus_sum_row(a, b, out, i);
}
__global__ void __launch_bounds__(PK_THREADS) sigma_kernel(const uint32_t* mc, const uint32_t* mr, const fe_t* dp, const fe_t* op, fe_t* out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sigma_row(mc, mr, dp, op, out, i);
}


}  // namespace b200zk

// ------------------------------------------------------------------ proving key
struct b200zk_pk {
    b200zk_ctx* ctx = nullptr;
    b200zk_params* params = nullptr;
    b200zk_domain* dom = nullptr;
    CsDesc cs;
    uint32_t n = 0, ext_n = 0, S = 0, chunk = 0, q = 0, L = 0, P = 0;
    fe_t *fixed_values = nullptr, *fixed_polys = nullptr, *fixed_cosets = nullptr;
    fe_t *perm_values = nullptr, *perm_polys = nullptr, *perm_cosets = nullptr;
    fe_t *l0 = nullptr, *l_last = nullptr, *l_active = nullptr;
    fe_t* omega_pows = nullptr;
    // quotient cosets c_j = zeta * extended_omega^j, j < q (see prover_kernels.cuh): ext_n = q * n
    fe_t *coset_pow = nullptr, *coset_pow_inv = nullptr;      // c_j^r and c_j^-r, [q][n]
    fe_t *coset_fac = nullptr, *vinv = nullptr;               // extended_omega^j [q]; inverse Vandermonde of c_j^n [q*q]
    std::vector<HFr> coset_t;                                  // 1 / (c_j^n - 1): the vanishing polynomial on coset j
    // lookup terms on lk_cosets_n < q cosets (prover_kernels.cuh, lookup_extrapolate_row); == q: not split
    uint32_t lk_cosets_n = 0;
    // lookup cosets and table values of ALL lookups kept side by side (4 arrays of lk_cosets_n * n per lookup), so that they
    // are computed on the side stream while the transcript-bound phases run; false = one lookup at a time after y (less memory)
    bool lk_early = false;
    fe_t *lk_lambda = nullptr, *coset_t_dev = nullptr;        // extrapolation weights [(q - CL) x CL]; coset_t on the device [q]
    // device program data
    uint32_t* d_prog = nullptr;                       // gates program | lookup programs
    uint32_t gates_len = 0;
    std::vector<std::pair<uint32_t, uint32_t>> lookup_prog;     // (off, len) into d_prog
    fe_t* d_consts = nullptr;
    int32_t *d_q_adv = nullptr, *d_q_fix = nullptr, *d_q_inst = nullptr;
    const fe_t** d_ptrs = nullptr;                    // pointer tables, see PtrTab
    char* arena = nullptr;
    size_t arena_bytes = 0;
    uint32_t* d_err = nullptr;
    float phase_ms[7] = {0, 0, 0, 0, 0, 0, 0};
    struct TimerEv { cudaEvent_t e0, e1; int slot; };
    std::vector<TimerEv> timer_events;                // pool, grown on demand
    size_t timer_used = 0;
    std::vector<void*> owned;
    // upload pipeline of create_proof: the advice columns arrive on copy_stream one by one (H2D from the caller's
    // buffers, or D2D), col_events[c] marks column c complete (blinding rows included)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_rnd = nullptr;
    std::vector<cudaEvent_t> col_events;
    std::map<std::string, std::pair<const void*, size_t>> dbg;     // buffers of the last proof, for b200zk_pk_debug_buffer
    std::vector<std::pair<const char*, double>> trace;             // host time (ms since the call) at the proof's synchronisation points
};

namespace b200zk {

// pointer tables in d_ptrs: [fixed_values F][fixed_cosets F][advice_values A][advice_cosets A][inst_values I][inst_cosets I]
struct PtrTab {
    uint32_t F, A, I;
    uint32_t fixed_values() const { return 0; }
    uint32_t fixed_cosets() const { return F; }
    uint32_t advice_values() const { return 2 * F; }
    uint32_t advice_cosets() const { return 2 * F + A; }
    uint32_t inst_values() const { return 2 * F + 2 * A; }
    uint32_t inst_cosets() const { return 2 * F + 2 * A + I; }
    uint32_t total() const { return 2 * (F + A + I); }
};

template <class T> static int32_t dev_alloc(b200zk_pk* pk, T** p, size_t count) {
    cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) return fail(pk->ctx, B200ZK_ENOMEM, "cudaMalloc(pk)", cudaGetErrorString(e));
    pk->owned.push_back(*p);
    return B200ZK_OK;
}

struct Arena {
    char* base; size_t cap, off = 0; bool ok = true;
    template <class T> T* take(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) / 256 * 256;
        if (off + bytes > cap) { ok = false; return nullptr; }
        T* p = (T*)(base + off); off += bytes; return p;
    }
};

static size_t arena_need(const b200zk_pk* pk) {
    size_t n = pk->n, ext = pk->ext_n, A = pk->cs.A, I = pk->cs.I, L = pk->L, S = pk->S;
    size_t draws = b200zk_pk_rng_draws(pk);
    size_t elems = 2 * A * n + 2 * I * n + draws + 7 * L * n + S * n + S * ext + 3 * n   // columns, lookups, perm, tmp
                   + ext * (1 + A + I + 4 + 1)                                            // h, cosets, lookup cosets + table_value, h_L
                   + n * (1 + 8 + 4 + 8);                                                 // h_poly, shplonk set sums, hx/lx/tmp, per-set quotients (sharded)
    if (pk->lk_early) elems += 4 * L * (size_t)pk->lk_cosets_n * n + 64;
    return elems * sizeof(fe_t) + (size_t)draws * 64 + (64 << 10);
}

// Per-phase device time without extra synchronisation: event pairs from a pool owned by the pk are
// recorded around each phase and summed once at the end of the proof (phase_timers_collect).
struct PhaseTimer {
    b200zk_pk* pk; size_t idx;
    PhaseTimer(b200zk_pk* pk_, int slot) : pk(pk_) {
        idx = pk->timer_used++;
        if (idx >= pk->timer_events.size()) {
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            pk->timer_events.push_back({a, b, slot});
        }
        pk->timer_events[idx].slot = slot;
        cudaEventRecord(pk->timer_events[idx].e0, pk->ctx->stream);
    }
    ~PhaseTimer() { cudaEventRecord(pk->timer_events[idx].e1, pk->ctx->stream); }
};
static void phase_timers_collect(b200zk_pk* pk) {
    cudaStreamSynchronize(pk->ctx->stream);
    for (size_t i = 0; i < pk->timer_used; ++i) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, pk->timer_events[i].e0, pk->timer_events[i].e1) == cudaSuccess) pk->phase_ms[pk->timer_events[i].slot] += ms;
    }
    pk->timer_used = 0;
}
enum { PH_MSM = 0, PH_NTT = 1, PH_QUOT = 2, PH_LOOKUP = 3, PH_PERM = 4, PH_OPEN = 5, PH_OTHER = 6 };

static int cmp_canonical(const HFr& a, const HFr& b) {
    uint64_t x[4], y[4]; a.to_canonical(x); b.to_canonical(y);
    for (int i = 3; i >= 0; --i) if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
    return 0;
}
struct FrLess { bool operator()(const HFr& a, const HFr& b) const { return cmp_canonical(a, b) < 0; } };

// arithmetic::lagrange_interpolate (host, <= 8 points): the unique polynomial of degree < m
static std::vector<HFr> lagrange_interpolate(const std::vector<HFr>& pts, const std::vector<HFr>& evals) {
    size_t m = pts.size();
    std::vector<HFr> poly(m, HFr::zero());
    for (size_t j = 0; j < m; ++j) {
        std::vector<HFr> num(1, HFr::one());
        HFr den = HFr::one();
        for (size_t k = 0; k < m; ++k) {
            if (k == j) continue;
            std::vector<HFr> nw(num.size() + 1, HFr::zero());
            for (size_t i = 0; i < num.size(); ++i) { nw[i] = nw[i] - num[i] * pts[k]; nw[i + 1] = nw[i + 1] + num[i]; }
            num.swap(nw);
            den = den * (pts[j] - pts[k]);
        }
        HFr s = evals[j] * den.inv();
        for (size_t i = 0; i < num.size(); ++i) poly[i] = poly[i] + num[i] * s;
    }
    return poly;
}
static HFr eval_small(const std::vector<HFr>& poly, const HFr& x) {
    HFr acc = HFr::zero();
    for (size_t i = poly.size(); i-- > 0;) acc = acc * x + poly[i];
    return acc;
}
static HFr rotate_omega(const b200zk_domain* d, const HFr& x, int rot) {
    return rot >= 0 ? x * d->omega.pow_u64((uint64_t)rot) : x * d->omega_inv.pow_u64((uint64_t)(-(int64_t)rot));
}

// Side-stream section: between construction and destruction every launch that goes through
// ctx->stream (NTTs, phase timers) lands on the low-priority side stream, ordered after everything
// the main stream has enqueued so far; the main stream consumes the results after waiting for
// ctx->ev_join (re-recorded at the end of every section; the side stream is in-order, so the last
// record covers all earlier sections).
struct SideStream {
    b200zk_ctx* ctx;
    explicit SideStream(b200zk_ctx* c) : ctx(c) {
        cudaEventRecord(ctx->ev_fork, ctx->stream);
        cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0);
        std::swap(ctx->stream, ctx->stream2);
        std::swap(ctx->ntt_scratch, ctx->ntt_scratch2);
    }
    ~SideStream() {
        cudaEventRecord(ctx->ev_join, ctx->stream);
        std::swap(ctx->stream, ctx->stream2);
        std::swap(ctx->ntt_scratch, ctx->ntt_scratch2);
    }
};

// several full-length columns in one batched launch sequence (msm_run_multi), <= 24 at a time
static int32_t commit_multi_dev(b200zk_pk* pk, const std::vector<const fe_t*>& cols, size_t len, bool lagrange, std::vector<HAffine>& outs) {
    PhaseTimer t(pk, PH_MSM);
    outs.resize(cols.size());
    for (size_t b = 0; b < cols.size(); b += 24) {
        uint32_t m = (uint32_t)std::min<size_t>(24, cols.size() - b);
        ZK_TRY(params_commit_multi(pk->params, cols.data() + b, m, len, lagrange, outs.data() + b));
    }
    return B200ZK_OK;
}
// `count` contiguous Lagrange columns to coefficient form in one launch per pass
static int32_t lagrange_to_coeff_batch(b200zk_pk* pk, fe_t* d_a, uint32_t count) {
    PhaseTimer t(pk, PH_NTT);
    HFr post[3] = {pk->dom->ifft_divisor, pk->dom->ifft_divisor, pk->dom->ifft_divisor};
    pk->ctx->ntt_sparse_hint = true;
    int32_t rc = ntt_rows_run(pk->ctx, d_a, count, pk->dom->omega_inv, pk->dom->k, post);
    pk->ctx->ntt_sparse_hint = false;
    return rc;
}
static int32_t lagrange_to_coeff(b200zk_pk* pk, fe_t* d_a) {
    PhaseTimer t(pk, PH_NTT);
    HFr post[3] = {pk->dom->ifft_divisor, pk->dom->ifft_divisor, pk->dom->ifft_divisor};
    pk->ctx->ntt_sparse_hint = true;
    int32_t rc = ntt_run(pk->ctx, d_a, pk->n, d_a, pk->dom->k, pk->dom->omega_inv, nullptr, post);
    pk->ctx->ntt_sparse_hint = false;
    return rc;
}
// coeff_to_extended restricted to the q cosets the quotient needs: coset j = size-n NTT of a_r * c_j^r
static int32_t coeff_to_extended(b200zk_pk* pk, const fe_t* d_coeffs, fe_t* d_out, uint32_t cosets = 0) {
    PhaseTimer t(pk, PH_NTT);
    return ntt_run_cosets(pk->ctx, d_coeffs, d_out, pk->dom->k, pk->dom->omega, pk->coset_pow, cosets ? cosets : pk->q);
}
static int32_t eval_dev(b200zk_pk* pk, const fe_t* d_poly, size_t len, const HFr& x, HFr* out) {
    return recurrence_run(pk->ctx, d_poly, nullptr, len, x, out);
}

static ExprArgs expr_args(const b200zk_pk* pk, uint32_t prog_off, uint32_t prog_len, bool extended, uint32_t mode,
                          const HFr ch[4], fe_t* out0, fe_t* out1) {
    PtrTab pt{pk->cs.F, pk->cs.A, pk->cs.I};
    ExprArgs a{};
    a.prog = pk->d_prog + prog_off; a.prog_len = prog_len; a.consts = pk->d_consts;
    a.fixed = pk->d_ptrs + (extended ? pt.fixed_cosets() : pt.fixed_values());
    a.advice = pk->d_ptrs + (extended ? pt.advice_cosets() : pt.advice_values());
    a.instance = pk->d_ptrs + (extended ? pt.inst_cosets() : pt.inst_values());
    a.q_fixed = pk->d_q_fix; a.q_advice = pk->d_q_adv; a.q_instance = pk->d_q_inst;
    a.log_size = pk->dom->k;
    a.rows = extended ? pk->ext_n : pk->n;
    a.rot_scale = 1u;
    for (int i = 0; i < 4; ++i) a.factors[i] = to_dev(ch[i]);
    { HFr yp = HFr::one(); for (int i = 0; i < 9; ++i) { a.ypow[i] = to_dev(yp); yp = yp * ch[EXF_Y]; } }
    a.mode = mode; a.out0 = out0; a.out1 = out1;
    return a;
}

// ------------------------------------------------------------------ common sub-expressions
// One evaluator program = a list of expressions (postfix ranges of cs.prog), each followed by a
// FOLD word.  The trees are hash-consed (ADD / MUL operands unordered); a node that the emission
// reaches more than once and that contains at least one multiplication is computed once and kept
// in a TEE/TMP slot (expr.cuh).  When more than EX_TMPS nodes qualify, those saving the most
// multiplications win.
struct ExFold { uint32_t off, len, fold_word; };

static bool ex_share_common(const std::vector<uint32_t>& src, const std::vector<ExFold>& items, std::vector<uint32_t>& out,
                            bool group_common_factor = false) {
    struct Node { uint32_t op, arg; int l, r; uint32_t muls; };
    std::vector<Node> nodes;
    std::map<std::array<int64_t, 4>, int> intern;
    auto make = [&](uint32_t op, uint32_t arg, int l, int r) {
        int a = l, b = r;
        if ((op == EX_ADD || op == EX_MUL) && a > b) std::swap(a, b);
        std::array<int64_t, 4> key{(int64_t)op, (int64_t)arg, a, b};
        auto it = intern.find(key);
        if (it != intern.end()) return it->second;
        uint32_t muls = (op == EX_MUL || op == EX_SCALE ? 1u : 0u) + (l >= 0 ? nodes[l].muls : 0u) + (r >= 0 ? nodes[r].muls : 0u);
        nodes.push_back({op, arg, l, r, muls});
        intern[key] = (int)nodes.size() - 1;
        return (int)nodes.size() - 1;
    };
    std::vector<int> roots;
    for (auto& it : items) {
        std::vector<int> st;
        for (uint32_t i = it.off; i < it.off + it.len; ++i) {
            uint32_t op = src[i] & 0xff, arg = src[i] >> 8;
            if (op <= EX_INSTANCE) st.push_back(make(op, arg, -1, -1));
            else if (op == EX_NEG || op == EX_SCALE) { if (st.empty()) return false; st.back() = make(op, arg, st.back(), -1); }
            else if (op == EX_ADD || op == EX_MUL) {
                if (st.size() < 2) return false;
                int r = st.back(); st.pop_back();
                st.back() = make(op, 0, st.back(), r);
            } else return false;
        }
        if (st.size() != 1) return false;
        roots.push_back(st[0]);
    }
    // how often the emission asks for each node when shared nodes are expanded once
    std::vector<uint32_t> req(nodes.size(), 0);
    std::vector<int> work;
    for (int root : roots) {
        work.push_back(root);
        while (!work.empty()) {
            int id = work.back(); work.pop_back();
            if (++req[id] == 1) { if (nodes[id].l >= 0) work.push_back(nodes[id].l); if (nodes[id].r >= 0) work.push_back(nodes[id].r); }
        }
    }
    std::vector<int> cand;
    for (size_t id = 0; id < nodes.size(); ++id) if (req[id] >= 2 && nodes[id].muls >= 1) cand.push_back((int)id);
    std::sort(cand.begin(), cand.end(), [&](int a, int b) {
        uint64_t sa = (uint64_t)(req[a] - 1) * nodes[a].muls, sb = (uint64_t)(req[b] - 1) * nodes[b].muls;
        return sa != sb ? sa > sb : a < b;
    });
    if (cand.size() > (size_t)EX_TMPS) cand.resize(EX_TMPS);
    std::vector<int> slot(nodes.size(), -1);
    for (size_t i = 0; i < cand.size(); ++i) slot[cand[i]] = (int)i;
    std::vector<char> done(nodes.size(), 0);
    // iterative postfix emission: (node, phase)
    auto emit = [&](int root) {
        std::vector<std::pair<int, int>> stk{{root, 0}};
        while (!stk.empty()) {
            auto [id, phase] = stk.back(); stk.pop_back();
            const Node& nd = nodes[id];
            if (phase == 0) {
                if (slot[id] >= 0 && done[id]) { out.push_back(EX_TMP | ((uint32_t)slot[id] << 8)); continue; }
                stk.push_back({id, 1});
                if (nd.r >= 0) stk.push_back({nd.r, 0});
                if (nd.l >= 0) stk.push_back({nd.l, 0});
            } else {
                out.push_back(nd.op | (nd.arg << 8));
                if (slot[id] >= 0) { out.push_back(EX_TEE | ((uint32_t)slot[id] << 8)); done[id] = 1; }
            }
        }
    };
    // common factor of two product roots (-1 if none)
    auto common = [&](int a, int b) {
        if (nodes[a].op != EX_MUL || nodes[b].op != EX_MUL) return -1;
        for (int x : {nodes[a].l, nodes[a].r}) if (x == nodes[b].l || x == nodes[b].r) return x;
        return -1;
    };
    for (size_t k = 0; k < roots.size();) {
        size_t m = 1;
        int f = -1;
        if (group_common_factor && k + 1 < roots.size()) {
            f = common(roots[k], roots[k + 1]);
            if (f >= 0) {
                m = 2;
                while (k + m < roots.size() && m < 8 && nodes[roots[k + m]].op == EX_MUL &&
                       (nodes[roots[k + m]].l == f || nodes[roots[k + m]].r == f)) ++m;
            }
        }
        if (f < 0) {
            emit(roots[k]);
            out.push_back(items[k].fold_word);
            ++k;
            continue;
        }
        for (size_t j = 0; j < m; ++j) {                        // cofactors into acc1 by Horner in y
            const Node& nd = nodes[roots[k + j]];
            emit(nd.l == f ? nd.r : nd.l);
            out.push_back(j == 0 ? (uint32_t)EX_SET1 : (EX_FOLD | ((1u << 4 | EXF_Y) << 8)));
        }
        emit(f);
        out.push_back(EX_GROUP | ((uint32_t)m << 8));
        k += m;
    }
    return true;
}

// ------------------------------------------------------------------ one proof over several ranks
// Rank layout of a sharded create_proof (comm.hpp).  G = 1 degenerates to the single-GPU prover: every
// `mine` test is true, every range is the whole, every exchange is a no-op.
struct Shard {
    Comm* cm; int G, R; uint32_t q;
    explicit Shard(b200zk_pk* pk) : cm(pk->ctx->comm), G(cm ? cm->world : 1), R(cm ? cm->rank : 0), q(pk->q) {}
    bool on() const { return G > 1; }
    // columns / lookups / permutation sets round-robin
    int owner(uint32_t i) const { return (int)(i % (uint32_t)G); }
    bool mine(uint32_t i) const { return owner(i) == R; }
    // quotient cosets in contiguous blocks: rank r evaluates cosets [coset_lo(r), coset_lo(r + 1))
    uint32_t coset_lo(int r) const { return (uint32_t)((uint64_t)q * (uint32_t)r / (uint32_t)G); }
    int coset_owner(uint32_t j) const { for (int r = 0; r < G; ++r) if (j >= coset_lo(r) && j < coset_lo(r + 1)) return r; return 0; }
    // point range of a dense commit
    size_t pt_lo(size_t len, int r) const { return len * (size_t)r / (size_t)G; }
    // Lookup arguments and permutation sets go to the least loaded rank, the load being what a rank already carries for its
    // quotient cosets (q is rarely a multiple of G: with 8 ranks and 5 cosets three ranks own none and would otherwise idle
    // through half of the proof).  Weights are in units of one size-n transform: a coset costs its column extensions plus the
    // quotient kernels, a lookup its sort, grand product, three transforms and commitments, a permutation set its product.
    // Measured on 8 GPUs (MST k = 20): round-robin 39.2 ms; this placement — all 8 lookups and the 4 sets on the three ranks
    // without a coset — 36.5 ms; weighting only the work before the challenge y (2 + 2 + 2 lookups on those ranks, 1 + 1 and the
    // sets on coset owners) 38.0 ms: the coset owners are the busy ranks from start to end, anything added to them costs.
    std::vector<int> lk_own, set_own;
    std::vector<uint32_t> set_slot;                     // index of set s among its owner's sets
    uint32_t set_slots = 0;
    void assign(uint32_t L, uint32_t S, uint32_t columns) {
        std::vector<uint64_t> load((size_t)G, 0);
        const uint64_t w_coset = columns + 30, w_lookup = 15, w_set = 8;
        for (int r = 0; r < G; ++r) load[r] = (uint64_t)(coset_lo(r + 1) - coset_lo(r)) * w_coset;
        auto lightest = [&]() { int b = 0; for (int r = 1; r < G; ++r) if (load[r] < load[b]) b = r; return b; };
        lk_own.resize(L); set_own.resize(S); set_slot.resize(S);
        for (uint32_t l = 0; l < L; ++l) { int r = lightest(); lk_own[l] = r; load[r] += w_lookup; }
        std::vector<uint32_t> used((size_t)G, 0);
        for (uint32_t t = 0; t < S; ++t) { int r = lightest(); set_own[t] = r; set_slot[t] = used[r]++; load[r] += w_set; }
        for (uint32_t u : used) set_slots = std::max(set_slots, u);
    }
    int lk_owner(uint32_t l) const { return lk_own[l]; }
    bool lk_mine(uint32_t l) const { return lk_own[l] == R; }
    int set_owner(uint32_t t) const { return set_own[t]; }
    bool set_mine(uint32_t t) const { return set_own[t] == R; }
};

// Commitments to `cols`, column i computed by rank owners[i]; every rank ends up with all of them.  `flag` (optional)
// is OR-ed over the ranks on the way (the lookup permutation's "input not in table" bit).
static int32_t commit_multi_split(b200zk_pk* pk, const Shard& sh, const std::vector<const fe_t*>& cols, const std::vector<int>& owners,
                                  size_t len, bool lagrange, std::vector<HAffine>& outs, uint32_t* flag = nullptr) {
    if (!sh.on()) return commit_multi_dev(pk, cols, len, lagrange, outs);
    std::vector<const fe_t*> my;
    std::vector<size_t> slot_of(cols.size());
    std::vector<size_t> used(sh.G, 0);
    for (size_t i = 0; i < cols.size(); ++i) { slot_of[i] = used[owners[i]]++; if (owners[i] == sh.R) my.push_back(cols[i]); }
    size_t slots = 0;
    for (size_t u : used) slots = std::max(slots, u);
    std::vector<HAffine> mp;
    ZK_TRY(commit_multi_dev(pk, my, len, lagrange, mp));
    const size_t bytes = slots * sizeof(HAffine) + 8;
    std::vector<uint8_t> send(bytes, 0), recv(bytes * sh.G);
    if (!mp.empty()) memcpy(send.data(), mp.data(), mp.size() * sizeof(HAffine));
    if (flag) memcpy(send.data() + slots * sizeof(HAffine), flag, 4);
    {
        PhaseTimer t(pk, PH_OTHER);
        ZK_TRY(sh.cm->allgather_host(pk->ctx, send.data(), bytes, recv.data(), pk->ctx->stream));
    }
    outs.resize(cols.size());
    for (size_t i = 0; i < cols.size(); ++i) memcpy(&outs[i], recv.data() + (size_t)owners[i] * bytes + slot_of[i] * sizeof(HAffine), sizeof(HAffine));
    if (flag) for (int r = 0; r < sh.G; ++r) { uint32_t f; memcpy(&f, recv.data() + (size_t)r * bytes + slots * sizeof(HAffine), 4); *flag |= f; }
    return B200ZK_OK;
}

// Dense commitments sharded by point range: every rank multiplies its slice [lo, hi) of every
// column, the G partial sums per column are exchanged (64 bytes each) and added on the host.
static int32_t commit_multi_range(b200zk_pk* pk, const Shard& sh, const std::vector<const fe_t*>& cols, size_t len, bool lagrange,
                                  std::vector<HAffine>& outs) {
    if (!sh.on()) return commit_multi_dev(pk, cols, len, lagrange, outs);
    const uint32_t m = (uint32_t)cols.size();
    std::vector<HAffine> part(m);
    {
        PhaseTimer t(pk, PH_MSM);
        ZK_TRY(params_commit_range(pk->params, cols.data(), m, sh.pt_lo(len, sh.R), sh.pt_lo(len, sh.R + 1), lagrange, part.data()));
    }
    std::vector<HAffine> all((size_t)m * sh.G);
    {
        PhaseTimer t(pk, PH_OTHER);
        ZK_TRY(sh.cm->allgather_host(pk->ctx, part.data(), m * sizeof(HAffine), all.data(), pk->ctx->stream));
    }
    outs.resize(m);
    for (uint32_t i = 0; i < m; ++i) {
        host::HXyzz acc = host::hx_identity();
        for (int r = 0; r < sh.G; ++r) {
            const HAffine& a = all[(size_t)r * m + i];
            if (!(a.x.is_zero() && a.y.is_zero())) acc = host::hx_add(acc, host::hx_from_affine(a));
        }
        outs[i] = host::hx_to_affine(acc);
    }
    return B200ZK_OK;
}

// ------------------------------------------------------------------ create_proof
static int32_t prove(b200zk_pk* pk, const fe_t* d_advice_in, bool advice_on_device, const void* const* advice_host,
                     const void* const* instance_columns, const uint32_t* instance_lens, const void* rng_wide, bool rng_on_device,
                     const HFr& transcript_repr, std::vector<uint8_t>& proof_out) {
    b200zk_ctx* ctx = pk->ctx;
    const CsDesc& cs = pk->cs;
    const b200zk_domain* dom = pk->dom;
    cudaStream_t st = ctx->stream;
    for (uint32_t c = 0; c < cs.I; ++c) if (instance_lens[c] && !instance_columns[c]) return fail(ctx, B200ZK_EINVAL, "create_proof", "null instance column");
    if (!advice_on_device) for (uint32_t c = 0; c < cs.A; ++c) if (!advice_host[c]) return fail(ctx, B200ZK_EINVAL, "create_proof", "null advice column");
    cudaStreamSynchronize(ctx->stream2);                          // nothing of an earlier (failed) proof may still be writing the arena
    if (pk->copy_stream) cudaStreamSynchronize(pk->copy_stream);
    const size_t n = pk->n, ext = pk->ext_n;
    const uint32_t A = cs.A, I = cs.I, F = cs.F, L = pk->L, S = pk->S, bf = cs.bf;
    const size_t usable = n - (bf + 1);
    const uint32_t rot_scale = 1u;                               // rotations stay inside a coset (prover_kernels.cuh)
    for (float& f : pk->phase_ms) f = 0;
    pk->timer_used = 0;
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    Shard sh(pk);
    sh.assign(pk->L, pk->S, pk->cs.A + pk->cs.I + pk->S + 3 * pk->L);
    pk->trace.clear();
    const auto t_start = std::chrono::steady_clock::now();
    auto mark = [&](const char* label) {              // called right after a host synchronisation: where the wall clock of the proof goes
        pk->trace.push_back({label, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count()});
    };
    // quotient cosets this rank extends and evaluates
    const uint32_t cj0 = sh.coset_lo(sh.R), cj1 = sh.coset_lo(sh.R + 1);
    if (sh.on()) ZK_TRY(sh.cm->set_window(ctx, pk->arena, pk->arena_bytes));
    auto share = [&](const std::vector<CommPiece>& pieces) -> int32_t {
        if (!sh.on() || pieces.empty()) return B200ZK_OK;
        PhaseTimer t(pk, PH_OTHER);
        return sh.cm->share(ctx, pieces.data(), pieces.size(), ctx->stream);
    };
    auto my_cosets = [&](const fe_t* d_coeffs, fe_t* d_out, uint32_t j0, uint32_t j1) -> int32_t {     // cosets [j0, j1) of d_out
        if (j0 >= j1) return B200ZK_OK;
        PhaseTimer t(pk, PH_NTT);
        return ntt_run_cosets(ctx, d_coeffs, d_out + (size_t)j0 * pk->n, pk->dom->k, pk->dom->omega, pk->coset_pow + (size_t)j0 * pk->n, j1 - j0);
    };
    ZK_CUDA(ctx, cudaMemsetAsync(pk->d_err, 0, 4, st));
    host::Transcript tr;
    Arena ar{pk->arena, pk->arena_bytes};
    const size_t draws = b200zk_pk_rng_draws(pk);

    // ---- arena carve-up
    fe_t* advice_values = ar.take<fe_t>(A * n);
    fe_t* advice_polys = ar.take<fe_t>(A * n);
    fe_t* inst_values = ar.take<fe_t>(I * n);
    fe_t* inst_polys = ar.take<fe_t>(I * n);
    uint32_t* wide = ar.take<uint32_t>(draws * 16);
    fe_t* rnd = ar.take<fe_t>(draws);
    fe_t* lk_bufs = ar.take<fe_t>(7 * (size_t)L * n);            // per lookup: cin ctab pin ptab pin_poly ptab_poly z_poly
    fe_t* perm_polys = ar.take<fe_t>((size_t)S * n);
    fe_t* perm_cosets = ar.take<fe_t>((size_t)S * ext);
    fe_t* tmp_n = ar.take<fe_t>(3 * n);
    fe_t* h = ar.take<fe_t>(ext);
    fe_t* advice_cosets = ar.take<fe_t>(A * ext);
    fe_t* inst_cosets = ar.take<fe_t>(I * ext);
    fe_t* lk_cosets = ar.take<fe_t>(4 * ext);                    // z, a', s', table_value
    fe_t* h_lk = ar.take<fe_t>(ext);                             // lookup part of the quotient when it runs on fewer cosets
    const uint32_t CL = pk->lk_cosets_n;
    const bool lk_split = L && CL < pk->q;
    const size_t lk_span = (size_t)(lk_split ? CL : pk->q) * n;  // rows of one lookup coset array
    fe_t* lk_all = pk->lk_early ? ar.take<fe_t>(4 * (size_t)L * lk_span) : nullptr;
    // arrays of lookup l: product z, permuted input a', permuted table s', table value (all indexed by global coset row)
    auto LKC = [&](uint32_t l, uint32_t which) { return pk->lk_early ? lk_all + ((size_t)l * 4 + which) * lk_span : lk_cosets + (size_t)which * ext; };
    // this rank's share of the cosets the lookup terms run on
    const uint32_t lj0 = std::min(cj0, lk_split ? CL : pk->q), lj1 = std::min(cj1, lk_split ? CL : pk->q);
    const uint32_t lk_row0 = lj0 * (uint32_t)n, lk_rows = (lj1 - lj0) * (uint32_t)n;
    auto lookup_table_value = [&](uint32_t l, const HFr* chv) {  // (compressed input + beta)(compressed table + gamma) on this rank's lookup cosets
        PhaseTimer t(pk, PH_QUOT);
        ExprArgs ea = expr_args(pk, pk->lookup_prog[l].first, pk->lookup_prog[l].second, true, EXM_LOOKUP_PROD, chv, LKC(l, 3), nullptr);
        ea.row0 = lk_row0; ea.rows = lk_rows;
        expr_kernel<<<nb(lk_rows), PK_THREADS, 0, ctx->stream>>>(ea);
        ctx->launches++;
    };
    fe_t* h_poly = ar.take<fe_t>(n);
    fe_t* set_sums = ar.take<fe_t>(8 * n);
    fe_t* sh_tmp = ar.take<fe_t>(4 * n);
    fe_t* set_quot = ar.take<fe_t>(8 * n);                       // sharded proof: Q_i of every rotation set, computed by its owner
    if (!ar.ok) return fail(ctx, B200ZK_ENOMEM, "create_proof", "arena too small");
    auto LK = [&](uint32_t l, uint32_t which) { return lk_bufs + ((size_t)l * 7 + which) * n; };
    pk->dbg.clear();
    pk->dbg["advice_values"] = {advice_values, A * n}; pk->dbg["advice_polys"] = {advice_polys, A * n};
    pk->dbg["lookups"] = {lk_bufs, 7 * (size_t)L * n}; pk->dbg["perm_polys"] = {perm_polys, (size_t)S * n};
    pk->dbg["h"] = {h, ext}; pk->dbg["rnd"] = {rnd, draws}; pk->dbg["tmp_n"] = {tmp_n, 3 * n};
    pk->dbg["advice_cosets"] = {advice_cosets, A * ext}; pk->dbg["h_poly"] = {h_poly, n};

    // ---- pointer tables for this proof
    {
        PtrTab pt{F, A, I};
        std::vector<const fe_t*> tab(pt.total());
        for (uint32_t c = 0; c < F; ++c) { tab[pt.fixed_values() + c] = pk->fixed_values + (size_t)c * n; tab[pt.fixed_cosets() + c] = pk->fixed_cosets + (size_t)c * ext; }
        for (uint32_t c = 0; c < A; ++c) { tab[pt.advice_values() + c] = advice_values + (size_t)c * n; tab[pt.advice_cosets() + c] = advice_cosets + (size_t)c * ext; }
        for (uint32_t c = 0; c < I; ++c) { tab[pt.inst_values() + c] = inst_values + (size_t)c * n; tab[pt.inst_cosets() + c] = inst_cosets + (size_t)c * ext; }
        ZK_CUDA(ctx, cudaMemcpyAsync(pk->d_ptrs, tab.data(), tab.size() * sizeof(void*), cudaMemcpyHostToDevice, st));
        ZK_CUDA(ctx, cudaStreamSynchronize(st));
    }

    // ---- step 0/1: vk.hash_into, instances
    tr.common_scalar(transcript_repr);
    ZK_CUDA(ctx, cudaMemsetAsync(inst_values, 0, I * n * sizeof(fe_t), st));
    for (uint32_t c = 0; c < I; ++c) {
        uint32_t len = instance_lens[c];
        if (len > usable) return fail(ctx, B200ZK_ESYNTH, "create_proof", "InstanceTooLarge");
        const uint64_t* v = (const uint64_t*)instance_columns[c];
        for (uint32_t i = 0; i < len; ++i) tr.common_scalar(HFr::from_limbs(v + 4 * i));
        if (len) ZK_CUDA(ctx, cudaMemcpyAsync(inst_values + (size_t)c * n, v, (size_t)len * sizeof(fe_t), cudaMemcpyHostToDevice, st));
    }
    ZK_CUDA(ctx, cudaMemcpyAsync(inst_polys, inst_values, I * n * sizeof(fe_t), cudaMemcpyDeviceToDevice, st));
    for (uint32_t c = 0; c < I; ++c) ZK_TRY(lagrange_to_coeff(pk, inst_polys + (size_t)c * n));

    // ---- randomness: Fr::random = from_u512 of the caller's 64-byte draws, consumed in upstream order
    // Sharded proof from host buffers: the host -> device traffic is split as well — every rank uploads 1/G of the rng
    // stream and its own advice columns over its PCIe link, and the ranks exchange them over NVLink.
    const bool split_upload = sh.on() && !advice_on_device && !rng_on_device;
    if (split_upload) {
        const size_t d0 = draws * (size_t)sh.R / sh.G, d1 = draws * (size_t)(sh.R + 1) / sh.G;
        ZK_CUDA(ctx, cudaMemcpyAsync(wide + d0 * 16, (const char*)rng_wide + d0 * 64, (d1 - d0) * 64, cudaMemcpyHostToDevice, st));
        if (d1 > d0) { from_u512_kernel<<<nb(d1 - d0), PK_THREADS, 0, st>>>(wide + d0 * 16, rnd + d0, d1 - d0); ctx->launches++; }
        std::vector<CommPiece> pieces;
        for (int r = 0; r < sh.G; ++r) {
            const size_t a0 = draws * (size_t)r / sh.G, a1 = draws * (size_t)(r + 1) / sh.G;
            pieces.push_back({rnd + a0, (a1 - a0) * sizeof(fe_t), r});
        }
        ZK_TRY(share(pieces));
    } else {
        if (rng_on_device) ZK_CUDA(ctx, cudaMemcpyAsync(wide, rng_wide, draws * 64, cudaMemcpyDeviceToDevice, st));
        else ZK_CUDA(ctx, cudaMemcpyAsync(wide, rng_wide, draws * 64, cudaMemcpyHostToDevice, st));
        from_u512_kernel<<<nb(draws), PK_THREADS, 0, st>>>(wide, rnd, draws);
        ctx->launches++;
    }
    size_t rpos = 0;
    auto rng_take = [&](size_t count) { fe_t* p = rnd + rpos; rpos += count; return p; };
    auto copy_rows = [&](fe_t* dst, const fe_t* src, size_t count) {
        return cudaMemcpyAsync(dst, src, count * sizeof(fe_t), cudaMemcpyDeviceToDevice, st);
    };

    // ---- step 2: advice columns: blinding rows, blinds, commitments.
    // The columns arrive on the pk's copy stream one at a time (H2D from the caller's buffers, 32 MiB each at
    // k = 20, or D2D), each followed by its blinding rows; col_events[c] marks column c complete.  Everything that
    // does not need the transcript starts under that transfer: on the side stream the coefficient form and the
    // quotient cosets of column c as soon as it has landed (step 10, moved up), and — when the columns come from
    // the host — on the main stream the commitment to the vanishing argument's random polynomial (step 8), which
    // depends on the rng stream only.  The main stream then waits for the last column and commits all of them.
    if (!pk->copy_stream) {
        ZK_CUDA(ctx, cudaStreamCreateWithFlags(&pk->copy_stream, cudaStreamNonBlocking));
        ZK_CUDA(ctx, cudaEventCreateWithFlags(&pk->ev_rnd, cudaEventDisableTiming));
    }
    while (pk->col_events.size() < A) {
        cudaEvent_t e;
        ZK_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        pk->col_events.push_back(e);
    }
    const size_t rpos_advice = rpos;
    rng_take((size_t)A * (bf + 1));
    rng_take(A);                                                  // Blind per column (unused by KZG)
    ZK_CUDA(ctx, cudaEventRecord(pk->ev_rnd, st));
    ZK_CUDA(ctx, cudaStreamWaitEvent(pk->copy_stream, pk->ev_rnd, 0));      // rnd (and the arena's previous users on st) first
    for (uint32_t c = 0; c < A; ++c) {
        fe_t* col = advice_values + (size_t)c * n;
        if (split_upload && !sh.mine(c)) continue;                // arrives from its owner below
        if (advice_on_device) ZK_CUDA(ctx, cudaMemcpyAsync(col, d_advice_in + (size_t)c * n, n * sizeof(fe_t), cudaMemcpyDeviceToDevice, pk->copy_stream));
        else ZK_CUDA(ctx, cudaMemcpyAsync(col, advice_host[c], n * sizeof(fe_t), cudaMemcpyHostToDevice, pk->copy_stream));
        ZK_CUDA(ctx, cudaMemcpyAsync(col + usable, rnd + rpos_advice + (size_t)c * (bf + 1), (bf + 1) * sizeof(fe_t), cudaMemcpyDeviceToDevice, pk->copy_stream));
        ZK_CUDA(ctx, cudaEventRecord(pk->col_events[c], pk->copy_stream));
    }
    // position of the random polynomial in the rng stream (draw order: SURVEY.md 8(a7))
    const size_t rpos_random = rpos + (size_t)L * (2 * (bf + 1) + 2) + (size_t)S * (bf + 1) + (size_t)L * (bf + 1);
    HAffine random_pt = {host::HFq::zero(), host::HFq::zero()};
    const bool random_early = !advice_on_device;
    auto commit_random = [&]() -> int32_t {
        std::vector<HAffine> pts;
        ZK_TRY(commit_multi_range(pk, sh, {rnd + rpos_random}, n, false, pts));
        random_pt = pts[0];
        return B200ZK_OK;
    };
    if (split_upload) {                                           // random-polynomial commit under the upload, then the column exchange
        ZK_TRY(commit_random());
        std::vector<CommPiece> pieces;
        for (uint32_t c = 0; c < A; ++c) {
            if (sh.mine(c)) ZK_CUDA(ctx, cudaStreamWaitEvent(st, pk->col_events[c], 0));
            pieces.push_back({advice_values + (size_t)c * n, n * sizeof(fe_t), sh.owner(c)});
        }
        ZK_TRY(share(pieces));
    }
    {
        SideStream side(ctx);
        // columns that are already on the device arrive at once: all their inverse transforms as one batched launch per pass
        // (columns from the host keep the per-column pipeline under the upload)
        const bool batch_intt = advice_on_device && A > 1 && ((uint64_t)A << dom->k) <= 0xFFFFFFFFull;
        if (batch_intt) {
            ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, pk->col_events[A - 1], 0));
            ZK_CUDA(ctx, cudaMemcpyAsync(advice_polys, advice_values, (size_t)A * n * sizeof(fe_t), cudaMemcpyDeviceToDevice, ctx->stream));
            ZK_TRY(lagrange_to_coeff_batch(pk, advice_polys, A));
        }
        for (uint32_t c = 0; c < A; ++c) {
            if (!batch_intt) {
                if (!split_upload) ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, pk->col_events[c], 0));
                ZK_CUDA(ctx, cudaMemcpyAsync(advice_polys + (size_t)c * n, advice_values + (size_t)c * n, n * sizeof(fe_t), cudaMemcpyDeviceToDevice, ctx->stream));
                ZK_TRY(lagrange_to_coeff(pk, advice_polys + (size_t)c * n));
            }
            ZK_TRY(my_cosets(advice_polys + (size_t)c * n, advice_cosets + (size_t)c * ext, cj0, cj1));
        }
        for (uint32_t c = 0; c < I; ++c) ZK_TRY(my_cosets(inst_polys + (size_t)c * n, inst_cosets + (size_t)c * ext, cj0, cj1));
    }
    if (random_early && !split_upload) ZK_TRY(commit_random());
    if (A && !split_upload) ZK_CUDA(ctx, cudaStreamWaitEvent(st, pk->col_events[A - 1], 0));
    {
        std::vector<const fe_t*> cols;
        std::vector<int> owners;
        std::vector<HAffine> pts;
        for (uint32_t c = 0; c < A; ++c) { cols.push_back(advice_values + (size_t)c * n); owners.push_back(sh.owner(c)); }
        ZK_TRY(commit_multi_split(pk, sh, cols, owners, n, true, pts));
        for (uint32_t c = 0; c < A; ++c) tr.write_point(pts[c]);
        mark("advice_commits");
    }
    HFr ch[4];                                                    // theta, beta, gamma, y
    ch[EXF_THETA] = tr.squeeze_challenge();
    ch[EXF_BETA] = ch[EXF_GAMMA] = ch[EXF_Y] = HFr::zero();

    // ---- step 4: lookups, commit_permuted
    // lookup l lives on rank l mod G from here to its grand product; every rank walks the rng stream
    for (uint32_t l = 0; l < L; ++l) {
        fe_t *cin = LK(l, 0), *ctab = LK(l, 1), *pin = LK(l, 2), *ptab = LK(l, 3), *pin_poly = LK(l, 4), *ptab_poly = LK(l, 5);
        if (!sh.lk_mine(l)) { rng_take(2 * (bf + 1) + 2); continue; }
        {
            PhaseTimer t(pk, PH_LOOKUP);
            ExprArgs ea = expr_args(pk, pk->lookup_prog[l].first, pk->lookup_prog[l].second, false, EXM_ACC01, ch, cin, ctab);
            expr_kernel<<<nb(n), PK_THREADS, 0, st>>>(ea);
            ctx->launches++;
            ZK_TRY(lookup_permute_run(ctx, cin, ctab, (uint32_t)usable, pin, ptab, pk->d_err));
        }
        ZK_CUDA(ctx, copy_rows(pin + usable, rng_take(bf + 1), bf + 1));
        ZK_CUDA(ctx, copy_rows(ptab + usable, rng_take(bf + 1), bf + 1));
        ZK_CUDA(ctx, copy_rows(pin_poly, pin, n));
        ZK_TRY(lagrange_to_coeff(pk, pin_poly));
        rng_take(1);
        ZK_CUDA(ctx, copy_rows(ptab_poly, ptab, n));
        ZK_TRY(lagrange_to_coeff(pk, ptab_poly));
        rng_take(1);
    }
    {
        uint32_t err = 0;
        ZK_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, pk->d_err, 4, cudaMemcpyDeviceToHost, st));
        ZK_CUDA(ctx, cudaStreamSynchronize(st));
        err = *(const uint32_t*)ctx->pinned;
        if (L) {                                                  // permuted input / table commitments of all lookups, one batch
            std::vector<const fe_t*> cols;
            std::vector<int> owners;
            std::vector<HAffine> pts;
            for (uint32_t l = 0; l < L; ++l) { cols.push_back(LK(l, 2)); cols.push_back(LK(l, 3)); owners.push_back(sh.lk_owner(l)); owners.push_back(sh.lk_owner(l)); }
            ZK_TRY(commit_multi_split(pk, sh, cols, owners, n, true, pts, &err));
            for (auto& pt : pts) tr.write_point(pt);
            mark("lookup_permuted_commits");
        }
        if (err) return fail(ctx, B200ZK_ESYNTH, "create_proof", "ConstraintSystemFailure: lookup input not in table");
        // the permuted polynomials in coefficient form: every rank opens them, the coset owners extend them
        std::vector<CommPiece> pieces;
        for (uint32_t l = 0; l < L; ++l) { pieces.push_back({LK(l, 4), n * sizeof(fe_t), sh.lk_owner(l)}); pieces.push_back({LK(l, 5), n * sizeof(fe_t), sh.lk_owner(l)}); }
        ZK_TRY(share(pieces));
    }
    ch[EXF_BETA] = tr.squeeze_challenge();
    ch[EXF_GAMMA] = tr.squeeze_challenge();
    const HFr beta = ch[EXF_BETA], gamma = ch[EXF_GAMMA];
    if (pk->lk_early && lk_rows) {                               // side stream: a', s' cosets and the table values (theta, beta, gamma known)
        SideStream side(ctx);
        for (uint32_t l = 0; l < L; ++l) {
            ZK_TRY(my_cosets(LK(l, 4), LKC(l, 1), lj0, lj1));
            ZK_TRY(my_cosets(LK(l, 5), LKC(l, 2), lj0, lj1));
            lookup_table_value(l, ch);
        }
    }

    // ---- step 6: permutation argument
    auto column_values = [&](uint32_t type, uint32_t idx) -> const fe_t* {
        return type == 0 ? advice_values + (size_t)idx * n : type == 1 ? pk->fixed_values + (size_t)idx * n : inst_values + (size_t)idx * n;
    };
    auto column_cosets = [&](uint32_t type, uint32_t idx) -> const fe_t* {
        return type == 0 ? advice_cosets + (size_t)idx * ext : type == 1 ? pk->fixed_cosets + (size_t)idx * ext : inst_cosets + (size_t)idx * ext;
    };
    {
        // Upstream chains the sets through z_s[0] = z_{s-1}[n - (bf + 1)].  On one GPU that is a 32-byte read-back per
        // set.  Sharded, set s lives on rank s mod G: every rank runs its sets from z[0] = 1, the S totals
        // t_s = z'_s[n - (bf + 1)] are exchanged, and set s is scaled by t_0 ... t_{s-1} — the same field elements.
        HFr last_z = HFr::one(), deltaomega = HFr::one();
        std::vector<const fe_t*> blind_rows(S);
        for (uint32_t s = 0; s < S; ++s) {
            fe_t* z = perm_polys + (size_t)s * n;
            uint32_t c0 = s * pk->chunk, c1 = std::min<uint32_t>(c0 + pk->chunk, pk->P);
            PermLagArgs pa{};
            pa.ncols = c1 - c0; pa.n = (uint32_t)n;
            for (uint32_t j = c0; j < c1; ++j) {
                pa.values[j - c0] = column_values(cs.perm[j].first, cs.perm[j].second);
                pa.sigma[j - c0] = pk->perm_values + (size_t)j * n;
                pa.coef[j - c0] = to_dev(deltaomega * beta);
                deltaomega = deltaomega * host::fr_delta();
            }
            blind_rows[s] = rng_take(bf);
            rng_take(1);
            if (!sh.set_mine(s)) continue;
            {
                PhaseTimer t(pk, PH_PERM);
                pa.beta = to_dev(beta); pa.gamma = to_dev(gamma); pa.omega_pows = pk->omega_pows; pa.out = tmp_n;
                perm_den_kernel<<<nb(n), PK_THREADS, 0, st>>>(pa);
                ctx->launches++;
                ZK_TRY(batch_invert_run(ctx, tmp_n, n, 0));
                perm_num_kernel<<<nb(n), PK_THREADS, 0, st>>>(pa);
                ctx->launches++;
                ZK_TRY(prefix_product_run(ctx, tmp_n, z, n, sh.on() ? HFr::one() : last_z));
            }
            ZK_CUDA(ctx, cudaMemcpyAsync((fe_t*)ctx->pinned + (sh.on() ? sh.set_slot[s] : 0), z + (n - (bf + 1)), sizeof(fe_t), cudaMemcpyDeviceToHost, st));
            if (!sh.on()) {
                ZK_CUDA(ctx, copy_rows(z + (n - bf), blind_rows[s], bf));
                ZK_CUDA(ctx, cudaStreamSynchronize(st));
                last_z = HFr::from_limbs(ctx->pinned);
            }
        }
        if (sh.on() && S) {
            const size_t slots = sh.set_slots;
            std::vector<HFr> mine_t(slots, HFr::one()), all_t(slots * sh.G);
            ZK_CUDA(ctx, cudaStreamSynchronize(st));
            for (uint32_t s = 0; s < S; ++s) if (sh.set_mine(s)) mine_t[sh.set_slot[s]] = HFr::from_limbs((const fe_t*)ctx->pinned + sh.set_slot[s]);
            {
                PhaseTimer t(pk, PH_OTHER);
                ZK_TRY(sh.cm->allgather_host(ctx, mine_t.data(), slots * sizeof(HFr), all_t.data(), st));
            }
            HFr carry = HFr::one();
            for (uint32_t s = 0; s < S; ++s) {
                fe_t* z = perm_polys + (size_t)s * n;
                if (sh.set_mine(s)) {
                    if (s) { scale_kernel<<<nb(n - bf), PK_THREADS, 0, st>>>(z, to_dev(carry), n - bf); ctx->launches++; }
                    ZK_CUDA(ctx, copy_rows(z + (n - bf), blind_rows[s], bf));
                }
                carry = carry * all_t[(size_t)sh.set_owner(s) * slots + sh.set_slot[s]];
            }
            mark("perm_products");
        }
        if (S) {                                                  // the S grand products: one batch of commitments, then coefficients and cosets
            std::vector<const fe_t*> cols;
            std::vector<int> owners;
            std::vector<HAffine> pts;
            for (uint32_t s = 0; s < S; ++s) { cols.push_back(perm_polys + (size_t)s * n); owners.push_back(sh.set_owner(s)); }
            ZK_TRY(commit_multi_split(pk, sh, cols, owners, n, true, pts));
            for (auto& pt : pts) tr.write_point(pt);
            mark("permutation_commits");
        }
        {
            std::vector<CommPiece> pieces;
            const bool batch_intt = !sh.on() && S > 1 && ((uint64_t)S << dom->k) <= 0xFFFFFFFFull;      // one rank: all sets in one launch per pass
            if (batch_intt) ZK_TRY(lagrange_to_coeff_batch(pk, perm_polys, S));
            for (uint32_t s = 0; s < S; ++s) {
                fe_t* z = perm_polys + (size_t)s * n;
                if (!batch_intt && sh.set_mine(s)) ZK_TRY(lagrange_to_coeff(pk, z));
                pieces.push_back({z, n * sizeof(fe_t), sh.set_owner(s)});
            }
            ZK_TRY(share(pieces));
            for (uint32_t s = 0; s < S; ++s) ZK_TRY(my_cosets(perm_polys + (size_t)s * n, perm_cosets + (size_t)s * ext, cj0, cj1));
        }
    }

    // ---- step 7: lookups, commit_product
    for (uint32_t l = 0; l < L; ++l) {
        fe_t* z = LK(l, 6);
        if (!sh.lk_mine(l)) { rng_take(bf + 1); continue; }
        {
            PhaseTimer t(pk, PH_LOOKUP);
            LookupProdArgs la{LK(l, 2), LK(l, 3), LK(l, 0), LK(l, 1), to_dev(beta), to_dev(gamma), tmp_n, (uint32_t)n};
            lookup_den_kernel<<<nb(n), PK_THREADS, 0, st>>>(la);
            ctx->launches++;
            ZK_TRY(batch_invert_run(ctx, tmp_n, n, 0));
            lookup_num_kernel<<<nb(n), PK_THREADS, 0, st>>>(la);
            ctx->launches++;
            ZK_TRY(prefix_product_run(ctx, tmp_n, z, n, HFr::one()));
        }
        ZK_CUDA(ctx, copy_rows(z + (n - bf), rng_take(bf), bf));
        rng_take(1);
    }
    if (L) {
        std::vector<const fe_t*> cols;
        std::vector<int> owners;
        std::vector<HAffine> pts;
        std::vector<CommPiece> pieces;
        for (uint32_t l = 0; l < L; ++l) { cols.push_back(LK(l, 6)); owners.push_back(sh.lk_owner(l)); pieces.push_back({LK(l, 6), n * sizeof(fe_t), sh.lk_owner(l)}); }
        ZK_TRY(commit_multi_split(pk, sh, cols, owners, n, true, pts));
        for (auto& pt : pts) tr.write_point(pt);
        mark("lookup_product_commits");
        for (uint32_t l = 0; l < L; ++l) if (sh.lk_mine(l)) ZK_TRY(lagrange_to_coeff(pk, LK(l, 6)));
        ZK_TRY(share(pieces));
        if (pk->lk_early && lk_rows) {                           // side stream: the product cosets
            SideStream side(ctx);
            for (uint32_t l = 0; l < L; ++l) ZK_TRY(my_cosets(LK(l, 6), LKC(l, 0), lj0, lj1));
        }
    }

    // ---- step 8: vanishing argument, random polynomial
    if (rpos != rpos_random) return fail(ctx, B200ZK_EINVAL, "create_proof", "rng draw order");
    const fe_t* random_poly = rng_take(n);
    rng_take(1);
    if (!random_early) ZK_TRY(commit_random());
    tr.write_point(random_pt);
    ch[EXF_Y] = tr.squeeze_challenge();
    const HFr y = ch[EXF_Y];

    // ---- step 10: advice / instance cosets were started on the side stream after the advice commitments
    ZK_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_join, 0));

    // ---- step 11: evaluate_h
    // this rank's rows of the quotient: its cosets [cj0, cj1) (all of them on a single GPU)
    const uint32_t my_row0 = cj0 * (uint32_t)n, my_rows = (cj1 - cj0) * (uint32_t)n;
    if (my_rows) {
        PhaseTimer t(pk, PH_QUOT);
        if (pk->gates_len) {
            ExprArgs ea = expr_args(pk, 0, pk->gates_len, true, EXM_ACC0, ch, h, nullptr);
            ea.row0 = my_row0; ea.rows = my_rows;
            expr_kernel<<<nb(my_rows), PK_THREADS, 0, st>>>(ea);
            ctx->launches++;
        } else {
            ZK_CUDA(ctx, cudaMemsetAsync(h + my_row0, 0, (size_t)my_rows * sizeof(fe_t), st));
        }
        if (S) {
            QuotPermAArgs qa{};
            qa.h = h; qa.y = to_dev(y); qa.l0 = pk->l0; qa.l_last = pk->l_last; qa.nsets = S;
            qa.rows = my_rows; qa.row0 = my_row0; qa.log_ext = dom->k; qa.rot_scale = rot_scale; qa.last_rot = -(int32_t)(bf + 1);
            for (uint32_t s = 0; s < S; ++s) qa.z[s] = perm_cosets + (size_t)s * ext;
            quot_perm_a_kernel<<<nb(my_rows), PK_THREADS, 0, st>>>(qa);
            ctx->launches++;
            HFr cd = beta * host::fr_zeta();                      // delta_start = beta * ZETA
            for (uint32_t s = 0; s < S; ++s) {
                uint32_t c0 = s * pk->chunk, c1 = std::min<uint32_t>(c0 + pk->chunk, pk->P);
                QuotPermBArgs qb{};
                qb.h = h; qb.y = to_dev(y); qb.beta = to_dev(beta); qb.gamma = to_dev(gamma); qb.l_active = pk->l_active;
                qb.z = perm_cosets + (size_t)s * ext; qb.ncols = c1 - c0; qb.rows = my_rows; qb.row0 = my_row0; qb.log_ext = dom->k; qb.rot_scale = rot_scale;
                qb.omega_pows = pk->omega_pows; qb.coset_fac = pk->coset_fac;
                for (uint32_t j = c0; j < c1; ++j) {
                    qb.values[j - c0] = column_cosets(cs.perm[j].first, cs.perm[j].second);
                    qb.sigma[j - c0] = pk->perm_cosets + (size_t)j * ext;
                    qb.cdelta[j - c0] = to_dev(cd);
                    cd = cd * host::fr_delta();
                }
                quot_perm_b_kernel<<<nb(my_rows), PK_THREADS, 0, st>>>(qb);
                ctx->launches++;
            }
        }
    }
    // lookup terms: on all q cosets straight into h, or — when their degree allows — on the first CL cosets into
    // h_lk, extended to the other cosets after the per-coset iNTT (lookup_extrapolate_row)
    if (lk_split && lk_rows) ZK_CUDA(ctx, cudaMemsetAsync(h_lk + lk_row0, 0, (size_t)lk_rows * sizeof(fe_t), st));
    for (uint32_t l = 0; l < L && lk_rows; ++l) {
        fe_t *zc = LKC(l, 0), *ac = LKC(l, 1), *sc = LKC(l, 2), *tv = LKC(l, 3);
        if (!pk->lk_early) {
            ZK_TRY(my_cosets(LK(l, 6), zc, lj0, lj1));
            ZK_TRY(my_cosets(LK(l, 4), ac, lj0, lj1));
            ZK_TRY(my_cosets(LK(l, 5), sc, lj0, lj1));
            lookup_table_value(l, ch);
        }
        PhaseTimer t(pk, PH_QUOT);
        QuotLookupArgs ql{};
        ql.h = lk_split ? h_lk : h; ql.y = to_dev(y); ql.beta = to_dev(beta); ql.gamma = to_dev(gamma);
        ql.l0 = pk->l0; ql.l_last = pk->l_last; ql.l_active = pk->l_active; ql.z = zc; ql.a = ac; ql.s = sc; ql.table_value = tv;
        ql.log_ext = dom->k; ql.rot_scale = rot_scale; ql.rows = lk_rows; ql.row0 = lk_row0;
        { HFr yp = y * y; for (int i = 0; i < 4; ++i) { ql.ypow[i] = to_dev(yp); yp = yp * y; } }
        quot_lookup_kernel<<<nb(lk_rows), PK_THREADS, 0, st>>>(ql);
        ctx->launches++;
    }
    ZK_CUDA(ctx, cudaGetLastError());

    // ---- step 12: vanishing construct
    {
        // divide by the vanishing polynomial (constant c_j^n - 1 on coset j), per-coset iNTT, then
        // the q x q interpolation across cosets gives the coefficients of h piece by piece
        PhaseTimer t(pk, PH_NTT);
        HFr y_lk = HFr::one();                                    // h = h_gates_perm * y^(5 L) + h_lk
        if (lk_split) {
            for (uint32_t i = 0; i < 5 * L; ++i) y_lk = y_lk * y;
            for (uint32_t j = lj0; j < lj1; ++j) {
                HFr post[3] = {dom->ifft_divisor, dom->ifft_divisor, dom->ifft_divisor};
                ZK_TRY(ntt_run(ctx, h_lk + j * n, (uint32_t)n, h_lk + j * n, dom->k, dom->omega_inv, nullptr, post));
            }
        }
        for (uint32_t j = cj0; j < cj1; ++j) {
            HFr f = dom->ifft_divisor * pk->coset_t[j] * y_lk;
            HFr post[3] = {f, f, f};
            ZK_TRY(ntt_run(ctx, h + j * n, (uint32_t)n, h + j * n, dom->k, dom->omega_inv, nullptr, post));
        }
        if (sh.on()) {                                            // every coset's coefficients to every rank: q (+ CL) x n elements
            std::vector<CommPiece> pieces;
            for (uint32_t j = 0; j < pk->q; ++j) pieces.push_back({h + (size_t)j * n, n * sizeof(fe_t), sh.coset_owner(j)});
            if (lk_split) for (uint32_t j = 0; j < CL; ++j) pieces.push_back({h_lk + (size_t)j * n, n * sizeof(fe_t), sh.coset_owner(j)});
            ZK_TRY(sh.cm->share(ctx, pieces.data(), pieces.size(), st));
        }
        if (lk_split) {
            lookup_extrapolate_kernel<<<nb(n), PK_THREADS, 0, st>>>(h_lk, pk->coset_pow_inv, pk->lk_lambda, pk->coset_t_dev, CL, pk->q, n);
            ctx->launches++;
        }
        coset_interpolate_kernel<<<nb(n), PK_THREADS, 0, st>>>(h, pk->coset_pow_inv, pk->vinv, pk->q, n, lk_cosets, lk_split ? h_lk : nullptr);
        ctx->launches++;
        ZK_CUDA(ctx, cudaGetLastError());
        h = lk_cosets;                                            // q pieces of n coefficients
    }
    const uint32_t q = pk->q;
    rng_take(q);                                                  // h_blinds
    {
        std::vector<const fe_t*> cols;
        std::vector<HAffine> pts;
        for (uint32_t i = 0; i < q; ++i) cols.push_back(h + (size_t)i * n);
        ZK_TRY(commit_multi_range(pk, sh, cols, n, false, pts));
        for (auto& pt : pts) tr.write_point(pt);
        mark("quotient_and_h_commits");
    }
    if (rpos != draws) return fail(ctx, B200ZK_EINVAL, "create_proof", "internal: rng draw count mismatch");
    const HFr x = tr.squeeze_challenge();
    const HFr xn = x.pow_u64(n);

    // ---- step 14: evaluations
    std::unique_ptr<PhaseTimer> open_timer(new PhaseTimer(pk, PH_OPEN));
    std::map<std::pair<const fe_t*, std::array<uint64_t, 4>>, HFr> eval_cache;
    auto eval_at = [&](const fe_t* poly, const HFr& pt, HFr* out) -> int32_t {
        std::array<uint64_t, 4> key = {pt.v[0], pt.v[1], pt.v[2], pt.v[3]};
        auto it = eval_cache.find({poly, key});
        if (it != eval_cache.end()) { *out = it->second; return B200ZK_OK; }
        ZK_TRY(eval_dev(pk, poly, n, pt, out));
        eval_cache[{poly, key}] = *out;
        return B200ZK_OK;
    };
    struct Query { const fe_t* poly; HFr point; };
    std::vector<Query> queries;
    int32_t rc = B200ZK_OK;
    auto finish = [&](int32_t r) { open_timer.reset(); return r; };
    const HFr x_next = rotate_omega(dom, x, 1), x_prev = rotate_omega(dom, x, -1), x_last = rotate_omega(dom, x, -(int)(bf + 1));
    // vanishing::evaluate: h_poly = sum_i (x^n)^i h_piece_i
    fold_pieces_kernel<<<nb(n), PK_THREADS, 0, st>>>(h, q, n, to_dev(xn), h_poly);
    ctx->launches++;
    // multiopen queries in upstream order (step 15); every evaluation written in step 14 is one of
    // them, so all of them are evaluated here in ONE batched launch pair and served from the cache.
    for (size_t i = 0; i < cs.adv_q.size(); i += 2) queries.push_back({advice_polys + (size_t)cs.adv_q[i] * n, rotate_omega(dom, x, cs.adv_q[i + 1])});
    for (uint32_t s = 0; s < S; ++s) { queries.push_back({perm_polys + (size_t)s * n, x}); queries.push_back({perm_polys + (size_t)s * n, x_next}); }
    for (uint32_t s = S; s-- > 0;) if (s + 1 < S) queries.push_back({perm_polys + (size_t)s * n, x_last});
    for (uint32_t l = 0; l < L; ++l) {
        queries.push_back({LK(l, 6), x}); queries.push_back({LK(l, 4), x}); queries.push_back({LK(l, 5), x});
        queries.push_back({LK(l, 4), x_prev}); queries.push_back({LK(l, 6), x_next});
    }
    for (size_t i = 0; i < cs.fix_q.size(); i += 2) queries.push_back({pk->fixed_polys + (size_t)cs.fix_q[i] * n, rotate_omega(dom, x, cs.fix_q[i + 1])});
    for (uint32_t j = 0; j < pk->P; ++j) queries.push_back({pk->perm_polys + (size_t)j * n, x});
    queries.push_back({h_poly, x});
    queries.push_back({random_poly, x});
    {
        std::vector<const fe_t*> bp; std::vector<HFr> bx;
        std::set<std::pair<const fe_t*, std::array<uint64_t, 4>>> seen;
        for (auto& qy : queries) {
            std::array<uint64_t, 4> key = {qy.point.v[0], qy.point.v[1], qy.point.v[2], qy.point.v[3]};
            if (seen.insert({qy.poly, key}).second) { bp.push_back(qy.poly); bx.push_back(qy.point); }
        }
        std::vector<HFr> vals(bp.size());
        if (!sh.on()) {
            rc = eval_batch_run(ctx, bp.data(), bx.data(), bp.size(), n, vals.data());
            if (rc != B200ZK_OK) return finish(rc);
        } else {                                                  // evaluation i on rank i mod G, 32 bytes each back to everybody
            std::vector<const fe_t*> mp; std::vector<HFr> mx;
            for (size_t i = 0; i < bp.size(); ++i) if (sh.mine((uint32_t)i)) { mp.push_back(bp[i]); mx.push_back(bx[i]); }
            const size_t slots = (bp.size() + sh.G - 1) / sh.G;
            std::vector<HFr> mv(slots, HFr::zero()), all(slots * sh.G);
            rc = eval_batch_run(ctx, mp.data(), mx.data(), mp.size(), n, mv.data());
            if (rc != B200ZK_OK) return finish(rc);
            for (size_t b0 = 0; b0 < slots; b0 += COMM_HOST_MAX / sizeof(HFr)) {       // in pieces the staging buffers hold
                const size_t m = std::min(slots - b0, COMM_HOST_MAX / sizeof(HFr));
                std::vector<HFr> part(m * sh.G);
                rc = sh.cm->allgather_host(ctx, mv.data() + b0, m * sizeof(HFr), part.data(), st);
                if (rc != B200ZK_OK) return finish(rc);
                for (int r = 0; r < sh.G; ++r) for (size_t i = 0; i < m; ++i) all[(size_t)r * slots + b0 + i] = part[(size_t)r * m + i];
            }
            for (size_t i = 0; i < bp.size(); ++i) vals[i] = all[(i % sh.G) * slots + i / sh.G];
        }
        for (size_t i = 0; i < bp.size(); ++i) eval_cache[{bp[i], {bx[i].v[0], bx[i].v[1], bx[i].v[2], bx[i].v[3]}}] = vals[i];
        mark("evaluations");
    }
    for (size_t i = 0; i < cs.adv_q.size() && rc == B200ZK_OK; i += 2) {
        HFr e; rc = eval_at(advice_polys + (size_t)cs.adv_q[i] * n, rotate_omega(dom, x, cs.adv_q[i + 1]), &e);
        tr.write_scalar(e);
    }
    for (size_t i = 0; i < cs.fix_q.size() && rc == B200ZK_OK; i += 2) {
        HFr e; rc = eval_at(pk->fixed_polys + (size_t)cs.fix_q[i] * n, rotate_omega(dom, x, cs.fix_q[i + 1]), &e);
        tr.write_scalar(e);
    }
    if (rc != B200ZK_OK) return finish(rc);
    {
        HFr e; rc = eval_at(random_poly, x, &e); tr.write_scalar(e);
    }
    for (uint32_t j = 0; j < pk->P && rc == B200ZK_OK; ++j) {
        HFr e; rc = eval_at(pk->perm_polys + (size_t)j * n, x, &e); tr.write_scalar(e);
    }
    for (uint32_t s = 0; s < S && rc == B200ZK_OK; ++s) {
        const fe_t* zp = perm_polys + (size_t)s * n;
        HFr e;
        rc = eval_at(zp, x, &e); tr.write_scalar(e);
        if (rc == B200ZK_OK) { rc = eval_at(zp, x_next, &e); tr.write_scalar(e); }
        if (rc == B200ZK_OK && s + 1 < S) { rc = eval_at(zp, x_last, &e); tr.write_scalar(e); }
    }
    for (uint32_t l = 0; l < L && rc == B200ZK_OK; ++l) {
        HFr e;
        const fe_t *zp = LK(l, 6), *ap = LK(l, 4), *sp = LK(l, 5);
        rc = eval_at(zp, x, &e); tr.write_scalar(e);
        if (rc == B200ZK_OK) { rc = eval_at(zp, x_next, &e); tr.write_scalar(e); }
        if (rc == B200ZK_OK) { rc = eval_at(ap, x, &e); tr.write_scalar(e); }
        if (rc == B200ZK_OK) { rc = eval_at(ap, x_prev, &e); tr.write_scalar(e); }
        if (rc == B200ZK_OK) { rc = eval_at(sp, x, &e); tr.write_scalar(e); }
    }
    if (rc != B200ZK_OK) return finish(rc);

    // ---- step 15: ProverSHPLONK::create_proof over `queries` (built above)
    const HFr sy = tr.squeeze_challenge();
    // construct_intermediate_sets
    struct CommSet { const fe_t* poly; std::set<HFr, FrLess> pts; };
    std::vector<CommSet> comm_sets;
    std::set<HFr, FrLess> super_points;
    for (auto& qy : queries) {
        super_points.insert(qy.point);
        auto it = std::find_if(comm_sets.begin(), comm_sets.end(), [&](const CommSet& c) { return c.poly == qy.poly; });
        if (it == comm_sets.end()) { comm_sets.push_back({qy.poly, {}}); it = comm_sets.end() - 1; }
        it->pts.insert(qy.point);
    }
    struct RotSet { std::vector<HFr> pts; std::vector<const fe_t*> polys; };
    std::vector<RotSet> rot_sets;
    for (auto& c : comm_sets) {
        std::vector<HFr> pts(c.pts.begin(), c.pts.end());
        auto it = std::find_if(rot_sets.begin(), rot_sets.end(), [&](const RotSet& r) {
            return r.pts.size() == pts.size() && std::equal(pts.begin(), pts.end(), r.pts.begin());
        });
        if (it == rot_sets.end()) { rot_sets.push_back({pts, {}}); it = rot_sets.end() - 1; }
        if (std::find(it->polys.begin(), it->polys.end(), c.poly) == it->polys.end()) it->polys.push_back(c.poly);
    }
    if (rot_sets.size() > 8) return finish(fail(ctx, B200ZK_EINVAL, "create_proof", "more than 8 rotation sets"));
    // low-degree equivalents
    std::vector<std::vector<std::vector<HFr>>> low(rot_sets.size());
    for (size_t i = 0; i < rot_sets.size(); ++i) {
        const size_t m = rot_sets[i].pts.size();
        if (m > 8) return finish(fail(ctx, B200ZK_EINVAL, "create_proof", "rotation set with more than 8 points"));
        // the Lagrange basis of the set's points once (m field inversions), then every polynomial of the set is a
        // combination of it: the same interpolant as arithmetic::lagrange_interpolate per polynomial, without an
        // inversion per polynomial and point (they were ~2 ms of host time per proof at every size)
        std::vector<std::vector<HFr>> basis(m);
        for (size_t j = 0; j < m; ++j) { std::vector<HFr> unit(m, HFr::zero()); unit[j] = HFr::one(); basis[j] = lagrange_interpolate(rot_sets[i].pts, unit); }
        for (const fe_t* poly : rot_sets[i].polys) {
            std::vector<HFr> lowp(m, HFr::zero());
            for (size_t j = 0; j < m; ++j) {
                HFr e; rc = eval_at(poly, rot_sets[i].pts[j], &e); if (rc != B200ZK_OK) return finish(rc);
                for (size_t c = 0; c < m; ++c) lowp[c] = lowp[c] + basis[j][c] * e;
            }
            low[i].push_back(lowp);
        }
    }
    const HFr sv = tr.squeeze_challenge();
    fe_t *hx = sh_tmp, *lx = sh_tmp + n, *t0 = sh_tmp + 2 * n, *t1 = sh_tmp + 3 * n;
    ZK_CUDA(ctx, cudaMemsetAsync(hx, 0, n * sizeof(fe_t), st));
    {
        // Sharded: rotation set i (its linear combination A_i and the Kate divisions that give Q_i) lives on rank i mod G; both
        // polynomials are then broadcast — every rank needs A_i again for the linearisation polynomial — and h1 = sum v^i Q_i
        // is formed everywhere.
        HFr v_pow = HFr::one();
        for (size_t i = 0; i < rot_sets.size(); ++i) {
            fe_t* sum = set_sums + i * n;                        // A_i = sum_j y^j P_ij
            if (sh.on() && !sh.mine((uint32_t)i)) continue;
            HFr y_pow = HFr::one();
            std::vector<HFr> rlow(rot_sets[i].pts.size(), HFr::zero());   // R_i = sum_j y^j r_ij
            for (size_t j0 = 0; j0 < rot_sets[i].polys.size(); j0 += 16) {
                LinCombArgs la{};
                la.m = (uint32_t)std::min<size_t>(16, rot_sets[i].polys.size() - j0);
                for (uint32_t j = 0; j < la.m; ++j) {
                    la.p[j] = rot_sets[i].polys[j0 + j];
                    la.s[j] = to_dev(y_pow);
                    for (size_t c = 0; c < low[i][j0 + j].size(); ++c) rlow[c] = rlow[c] + low[i][j0 + j][c] * y_pow;
                    y_pow = y_pow * sy;
                }
                lincomb_kernel<<<nb(n), PK_THREADS, 0, st>>>(sum, la, n, j0 == 0 ? 1 : 0);
                ctx->launches++;
            }
            // N_i = A_i - R_i ; Q_i = N_i / prod (X - p)
            ZK_CUDA(ctx, copy_rows(t0, sum, n));
            LowCoeffs lc{}; lc.m = (uint32_t)rlow.size();
            for (size_t c = 0; c < rlow.size(); ++c) lc.c[c] = to_dev(rlow[c]);
            sub_low_kernel<<<1, 8, 0, st>>>(t0, lc);
            ctx->launches++;
            fe_t *src = t0, *dst = t1;
            size_t len = n;
            for (auto& pt : rot_sets[i].pts) {
                ZK_TRY(recurrence_run(ctx, src + 1, dst, len - 1, pt, nullptr));   // kate_division
                --len; std::swap(src, dst);
            }
            if (sh.on()) {                                        // keep Q_i (zero-padded to n) for the exchange
                ZK_CUDA(ctx, cudaMemsetAsync(set_quot + i * n, 0, n * sizeof(fe_t), st));
                ZK_CUDA(ctx, copy_rows(set_quot + i * n, src, len));
            } else {
                axpy_kernel<<<nb(len), PK_THREADS, 0, st>>>(hx, src, to_dev(v_pow), len, 0);
                ctx->launches++;
                v_pow = v_pow * sv;
            }
        }
        if (sh.on()) {
            std::vector<CommPiece> pieces;
            for (size_t i = 0; i < rot_sets.size(); ++i) {
                pieces.push_back({set_sums + i * n, n * sizeof(fe_t), sh.owner((uint32_t)i)});
                pieces.push_back({set_quot + i * n, n * sizeof(fe_t), sh.owner((uint32_t)i)});
            }
            rc = share(pieces);
            if (rc != B200ZK_OK) return finish(rc);
            for (size_t i = 0; i < rot_sets.size(); ++i) {
                axpy_kernel<<<nb(n), PK_THREADS, 0, st>>>(hx, set_quot + i * n, to_dev(v_pow), n, 0);
                ctx->launches++;
                v_pow = v_pow * sv;
            }
        }
    }
    {
        open_timer.reset();
        std::vector<HAffine> pt;
        ZK_TRY(commit_multi_range(pk, sh, {hx}, n, false, pt));
        tr.write_point(pt[0]);
        mark("shplonk_h1_commit");
        open_timer.reset(new PhaseTimer(pk, PH_OPEN));
    }
    const HFr su = tr.squeeze_challenge();
    {
        auto vanish = [&](const std::vector<HFr>& roots) { HFr acc = HFr::one(); for (auto& r : roots) acc = acc * (su - r); return acc; };
        HFr v_pow = HFr::one(), const_term = HFr::zero(), z0_diff = HFr::zero();
        for (size_t i = 0; i < rot_sets.size(); ++i) {
            std::vector<HFr> diffs;
            for (auto& p : super_points)
                if (std::find_if(rot_sets[i].pts.begin(), rot_sets[i].pts.end(), [&](const HFr& o) { return o == p; }) == rot_sets[i].pts.end()) diffs.push_back(p);
            HFr z_i = vanish(diffs);
            if (i == 0) z0_diff = z_i;
            HFr y_pow = HFr::one(), c_i = HFr::zero();
            for (size_t j = 0; j < rot_sets[i].polys.size(); ++j) { c_i = c_i + eval_small(low[i][j], su) * y_pow; y_pow = y_pow * sy; }
            HFr coef = z_i * v_pow;
            axpy_kernel<<<nb(n), PK_THREADS, 0, st>>>(lx, set_sums + i * n, to_dev(coef), n, i == 0 ? 1 : 0);
            ctx->launches++;
            const_term = const_term + coef * c_i;
            v_pow = v_pow * sv;
        }
        std::vector<HFr> sp(super_points.begin(), super_points.end());
        HFr zt = vanish(sp);
        axpy_kernel<<<nb(n), PK_THREADS, 0, st>>>(lx, hx, to_dev(zt.neg()), n, 0);
        LowCoeffs lc{}; lc.m = 1; lc.c[0] = to_dev(const_term);
        sub_low_kernel<<<1, 8, 0, st>>>(lx, lc);
        ctx->launches += 2;
        ZK_TRY(recurrence_run(ctx, lx + 1, t0, n - 1, su, nullptr));              // (L(X)) / (X - u)
        scale_kernel<<<nb(n - 1), PK_THREADS, 0, st>>>(t0, to_dev(z0_diff.inv()), n - 1);
        ctx->launches++;
        ZK_CUDA(ctx, cudaGetLastError());
        open_timer.reset();
        std::vector<HAffine> pt;
        ZK_TRY(commit_multi_range(pk, sh, {t0}, n - 1, false, pt));
        tr.write_point(pt[0]);
    }
    phase_timers_collect(pk);
    mark("end");
    proof_out = tr.proof();
    return B200ZK_OK;
}

}  // namespace b200zk

// ------------------------------------------------------------------ C ABI
extern "C" {

size_t b200zk_pk_rng_draws(const b200zk_pk* pk) {
    if (!pk) return 0;
    const CsDesc& cs = pk->cs;
    size_t bf = cs.bf;
    return cs.A * (bf + 1) + cs.A + pk->L * (2 * (bf + 1) + 2) + pk->S * (bf + 1) + pk->L * (bf + 1) + pk->n + 1 + pk->q;
}

size_t b200zk_pk_proof_size(const b200zk_pk* pk) {
    if (!pk) return 0;
    const CsDesc& cs = pk->cs;
    size_t S = pk->S, L = pk->L;
    size_t points = cs.A + 2 * L + S + L + 1 + pk->q + 2;
    size_t evals = cs.adv_q.size() / 2 + cs.fix_q.size() / 2 + 1 + pk->P + (S ? 3 * S - 1 : 0) + 5 * L;
    return 32 * (points + evals);
}

uint32_t b200zk_pk_blinding_factors(const b200zk_pk* pk) { return pk ? pk->cs.bf : 0; }
uint32_t b200zk_pk_degree(const b200zk_pk* pk) { return pk ? pk->cs.degree : 0; }

// debugging aid: copy a named device buffer of the last proof to the host (elements of 32 bytes)
int32_t b200zk_pk_debug_buffer(b200zk_pk* pk, const char* name, void* host_out, size_t max_elems, size_t* count) {
    if (!pk || !name || !host_out) return B200ZK_EINVAL;
    auto it = pk->dbg.find(name);
    if (it == pk->dbg.end()) return fail(pk->ctx, B200ZK_EINVAL, "debug_buffer", "unknown buffer");
    size_t c = std::min(max_elems, it->second.second);
    if (count) *count = it->second.second;
    ZK_CUDA(pk->ctx, cudaMemcpyAsync(host_out, it->second.first, c * sizeof(fe_t), cudaMemcpyDeviceToHost, pk->ctx->stream));
    ZK_CUDA(pk->ctx, cudaStreamSynchronize(pk->ctx->stream));
    return B200ZK_OK;
}

// Field multiplications evaluate_h executes per row, as launched: out[0] custom gates (one row of every quotient coset),
// out[1] permutation terms (same rows), out[2] all lookups together (rows of the lk_cosets_n cosets the lookup terms run
// on), out[3] = q, out[4] = lk_cosets_n.  bench.py's roofline_quotient: work = n * (q * (gates + perm) + CL * lookups).
int32_t b200zk_pk_quotient_muls(const b200zk_pk* pk, uint32_t out5[5]) {
    if (!pk || !out5) return B200ZK_EINVAL;
    std::vector<uint32_t> prog(pk->gates_len + 1);
    // the programs live on the device; count from a host copy
    uint32_t total = pk->gates_len;
    for (auto& lp : pk->lookup_prog) total = std::max(total, lp.first + lp.second);
    prog.resize(total);
    if (total && cudaMemcpy(prog.data(), pk->d_prog, total * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return B200ZK_ECUDA;
    auto count = [&](uint32_t off, uint32_t len) {
        uint32_t m = 0;
        for (uint32_t i = off; i < off + len; ++i) {
            uint32_t op = prog[i] & 0xff;
            if (op == EX_MUL || op == EX_SCALE || op == EX_FOLD) m += 1;
            else if (op == EX_GROUP) m += 2;
        }
        return m;
    };
    out5[0] = count(0, pk->gates_len);
    uint32_t perm = 0;
    if (pk->S) {
        perm = 5 + 2 * (pk->S - 1);
        for (uint32_t s = 0; s < pk->S; ++s) { uint32_t c0 = s * pk->chunk, c1 = std::min<uint32_t>(c0 + pk->chunk, pk->P); perm += 3 + 4 * (c1 - c0); }
    }
    out5[1] = perm;
    uint32_t lk = 0;
    for (auto& lp : pk->lookup_prog) lk += count(lp.first, lp.second) + 1 + 13;       // table value product + quot_lookup_row
    out5[2] = lk; out5[3] = pk->q; out5[4] = pk->lk_cosets_n;
    return B200ZK_OK;
}

// host wall clock of the last create_proof at its synchronisation points: "label:ms;label:ms;..." (ms since the call)
int32_t b200zk_pk_last_trace(const b200zk_pk* pk, char* out, size_t cap) {
    if (!pk || !out || cap == 0) return B200ZK_EINVAL;
    std::string t;
    for (auto& kv : pk->trace) { char buf[96]; snprintf(buf, sizeof buf, "%s:%.3f;", kv.first, kv.second); t += buf; }
    if (t.size() + 1 > cap) return B200ZK_EINVAL;
    memcpy(out, t.c_str(), t.size() + 1);
    return B200ZK_OK;
}

int32_t b200zk_pk_last_phase_ms(const b200zk_pk* pk, float* out7) {
    if (!pk || !out7) return B200ZK_EINVAL;
    memcpy(out7, pk->phase_ms, sizeof(pk->phase_ms));
    return B200ZK_OK;
}

// keygen_vk's O(n) part (src/plonk/keygen.rs): the commitments to the fixed and the permutation
// polynomials a VerifyingKey holds, computed on the device from the pk's coefficient forms.
int32_t b200zk_pk_vk_commitments(b200zk_pk* pk, void* fixed_out, void* sigma_out) {
    if (!pk || (pk->cs.F && !fixed_out) || (pk->P && !sigma_out)) return B200ZK_EINVAL;
    cudaSetDevice(pk->ctx->device);
    size_t n = pk->n;
    auto run = [&](const fe_t* base, uint32_t count, void* out) -> int32_t {
        if (!count) return B200ZK_OK;
        std::vector<const fe_t*> cols;
        std::vector<HAffine> pts;
        for (uint32_t c = 0; c < count; ++c) cols.push_back(base + (size_t)c * n);
        ZK_TRY(commit_multi_dev(pk, cols, n, false, pts));
        for (uint32_t c = 0; c < count; ++c) { pts[c].x.store((uint8_t*)out + 64 * (size_t)c); pts[c].y.store((uint8_t*)out + 64 * (size_t)c + 32); }
        return B200ZK_OK;
    };
    ZK_TRY(run(pk->fixed_polys, pk->cs.F, fixed_out));
    ZK_TRY(run(pk->perm_polys, pk->P, sigma_out));
    pk->timer_used = 0;
    return B200ZK_OK;
}

void b200zk_pk_destroy(b200zk_pk* pk) {
    if (!pk) return;
    cudaSetDevice(pk->ctx->device);
    if (pk->copy_stream) { cudaStreamSynchronize(pk->copy_stream); cudaStreamDestroy(pk->copy_stream); }
    if (pk->ev_rnd) cudaEventDestroy(pk->ev_rnd);
    for (cudaEvent_t e : pk->col_events) cudaEventDestroy(e);
    cudaStreamSynchronize(pk->ctx->stream);
    for (void* p : pk->owned) cudaFree(p);
    for (auto& t : pk->timer_events) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
    if (pk->dom) b200zk_domain_destroy(pk->dom);
    delete pk;
}

int32_t b200zk_pk_create(b200zk_params* params, const uint32_t* cs_blob, size_t blob_words, const void* const* fixed_columns,
                         const uint32_t* map_col, const uint32_t* map_row, b200zk_pk** out) {
    if (!params || !cs_blob || !out) return B200ZK_EINVAL;
    b200zk_ctx* ctx = params->ctx;
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    b200zk_pk* pk = new (std::nothrow) b200zk_pk();
    if (!pk) return B200ZK_ENOMEM;
    pk->ctx = ctx; pk->params = params;
    auto bail = [&](int32_t rc) { b200zk_pk_destroy(pk); return rc; };
    if (!parse_cs(cs_blob, blob_words, pk->cs)) return bail(fail(ctx, B200ZK_EINVAL, "pk_create", "malformed constraint-system blob"));
    CsDesc& cs = pk->cs;
    if (cs.k != params->k || !params->d_g_lagrange) return bail(fail(ctx, B200ZK_EINVAL, "pk_create", "params do not match k / lagrange basis missing"));
    if (cs.F && !fixed_columns) return bail(B200ZK_EINVAL);
    int32_t rc = b200zk_domain_create(ctx, cs.degree, cs.k, &pk->dom);
    if (rc != B200ZK_OK) return bail(rc);
    const b200zk_domain* dom = pk->dom;
    pk->n = 1u << cs.k;
    pk->P = (uint32_t)cs.perm.size(); pk->L = (uint32_t)cs.lookups.size();
    pk->chunk = cs.degree - 2;
    pk->S = (pk->P + pk->chunk - 1) / pk->chunk;
    pk->q = dom->quotient_poly_degree;
    pk->ext_n = pk->q * pk->n;                                    // q cosets of size n (prover_kernels.cuh)
    if (pk->chunk > ZK_MAXC || pk->S > ZK_MAXC || pk->q > (uint32_t)ZK_MAXCOSETS || pk->q > (1u << (dom->extended_k - dom->k)) ||
        (uint64_t)pk->q * pk->n > 0xFFFFFFFFull)
        return bail(fail(ctx, B200ZK_EINVAL, "pk_create", "permutation too wide for this build"));
    if (pk->P && (!map_col || !map_row)) return bail(B200ZK_EINVAL);
    const size_t n = pk->n, ext = pk->ext_n;
    const uint32_t F = cs.F, P = pk->P;
    cudaStream_t st = ctx->stream;
#define PK_TRY(expr) do { int32_t _r = (expr); if (_r != B200ZK_OK) return bail(_r); } while (0)
#define PK_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return bail(fail(ctx, B200ZK_ECUDA, #expr, cudaGetErrorString(_e))); } while (0)
    PK_TRY(dev_alloc(pk, &pk->fixed_values, F * n)); PK_TRY(dev_alloc(pk, &pk->fixed_polys, F * n)); PK_TRY(dev_alloc(pk, &pk->fixed_cosets, F * ext));
    PK_TRY(dev_alloc(pk, &pk->perm_values, P * n)); PK_TRY(dev_alloc(pk, &pk->perm_polys, P * n)); PK_TRY(dev_alloc(pk, &pk->perm_cosets, P * ext));
    PK_TRY(dev_alloc(pk, &pk->l0, ext)); PK_TRY(dev_alloc(pk, &pk->l_last, ext)); PK_TRY(dev_alloc(pk, &pk->l_active, ext));
    PK_TRY(dev_alloc(pk, &pk->omega_pows, n));
    PK_TRY(dev_alloc(pk, &pk->d_err, 1));
    // quotient cosets: powers of c_j and 1/c_j, extended_omega^j, 1/(c_j^n - 1), inverse Vandermonde of c_j^n
    {
        const uint32_t C = pk->q;
        PK_TRY(dev_alloc(pk, &pk->coset_pow, (size_t)C * n)); PK_TRY(dev_alloc(pk, &pk->coset_pow_inv, (size_t)C * n));
        PK_TRY(dev_alloc(pk, &pk->coset_fac, C)); PK_TRY(dev_alloc(pk, &pk->vinv, (size_t)C * C));
        std::vector<HFr> y(C);
        std::vector<fe_t> fac(C);
        HFr w = HFr::one();
        for (uint32_t j = 0; j < C; ++j) {
            HFr cj = dom->g_coset * w;
            PK_TRY(powers_run(ctx, cj, n, pk->coset_pow + (size_t)j * n));
            PK_TRY(powers_run(ctx, cj.inv(), n, pk->coset_pow_inv + (size_t)j * n));
            y[j] = cj.pow_u64(n);
            pk->coset_t.push_back((y[j] - HFr::one()).inv());
            fac[j] = to_dev(w);
            w = w * dom->extended_omega;
        }
        // V[j][t] = y_j^t; vinv = V^-1 by Gauss-Jordan (C <= 32; the y_j are distinct, so V is regular)
        std::vector<std::vector<HFr>> m(C, std::vector<HFr>(2 * C, HFr::zero()));
        for (uint32_t j = 0; j < C; ++j) { HFr p = HFr::one(); for (uint32_t t = 0; t < C; ++t) { m[j][t] = p; p = p * y[j]; } m[j][C + j] = HFr::one(); }
        for (uint32_t c = 0; c < C; ++c) {
            uint32_t piv = c;
            while (piv < C && m[piv][c] == HFr::zero()) ++piv;
            if (piv == C) return bail(fail(ctx, B200ZK_EINVAL, "pk_create", "internal: singular coset Vandermonde"));
            std::swap(m[c], m[piv]);
            HFr inv = m[c][c].inv();
            for (auto& v : m[c]) v = v * inv;
            for (uint32_t r = 0; r < C; ++r) {
                if (r == c || m[r][c] == HFr::zero()) continue;
                HFr f = m[r][c];
                for (uint32_t k2 = 0; k2 < 2 * C; ++k2) m[r][k2] = m[r][k2] - f * m[c][k2];
            }
        }
        std::vector<fe_t> vin((size_t)C * C);
        for (uint32_t t = 0; t < C; ++t) for (uint32_t j = 0; j < C; ++j) vin[(size_t)t * C + j] = to_dev(m[t][C + j]);
        PK_CUDA(cudaMemcpyAsync(pk->coset_fac, fac.data(), C * sizeof(fe_t), cudaMemcpyHostToDevice, st));
        PK_CUDA(cudaMemcpyAsync(pk->vinv, vin.data(), vin.size() * sizeof(fe_t), cudaMemcpyHostToDevice, st));
        // cosets the lookup terms need: their degree is < (2 + deg input + deg table) * n  (circuit.rs, lookup
        // Argument::required_degree without the max(4, .) floor being relevant: it is >= 4 already)
        uint32_t CL = C;
        if (!cs.lookups.empty()) {
            uint32_t need = 0;
            for (auto& lk : cs.lookups) {
                uint32_t di = 1, dt = 1;
                for (auto& e : lk.ins) di = std::max(di, expr_degree(cs, e.first, e.second));
                for (auto& e : lk.tabs) dt = std::max(dt, expr_degree(cs, e.first, e.second));
                need = std::max(need, 2 + di + dt);
            }
            CL = std::min(C, need);
        }
        pk->lk_cosets_n = CL;
        std::vector<fe_t> lam((size_t)std::max<uint32_t>(C - CL, 1) * CL), tdev(C);
        for (uint32_t j = 0; j < C; ++j) tdev[j] = to_dev(pk->coset_t[j]);
        for (uint32_t jp = CL; jp < C; ++jp)
            for (uint32_t j = 0; j < CL; ++j) {
                HFr num = HFr::one(), den = HFr::one();
                for (uint32_t m2 = 0; m2 < CL; ++m2) { if (m2 == j) continue; num = num * (y[jp] - y[m2]); den = den * (y[j] - y[m2]); }
                lam[(size_t)(jp - CL) * CL + j] = to_dev(num * den.inv());
            }
        PK_TRY(dev_alloc(pk, &pk->lk_lambda, lam.size())); PK_TRY(dev_alloc(pk, &pk->coset_t_dev, C));
        PK_CUDA(cudaMemcpyAsync(pk->lk_lambda, lam.data(), lam.size() * sizeof(fe_t), cudaMemcpyHostToDevice, st));
        PK_CUDA(cudaMemcpyAsync(pk->coset_t_dev, tdev.data(), C * sizeof(fe_t), cudaMemcpyHostToDevice, st));
        PK_CUDA(cudaStreamSynchronize(st));
    }
    // fixed columns
    for (uint32_t c = 0; c < F; ++c) PK_CUDA(cudaMemcpyAsync(pk->fixed_values + (size_t)c * n, fixed_columns[c], n * sizeof(fe_t), cudaMemcpyHostToDevice, st));
    PK_CUDA(cudaMemcpyAsync(pk->fixed_polys, pk->fixed_values, F * n * sizeof(fe_t), cudaMemcpyDeviceToDevice, st));
    for (uint32_t c = 0; c < F; ++c) {
        PK_TRY(lagrange_to_coeff(pk, pk->fixed_polys + (size_t)c * n));
        PK_TRY(coeff_to_extended(pk, pk->fixed_polys + (size_t)c * n, pk->fixed_cosets + (size_t)c * ext));
    }
    PK_TRY(powers_run(ctx, dom->omega, n, pk->omega_pows));
    // permutation polynomials: sigma_c[r] = delta^map_col * omega^map_row
    if (P) {
        uint32_t *d_mc = nullptr, *d_mr = nullptr; fe_t* d_dp = nullptr;
        PK_TRY(dev_alloc(pk, &d_mc, P * n)); PK_TRY(dev_alloc(pk, &d_mr, P * n)); PK_TRY(dev_alloc(pk, &d_dp, P));
        for (size_t i = 0; i < (size_t)P * n; ++i) if (map_col[i] >= P || map_row[i] >= n) return bail(fail(ctx, B200ZK_EINVAL, "pk_create", "permutation mapping out of range"));
        PK_CUDA(cudaMemcpyAsync(d_mc, map_col, P * n * 4, cudaMemcpyHostToDevice, st));
        PK_CUDA(cudaMemcpyAsync(d_mr, map_row, P * n * 4, cudaMemcpyHostToDevice, st));
        PK_TRY(powers_run(ctx, host::fr_delta(), P, d_dp));
        sigma_kernel<<<nb((size_t)P * n), PK_THREADS, 0, st>>>(d_mc, d_mr, d_dp, pk->omega_pows, pk->perm_values, (size_t)P * n);
        ctx->launches++;
        PK_CUDA(cudaMemcpyAsync(pk->perm_polys, pk->perm_values, P * n * sizeof(fe_t), cudaMemcpyDeviceToDevice, st));
        for (uint32_t c = 0; c < P; ++c) {
            PK_TRY(lagrange_to_coeff(pk, pk->perm_polys + (size_t)c * n));
            PK_TRY(coeff_to_extended(pk, pk->perm_polys + (size_t)c * n, pk->perm_cosets + (size_t)c * ext));
        }
    }
    // l0, l_last, l_blind -> l_active_row
    {
        fe_t* tmp = nullptr; fe_t* lb = nullptr;
        PK_TRY(dev_alloc(pk, &tmp, n)); PK_TRY(dev_alloc(pk, &lb, ext));
        fe_t one = to_dev(HFr::one());
        auto indicator = [&](size_t lo, size_t hi, fe_t* out_ext) -> int32_t {
            ZK_CUDA(ctx, cudaMemsetAsync(tmp, 0, n * sizeof(fe_t), st));
            std::vector<fe_t> ones(hi - lo, one);
            ZK_CUDA(ctx, cudaMemcpyAsync(tmp + lo, ones.data(), ones.size() * sizeof(fe_t), cudaMemcpyHostToDevice, st));
            ZK_CUDA(ctx, cudaStreamSynchronize(st));
            ZK_TRY(lagrange_to_coeff(pk, tmp));
            return coeff_to_extended(pk, tmp, out_ext);
        };
        PK_TRY(indicator(0, 1, pk->l0));
        PK_TRY(indicator(n - cs.bf, n, lb));
        PK_TRY(indicator(n - cs.bf - 1, n - cs.bf, pk->l_last));
        one_minus_sum_kernel<<<nb(ext), PK_THREADS, 0, st>>>(pk->l_last, lb, pk->l_active, ext);
        ctx->launches++;
    }
    // programs: gates with FOLD(acc0, y); lookups with FOLD(acc0/acc1, theta)
    {
        std::vector<uint32_t> prog;
        {
            std::vector<ExFold> items;
            for (auto& g : cs.gates) items.push_back({g.first, g.second, EX_FOLD | ((0u << 4 | EXF_Y) << 8)});
            if (!ex_share_common(cs.prog, items, prog, true)) return bail(fail(ctx, B200ZK_EINVAL, "pk_create", "malformed expression program"));
        }
        pk->gates_len = (uint32_t)prog.size();
        for (auto& lk : cs.lookups) {
            uint32_t off = (uint32_t)prog.size();
            std::vector<ExFold> items;
            for (auto& e : lk.ins) items.push_back({e.first, e.second, EX_FOLD | ((0u << 4 | EXF_THETA) << 8)});
            for (auto& e : lk.tabs) items.push_back({e.first, e.second, EX_FOLD | ((1u << 4 | EXF_THETA) << 8)});
            if (!ex_share_common(cs.prog, items, prog)) return bail(fail(ctx, B200ZK_EINVAL, "pk_create", "malformed expression program"));
            pk->lookup_prog.push_back({off, (uint32_t)prog.size() - off});
        }
        // stack depth check
        int depth = 0, maxd = 0;
        for (uint32_t w : prog) {
            uint32_t op = w & 0xff;
            if (op <= EX_INSTANCE || op == EX_TMP) ++depth; else if (op == EX_ADD || op == EX_MUL || op == EX_FOLD || op == EX_SET1 || op == EX_GROUP) --depth;
            maxd = std::max(maxd, depth);
            if (depth < 0) return bail(fail(ctx, B200ZK_EINVAL, "pk_create", "malformed expression program"));
        }
        if (maxd > EX_STACK) return bail(fail(ctx, B200ZK_EINVAL, "pk_create", "expression too deep for the evaluator stack"));
        PK_TRY(dev_alloc(pk, &pk->d_prog, prog.size()));
        PK_CUDA(cudaMemcpyAsync(pk->d_prog, prog.data(), prog.size() * 4, cudaMemcpyHostToDevice, st));
        PK_TRY(dev_alloc(pk, &pk->d_consts, cs.consts.size()));
        PK_CUDA(cudaMemcpyAsync(pk->d_consts, cs.consts.data(), cs.consts.size() * sizeof(fe_t), cudaMemcpyHostToDevice, st));
        PK_TRY(dev_alloc(pk, &pk->d_q_adv, cs.adv_q.size())); PK_TRY(dev_alloc(pk, &pk->d_q_fix, cs.fix_q.size())); PK_TRY(dev_alloc(pk, &pk->d_q_inst, cs.inst_q.size()));
        PK_CUDA(cudaMemcpyAsync(pk->d_q_adv, cs.adv_q.data(), cs.adv_q.size() * 4, cudaMemcpyHostToDevice, st));
        PK_CUDA(cudaMemcpyAsync(pk->d_q_fix, cs.fix_q.data(), cs.fix_q.size() * 4, cudaMemcpyHostToDevice, st));
        PK_CUDA(cudaMemcpyAsync(pk->d_q_inst, cs.inst_q.data(), cs.inst_q.size() * 4, cudaMemcpyHostToDevice, st));
        PK_CUDA(cudaStreamSynchronize(st));
        PtrTab pt{cs.F, cs.A, cs.I};
        PK_TRY(dev_alloc(pk, &pk->d_ptrs, pt.total()));
    }
    {
        // early lookup cosets when their 4 L CL n elements are a modest share of the device (B200ZK_LOOKUP_EARLY=0/1 overrides)
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const size_t extra = 4 * (size_t)pk->L * pk->lk_cosets_n * pk->n * sizeof(fe_t);
        pk->lk_early = pk->L > 0 && extra <= free_b / 4;
        if (const char* e = getenv("B200ZK_LOOKUP_EARLY")) pk->lk_early = pk->L > 0 && e[0] != '0';
    }
    pk->arena_bytes = arena_need(pk);
    PK_TRY(dev_alloc(pk, &pk->arena, pk->arena_bytes));
    PK_CUDA(cudaStreamSynchronize(st));
    PK_CUDA(cudaGetLastError());
#undef PK_TRY
#undef PK_CUDA
    *out = pk;
    return B200ZK_OK;
}

static int32_t finish_proof(b200zk_pk* pk, int32_t rc, const std::vector<uint8_t>& proof, uint8_t* proof_out, size_t cap, size_t* proof_len) {
    // a rank that fails on its own (not the collective ConstraintSystemFailure / InstanceTooLarge) releases its peers
    if (rc != B200ZK_OK && rc != B200ZK_ESYNTH && pk->ctx->comm) pk->ctx->comm->abort();
    if (rc != B200ZK_OK) {
        // a failed proof may have left column uploads in flight: the caller's buffers are borrowed for the call only
        if (pk->copy_stream) cudaStreamSynchronize(pk->copy_stream);
        return rc;
    }
    if (proof_len) *proof_len = proof.size();
    if (proof.size() > cap) return fail(pk->ctx, B200ZK_EINVAL, "create_proof", "proof buffer too small");
    memcpy(proof_out, proof.data(), proof.size());
    return B200ZK_OK;
}

int32_t b200zk_create_proof(b200zk_pk* pk, const void* const* advice_columns, const void* const* instance_columns,
                            const uint32_t* instance_lens, const void* rng_wide, const void* transcript_repr,
                            uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
    if (!pk || !rng_wide || !transcript_repr || !proof_out || (pk->cs.A && !advice_columns) || (pk->cs.I && (!instance_columns || !instance_lens))) return B200ZK_EINVAL;
    std::vector<uint8_t> proof;
    int32_t rc = prove(pk, nullptr, false, advice_columns, instance_columns, instance_lens, rng_wide, false, HFr::from_limbs(transcript_repr), proof);
    return finish_proof(pk, rc, proof, proof_out, proof_cap, proof_len);
}

int32_t b200zk_create_proof_dev(b200zk_pk* pk, const void* d_advice, const void* const* instance_columns,
                                const uint32_t* instance_lens, const void* d_rng_wide, const void* transcript_repr,
                                uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
    if (!pk || !d_rng_wide || !transcript_repr || !proof_out || (pk->cs.A && !d_advice) || (pk->cs.I && (!instance_columns || !instance_lens))) return B200ZK_EINVAL;
    std::vector<uint8_t> proof;
    int32_t rc = prove(pk, (const fe_t*)d_advice, true, nullptr, instance_columns, instance_lens, d_rng_wide, true, HFr::from_limbs(transcript_repr), proof);
    return finish_proof(pk, rc, proof, proof_out, proof_cap, proof_len);
}

// ---- one create_proof over the GPUs of a group (b200zk_group_create): rank r = thread r on pks[r]'s device.
// Every rank gets the same inputs; the proof (identical on every rank) is rank 0's.
static int32_t group_prove(b200zk_group* g, b200zk_pk* const* pks, const void* const* d_advice_per_rank, const void* const* advice_columns,
                           const void* const* instance_columns, const uint32_t* instance_lens, const void* const* d_rng_per_rank,
                           const void* rng_wide, const void* transcript_repr, uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
    const uint32_t G = b200zk_group_size(g);
    if (!G || !pks || !transcript_repr || !proof_out) return B200ZK_EINVAL;
    for (uint32_t r = 0; r < G; ++r) {
        if (!pks[r] || pks[r]->ctx != b200zk_group_ctx(g, r)) return B200ZK_EINVAL;
        if (pks[r]->cs.I && (!instance_columns || !instance_lens)) return B200ZK_EINVAL;
        if (d_advice_per_rank ? (!d_rng_per_rank || (pks[r]->cs.A && !d_advice_per_rank[r]) || !d_rng_per_rank[r]) : (!rng_wide || (pks[r]->cs.A && !advice_columns))) return B200ZK_EINVAL;
    }
    b200zk_group_reset(g);
    std::vector<int32_t> rcs(G, B200ZK_OK);
    std::vector<std::vector<uint8_t>> proofs(G);
    const HFr repr = HFr::from_limbs(transcript_repr);
    auto body = [&](uint32_t r) {
        b200zk_pk* pk = pks[r];
        cudaSetDevice(pk->ctx->device);
        int32_t rc = d_advice_per_rank ? prove(pk, (const fe_t*)d_advice_per_rank[r], true, nullptr, instance_columns, instance_lens, d_rng_per_rank[r], true, repr, proofs[r])
                                       : prove(pk, nullptr, false, advice_columns, instance_columns, instance_lens, rng_wide, false, repr, proofs[r]);
        if (rc != B200ZK_OK && rc != B200ZK_ESYNTH) pk->ctx->comm->abort();
        if (rc != B200ZK_OK && pk->copy_stream) cudaStreamSynchronize(pk->copy_stream);
        rcs[r] = rc;
    };
    std::vector<std::thread> threads;
    for (uint32_t r = 1; r < G; ++r) threads.emplace_back(body, r);
    body(0);
    for (auto& t : threads) t.join();
    for (uint32_t r = 0; r < G; ++r) if (rcs[r] != B200ZK_OK) return rcs[r];
    for (uint32_t r = 1; r < G; ++r) if (proofs[r] != proofs[0]) return fail(pks[0]->ctx, B200ZK_ECUDA, "group_create_proof", "internal: ranks disagree on the proof");
    if (proof_len) *proof_len = proofs[0].size();
    if (proofs[0].size() > proof_cap) return fail(pks[0]->ctx, B200ZK_EINVAL, "create_proof", "proof buffer too small");
    memcpy(proof_out, proofs[0].data(), proofs[0].size());
    return B200ZK_OK;
}

int32_t b200zk_group_create_proof(b200zk_group* g, b200zk_pk* const* pks, const void* const* advice_columns, const void* const* instance_columns,
                                  const uint32_t* instance_lens, const void* rng_wide, const void* transcript_repr,
                                  uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
    return group_prove(g, pks, nullptr, advice_columns, instance_columns, instance_lens, nullptr, rng_wide, transcript_repr, proof_out, proof_cap, proof_len);
}

int32_t b200zk_group_create_proof_dev(b200zk_group* g, b200zk_pk* const* pks, const void* const* d_advice_per_rank, const void* const* instance_columns,
                                      const uint32_t* instance_lens, const void* const* d_rng_wide_per_rank, const void* transcript_repr,
                                      uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
    if (!d_advice_per_rank || !d_rng_wide_per_rank) return B200ZK_EINVAL;
    return group_prove(g, pks, d_advice_per_rank, nullptr, instance_columns, instance_lens, d_rng_wide_per_rank, nullptr, transcript_repr, proof_out, proof_cap, proof_len);
}

}  // extern "C"
