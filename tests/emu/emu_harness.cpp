// TEST AID: compiles the product's device headers with a host compiler (carry flag
// and block phases emulated, see csrc/ptx_chain.cuh and csrc/blockexec.cuh) so the
// limb algorithms and kernel index logic can be checked against the oracle on a box
// without a GPU.  Never linked into the product library.
#include <vector>
#include <cstring>
#include "../../halo2-experiments_b200/csrc/field.cuh"
#include "../../halo2-experiments_b200/csrc/ntt.cuh"
#include "../../halo2-experiments_b200/csrc/ntt_plan.hpp"
#include "../../halo2-experiments_b200/csrc/msm.cuh"
#include "../../halo2-experiments_b200/csrc/msm_plan.hpp"
#include "../../halo2-experiments_b200/csrc/poly.cuh"
#include "../../halo2-experiments_b200/csrc/transcript.hpp"
#include "../../halo2-experiments_b200/csrc/prover_kernels.cuh"

using namespace b200zk;

extern "C" {

// op: 0 add 1 sub 2 mul 3 sqr 4 inv 5 from_mont 6 to_mont 7 inv_gcd ; which: 0 Fr 1 Fq
void emu_field_op(int which, int op, const fe_t* a, const fe_t* b, fe_t* out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        if (which == 0) {
            switch (op) {
                case 0: out[i] = Fr::add(a[i], b[i]); break;
                case 1: out[i] = Fr::sub(a[i], b[i]); break;
                case 2: out[i] = Fr::mul(a[i], b[i]); break;
                case 3: out[i] = Fr::sqr(a[i]); break;
                case 4: out[i] = Fr::inv(a[i]); break;
                case 5: out[i] = Fr::from_mont(a[i]); break;
                case 6: out[i] = Fr::to_mont(a[i]); break;
                case 7: out[i] = Fr::inv_gcd(a[i]); break;
            }
        } else {
            switch (op) {
                case 0: out[i] = Fq::add(a[i], b[i]); break;
                case 1: out[i] = Fq::sub(a[i], b[i]); break;
                case 2: out[i] = Fq::mul(a[i], b[i]); break;
                case 3: out[i] = Fq::sqr(a[i]); break;
                case 4: out[i] = Fq::inv(a[i]); break;
                case 5: out[i] = Fq::from_mont(a[i]); break;
                case 6: out[i] = Fq::to_mont(a[i]); break;
                case 7: out[i] = Fq::inv_gcd(a[i]); break;
            }
        }
    }
}

// x * w mod p through the constant-operand multiplier: w plain, wq = floor(w * 2^256 / p)
void emu_mul_shoup(int which, const fe_t* x, const fe_t* w, const fe_t* wq, fe_t* out, size_t n) {
    for (size_t i = 0; i < n; ++i) out[i] = which == 0 ? Fr::mul_shoup(x[i], w[i], wq[i]) : Fq::mul_shoup(x[i], w[i], wq[i]);
}

// the [0, 2p) forms used inside a transform pass: mode 0 mul_shoup_lazy(x, w, wq), 1 add_lazy(x, w), 2 sub_raw(x, w), 3 reduce_2p(x)
void emu_lazy_op(int which, int mode, const fe_t* x, const fe_t* w, const fe_t* wq, fe_t* out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        if (which == 0) {
            if (mode == 0) out[i] = Fr::mul_shoup_lazy(x[i], w[i], wq[i]);
            else if (mode == 1) out[i] = Fr::add_lazy(x[i], w[i]);
            else if (mode == 2) out[i] = Fr::sub_raw(x[i], w[i]);
            else { out[i] = x[i]; Fr::reduce_2p(out[i]); }
        } else {
            if (mode == 0) out[i] = Fq::mul_shoup_lazy(x[i], w[i], wq[i]);
            else if (mode == 1) out[i] = Fq::add_lazy(x[i], w[i]);
            else if (mode == 2) out[i] = Fq::sub_raw(x[i], w[i]);
            else { out[i] = x[i]; Fq::reduce_2p(out[i]); }
        }
    }
}

void emu_field_consts(int which, fe_t* one, fe_t* r2) {
    if (which == 0) { *one = Fr::one(); *r2 = Fr::r2(); } else { *one = Fq::one(); *r2 = Fq::r2(); }
}

// Full multi-pass NTT through ntt_pass_block, mirroring the launch sequence of ntt.cu.
// pre/post: 3 elements each or null.
void emu_ntt(const fe_t* in, uint32_t n_in, fe_t* out, uint32_t log_n, const fe_t* omega,
             uint32_t max_log_m, uint32_t max_log_tw, uint32_t tile_cap_log, uint32_t nthreads,
             const fe_t* pre, const fe_t* post) {
    NttShape s = ntt_plan_shape(log_n, max_log_m, max_log_tw, tile_cap_log);
    size_t N = (size_t)1 << log_n;
    std::vector<fe_t> roots((size_t)1 << (s.log_roots ? s.log_roots - 1 : 0));
    std::vector<fe_t> lo((size_t)1 << s.tw_lo_bits), hi(N >> s.tw_lo_bits ? N >> s.tw_lo_bits : 1);
    fe_t w_r = Fr::pow_u64(*omega, 1ull << (log_n - s.log_roots));
    for (uint32_t i = 0; i < roots.size(); ++i) ntt_pow_table_thread(roots.data(), w_r, i, 0);
    for (uint32_t i = 0; i < lo.size(); ++i) ntt_pow_table_thread(lo.data(), *omega, i, 0);
    for (uint32_t i = 0; i < hi.size(); ++i) ntt_pow_table_thread(hi.data(), *omega, i, s.tw_lo_bits);
    std::vector<fe_t> scratch(N);
    for (uint32_t p = 0; p < s.npass; ++p) {
        const NttPassShape& q = s.pass[p];
        NttPassArgs a{};
        a.in = p == 0 ? in : scratch.data();
        a.out = q.is_last ? out : scratch.data();
        a.log_n = log_n; a.log_m = q.log_m; a.log_l = q.log_l; a.log_tw = q.log_tw; a.is_last = q.is_last;
        a.log_m1 = q.log_m1; a.log_mid = q.log_mid;
        a.n_in = p == 0 ? n_in : (uint32_t)N;
        a.use_pre = (p == 0 && pre) ? 1 : 0; a.use_post = (q.is_last && post) ? 1 : 0;
        if (pre) memcpy(a.pre, pre, 96);
        if (post) memcpy(a.post, post, 96);
        a.roots = roots.data(); a.log_roots = s.log_roots;
        a.tw_lo = lo.data(); a.tw_hi = hi.data(); a.tw_lo_bits = s.tw_lo_bits;
        a.tw_shift = log_n - q.log_m - q.log_l; a.l_offset = 0;
        std::vector<half_t> sm((size_t)2 << (q.log_m + q.log_tw));
        for (uint32_t b = 0; b < q.blocks; ++b) ntt_pass_block(a, b, nthreads, sm.data());
    }
}


// The warp-level kernel's pass plan (ntt_plan.hpp): out[0] = npass, then 8 words per pass:
// log_m, log_l, log_tw, is_last, log_m1, log_mid, log_m3, blocks.  Returns 0 when log_n is not eligible.
int emu_ntt_warp_plan(uint32_t log_n, uint32_t* out) {
    if (!ntt_warp_eligible(log_n)) return 0;
    NttShape s = ntt_plan_shape_warp(log_n);
    out[0] = s.npass;
    for (uint32_t p = 0; p < s.npass; ++p) {
        const NttPassShape& q = s.pass[p];
        uint32_t* o = out + 1 + 8 * p;
        o[0] = q.log_m; o[1] = q.log_l; o[2] = q.log_tw; o[3] = q.is_last; o[4] = q.log_m1; o[5] = q.log_mid; o[6] = q.log_m3; o[7] = q.blocks;
    }
    return 1;
}

// Full MSM through the block programs of msm.cuh + the host finish of msm_plan.hpp.
// out_affine: 64 bytes (x, y Montgomery).
void emu_msm(const fe_t* scalars, const affine_t* bases, uint32_t n, int force_c, int fast_max, uint32_t seg_len, uint32_t* out_affine64) {
    MsmShape s = msm_plan_shape(n, force_c);
    std::vector<uint32_t> counts(s.nbuckets, 0), offsets(s.nbuckets + 1), cursor(s.nbuckets), entries((size_t)n * s.nwin + 1);
    std::vector<xyzz_t> buckets(s.nbuckets), partials((size_t)s.nwin << s.log_t), wsum(s.nwin);
    MsmArgs a{};
    a.scalars = scalars; a.bases = bases; a.n = n; a.c = s.c; a.nwin = s.nwin; a.log_t = s.log_t;
    a.pre = 0; a.pre_stride = 0; a.nbuckets = (uint32_t)s.nbuckets;
    a.counts = counts.data(); a.offsets = offsets.data(); a.cursor = cursor.data(); a.entries = entries.data();
    a.buckets = buckets.data(); a.partials = partials.data(); a.window_sums = wsum.data();
    for (uint32_t i = 0; i < n; ++i) msm_count_thread(a, i);
    const uint32_t B = (uint32_t)s.nbuckets;
    uint32_t maxcnt = 0;
    auto scan = [&](ScanArgs sa) {   // mirrors run_scan() of msm.cu with a small block size
        const uint32_t T = 4, items = T * MSM_SCAN_PER_THREAD;
        const uint32_t nb = (sa.total + items - 1) / items;
        std::vector<uint32_t> bs(nb), sm(2 * T);
        sa.blocksums = bs.data();
        for (uint32_t b = 0; b < nb; ++b) scan_blocksum_block(sa, b, T, sm.data());
        scan_top_block(sa, nb, T, sm.data());
        for (uint32_t b = 0; b < nb; ++b) scan_final_block(sa, b, T, sm.data());
    };
    scan(ScanArgs{a.counts, a.offsets, a.cursor, nullptr, &maxcnt, B, 0, 0});
    for (uint32_t i = 0; i < n; ++i) msm_scatter_thread(a, i);
    if ((int)maxcnt <= fast_max) {
        for (uint32_t k = 0; k < B; ++k) msm_accumulate_thread(a, k);
    } else {
        uint32_t L = seg_len;
        size_t t1_bound = (size_t)n * s.nwin / L + B, t2_bound = t1_bound / L + B;
        std::vector<uint32_t> toff1(B + 1), toff2(B + 1);
        std::vector<xyzz_t> p1(t1_bound), p2(t2_bound);
        scan(ScanArgs{a.counts, toff1.data(), nullptr, nullptr, nullptr, B, 0, L});
        MsmTaskArgs t1{a.offsets, toff1.data(), B, toff1.data() + B, L, nullptr, p1.data()};
        for (uint32_t t = 0; t < t1_bound; ++t) msm_accumulate_task_thread(a, t1, t);
        scan(ScanArgs{toff1.data(), toff2.data(), nullptr, nullptr, nullptr, B, 1, L});
        MsmTaskArgs t2{toff1.data(), toff2.data(), B, toff2.data() + B, L, p1.data(), p2.data()};
        for (uint32_t t = 0; t < t2_bound; ++t) msm_combine_task_thread(t2, t);
        MsmTaskArgs t3{toff2.data(), nullptr, B, nullptr, 0, p2.data(), a.buckets};
        for (uint32_t b = 0; b < B; ++b) msm_combine_bucket_thread(t3, b);
    }
    for (uint32_t g = 0; g < (s.nwin << s.log_t); ++g) msm_reduce_thread(a, g);
    std::vector<xyzz_t> smx(8);
    for (uint32_t j = 0; j < s.nwin; ++j) msm_fold_block(a, j, 8, smx.data());
    host::HAffine r = host::hx_to_affine(msm_finish(wsum.data(), s.nwin, s.c));
    memcpy(out_affine64, &r, 64);
}

// Fixed-base mode of msm.cu: table T[j*stride + i] = 2^(c j) * base_i (built here with the host
// curve code), one bucket set, bit-decomposition reduce.  n <= stride scalars.
void emu_msm_pre(const fe_t* scalars, const affine_t* bases, uint32_t n, uint32_t stride, uint32_t c, uint32_t seg_len, uint32_t* out_affine64) {
    MsmShape s{};
    s.c = c; s.nwin = (255 + c - 1) / c; s.log_t = (c - 1) > 2 ? (c - 1) - 2 : 0; s.nbuckets = (size_t)1 << (c - 1);
    std::vector<affine_t> table((size_t)s.nwin * stride);
    for (uint32_t i = 0; i < stride; ++i) {
        host::HAffine p; memcpy(&p, &bases[i], 64);
        host::HXyzz cur = host::hx_from_affine(p);
        for (uint32_t j = 0; j < s.nwin; ++j) {
            host::HAffine q = host::hx_to_affine(cur);
            memcpy(&table[(size_t)j * stride + i], &q, 64);
            for (uint32_t d = 0; d < c; ++d) cur = host::hx_dbl(cur);
        }
    }
    const uint32_t B = (uint32_t)s.nbuckets;
    std::vector<uint32_t> counts(B, 0), offsets(B + 1), cursor(B), entries((size_t)n * s.nwin + 1);
    std::vector<xyzz_t> buckets(B), partials((size_t)s.c << s.log_t), wsum(s.c);
    MsmArgs a{};
    a.scalars = scalars; a.bases = table.data(); a.n = n; a.c = s.c; a.nwin = s.nwin; a.log_t = s.log_t;
    a.pre = 1; a.pre_stride = stride; a.nbuckets = B;
    a.counts = counts.data(); a.offsets = offsets.data(); a.cursor = cursor.data(); a.entries = entries.data();
    a.buckets = buckets.data(); a.partials = partials.data(); a.window_sums = wsum.data();
    for (uint32_t i = 0; i < n; ++i) msm_count_thread(a, i);
    auto scan = [&](ScanArgs sa) {
        const uint32_t T = 4, items = T * MSM_SCAN_PER_THREAD;
        const uint32_t nb = (sa.total + items - 1) / items;
        std::vector<uint32_t> bs(nb), sm(2 * T);
        sa.blocksums = bs.data();
        for (uint32_t b = 0; b < nb; ++b) scan_blocksum_block(sa, b, T, sm.data());
        scan_top_block(sa, nb, T, sm.data());
        for (uint32_t b = 0; b < nb; ++b) scan_final_block(sa, b, T, sm.data());
    };
    scan(ScanArgs{a.counts, a.offsets, a.cursor, nullptr, nullptr, B, 0, 0});
    for (uint32_t i = 0; i < n; ++i) msm_scatter_thread(a, i);
    uint32_t L = seg_len;
    size_t t1_bound = (size_t)n * s.nwin / L + B, t2_bound = t1_bound / L + B;
    std::vector<uint32_t> toff1(B + 1), toff2(B + 1);
    std::vector<xyzz_t> p1(t1_bound), p2(t2_bound);
    scan(ScanArgs{a.counts, toff1.data(), nullptr, nullptr, nullptr, B, 0, L});
    MsmTaskArgs t1{a.offsets, toff1.data(), B, toff1.data() + B, L, nullptr, p1.data()};
    for (uint32_t t = 0; t < t1_bound; ++t) msm_accumulate_task_thread(a, t1, t);
    scan(ScanArgs{toff1.data(), toff2.data(), nullptr, nullptr, nullptr, B, 1, L});
    MsmTaskArgs t2{toff1.data(), toff2.data(), B, toff2.data() + B, L, p1.data(), p2.data()};
    for (uint32_t t = 0; t < t2_bound; ++t) msm_combine_task_thread(t2, t);
    MsmTaskArgs t3{toff2.data(), nullptr, B, nullptr, 0, p2.data(), a.buckets};
    for (uint32_t b = 0; b < B; ++b) msm_combine_bucket_thread(t3, b);
    for (uint32_t g = 0; g < (s.c << s.log_t); ++g) msm_reduce_bits_thread(a, g);
    std::vector<xyzz_t> smx(4);
    for (uint32_t t = 0; t < s.c; ++t) msm_fold_block(a, t, 4, smx.data());
    host::HAffine r = host::hx_to_affine(msm_finish_bits(wsum.data(), s.c));
    memcpy(out_affine64, &r, 64);
}

// host field self-check hooks: op 0 add 1 sub 2 mul 3 inv ; which 0 Fr 1 Fq
void emu_host_field_op(int which, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    using namespace host;
    for (size_t i = 0; i < n; ++i) {
        if (which == 0) {
            HFr x = HFr::from_limbs(a + 4 * i), y = HFr::from_limbs(b + 4 * i);
            HFr r = op == 0 ? x + y : op == 1 ? x - y : op == 2 ? x * y : x.inv();
            r.store(out + 4 * i);
        } else {
            HFq x = HFq::from_limbs(a + 4 * i), y = HFq::from_limbs(b + 4 * i);
            HFq r = op == 0 ? x + y : op == 1 ? x - y : op == 2 ? x * y : x.inv();
            r.store(out + 4 * i);
        }
    }
}
void emu_host_fr_consts(uint64_t* root, uint64_t* zeta, uint64_t* delta, uint64_t* wide_in8, uint64_t* wide_out) {
    host::fr_root_of_unity().store(root); host::fr_zeta().store(zeta); host::fr_delta().store(delta);
    host::HFr::from_u512(wide_in8).store(wide_out);
}


// transcript.hpp: replay a script of ops: 0 = squeeze (writes 32 B Montgomery challenge to out),
// 1 = common_scalar(next 32 B Montgomery), 2 = write_point(next 64 B affine Montgomery),
// 3 = write_scalar(next 32 B).  Returns the proof length; proof bytes copied to proof_out.
size_t emu_transcript(const uint8_t* ops, size_t nops, const uint64_t* data, uint64_t* out, uint8_t* proof_out) {
    host::Transcript t;
    size_t d = 0, o = 0;
    for (size_t i = 0; i < nops; ++i) {
        if (ops[i] == 0) { t.squeeze_challenge().store(out + 4 * o++); }
        else if (ops[i] == 1) { t.common_scalar(host::HFr::from_limbs(data + d)); d += 4; }
        else if (ops[i] == 2) { host::HAffine p{host::HFq::from_limbs(data + d), host::HFq::from_limbs(data + d + 4)}; t.write_point(p); d += 8; }
        else { t.write_scalar(host::HFr::from_limbs(data + d)); d += 4; }
    }
    memcpy(proof_out, t.proof().data(), t.proof().size());
    return t.proof().size();
}

// prover_kernels.cuh: Fr::random's from_u512 on the device path
void emu_from_u512(const uint32_t* wide, fe_t* out, size_t n) {
    for (size_t i = 0; i < n; ++i) out[i] = from_u512_row(wide + 16 * i);
}

// poly.cuh: launch sequences mirrored from poly.cu
void emu_batch_invert(int which, fe_t* a, size_t n, size_t lanes) {
    std::vector<fe_t> scratch(n ? n : 1);
    for (size_t t = 0; t < lanes; ++t) {
        if (which == 0) batch_invert_lane<Fr>(a, scratch.data(), n, t, lanes); else batch_invert_lane<Fq>(a, scratch.data(), n, t, lanes);
    }
}
// y may be null; head_out receives y[0]
void emu_recurrence(const fe_t* a, fe_t* y, size_t n, size_t m, uint32_t T, const fe_t* b, fe_t* head_out) {
    size_t C = (n + m - 1) / m;
    std::vector<fe_t> heads(C), carries(C), sm(2 * T);
    for (size_t c = 0; c < C; ++c) recur_local_chunk(a, y, n, m, c, *b, heads.data());
    recur_carries_block(heads.data(), carries.data(), n, m, *b, T, sm.data());
    if (y) for (size_t c = 0; c < C; ++c) recur_apply_chunk(y, n, m, c, *b, carries.data());
    *head_out = heads[0];
}
void emu_prefix_product(const fe_t* p, fe_t* z, size_t n, size_t m, uint32_t T, const fe_t* z0) {
    size_t C = (n + m - 1) / m;
    std::vector<fe_t> prods(C), sm(2 * T);
    for (size_t c = 0; c < C; ++c) prodscan_chunk_product(p, n, m, c, prods.data());
    prodscan_carries_block(prods.data(), C, *z0, T, sm.data());
    for (size_t c = 0; c < C; ++c) prodscan_write_chunk(p, z, n, m, c, prods.data());
}

// prover_kernels.cuh: the quotient's coset machinery (vanishing::construct without the extended iNTT)
void emu_lookup_extrapolate(fe_t* g, const fe_t* inv_pow, const fe_t* lambda, const fe_t* tinv, uint32_t CL, uint32_t C, size_t n) {
    for (size_t r = 0; r < n; ++r) lookup_extrapolate_row(g, inv_pow, lambda, tinv, CL, C, n, r);
}
void emu_coset_interpolate(const fe_t* g, const fe_t* inv_pow, const fe_t* vinv, uint32_t C, size_t n, fe_t* out, const fe_t* extra) {
    for (size_t r = 0; r < n; ++r) coset_interpolate_row(g, inv_pow, vinv, C, n, out, r, extra);
}

}  // extern "C"
