// O(n) field-vector primitives used around the MSM / NTT kernels:
//   batch inversion          ff::BatchInvert / poly::batch_invert_assigned (halo2_proofs
//                            v2023_02_02 src/poly.rs, used by permutation/lookup provers)
//   linear recurrences       arithmetic::eval_polynomial (Horner) and kate_division
//   prefix products          the grand-product columns of plonk/permutation/prover.rs and
//                            plonk/lookup/prover.rs
// All reached from create_proof at /root/reference/src/circuits/utils.rs:40-48.
// Written as per-thread bodies over grid-stride lanes so that every global access is
// coalesced: lane t owns elements t, t + T, t + 2T, ...
#pragma once
#include "field.cuh"
#include "blockexec.cuh"

namespace b200zk {

// ---- batch inversion (Montgomery's trick per lane) ------------------------------
// Lane t of T: elements i = t + j*T.  scratch holds the running prefix products.
// Zeros are skipped (stay zero), as halo2's BatchInvert does.
template <class F> ZK_D void batch_invert_lane(fe_t* a, fe_t* scratch, size_t n, size_t t, size_t T) {
    fe_t acc = F::one();
    for (size_t i = t; i < n; i += T) {
        fe_t v = a[i];
        scratch[i] = acc;
        if (!F::is_zero(v)) acc = F::mul(acc, v);
    }
    acc = F::inv_gcd(acc);                                  // division-step inversion: ~6x shorter than the Fermat chain a lane would wait for
    size_t cnt = t < n ? (n - t + T - 1) / T : 0;
    for (size_t j = cnt; j-- > 0;) {
        size_t i = t + j * T;
        fe_t v = a[i];
        if (F::is_zero(v)) continue;
        fe_t pre = scratch[i];
        a[i] = F::mul(acc, pre);
        acc = F::mul(acc, v);
    }
}

// ---- first-order linear recurrence  y[i] = a[i] + b * y[i+1],  y[n] = 0 ----------
// (Horner: eval_polynomial(a, b) = y[0]; kate_division(a, b)[i] = y[i+1].)
// Contiguous chunks: chunk c covers [c*m, min(n, (c+1)*m)).
// Pass 1: local recurrence with zero carry-in, written to y; chunk head saved.
ZK_D void recur_local_chunk(const fe_t* a, fe_t* y, size_t n, size_t m, size_t c, const fe_t& b, fe_t* heads) {
    size_t lo = c * m, hi = lo + m < n ? lo + m : n;
    if (lo >= n) return;
    fe_t acc = Fr::zero();
    for (size_t i = hi; i-- > lo;) {
        fe_t v = a[i];
        acc = Fr::add(Fr::mul(acc, b), v);
        if (y) y[i] = acc;
    }
    heads[c] = acc;
}
// Pass 2 (one block of T threads, T a power of two): right-to-left scan of the affine maps
// x -> heads[c] + B x with B = b^m (chunks are treated as zero-padded to full length, which
// leaves every y unchanged).  On exit carries[c] = true y[(c+1)*m] and heads[0] = y[0].
// sm: 2*T field elements.
ZK_D void recur_carries_block(fe_t* heads, fe_t* carries, size_t n, size_t m, const fe_t& b, uint32_t T, fe_t* sm) {
    const size_t C = (n + m - 1) / m;
    const size_t r = (C + T - 1) / T;
    const fe_t B = Fr::pow_u64(b, m);
    // A: thread t owns chunks [t*r, (t+1)*r); local aggregate H_t = sum_j heads[lo+j] B^j
    ZK_PHASE_BEGIN(tid, T)
    size_t lo = (size_t)tid * r, hi = lo + r < C ? lo + r : C;
    fe_t acc = Fr::zero();
    for (size_t c = hi; c-- > lo;) { fe_t h = heads[c]; acc = Fr::add(Fr::mul(acc, B), h); }
    sm[tid] = acc;
    ZK_PHASE_END
    // B: suffix scan  S_t = H_t + B^r H_{t+1} + B^{2r} H_{t+2} + ...   (ping-pong buffers)
    fe_t step = Fr::pow_u64(B, r);
    uint32_t cur = 0;
    for (uint32_t d = 1; d < T; d <<= 1) {
        ZK_PHASE_BEGIN(tid, T)
        fe_t v = sm[cur * T + tid];
        if (tid + d < T) { fe_t o = sm[cur * T + tid + d]; v = Fr::add(v, Fr::mul(step, o)); }
        sm[(cur ^ 1) * T + tid] = v;
        ZK_PHASE_END
        step = Fr::sqr(step);
        cur ^= 1;
    }
    // C: walk the owned chunks right to left from the carry entering the thread's range
    ZK_PHASE_BEGIN(tid, T)
    size_t lo = (size_t)tid * r, hi = lo + r < C ? lo + r : C;
    fe_t carry = (tid + 1 < T) ? sm[cur * T + tid + 1] : Fr::zero();
    if (lo >= C) carry = Fr::zero();
    for (size_t c = hi; c-- > lo;) {
        carries[c] = carry;
        fe_t h = heads[c];
        carry = Fr::add(h, Fr::mul(B, carry));
        heads[c] = carry;
    }
    ZK_PHASE_END
}
// Pass 3: y[i] += b^(hi - i) * carry[c]
ZK_D void recur_apply_chunk(fe_t* y, size_t n, size_t m, size_t c, const fe_t& b, const fe_t* carries) {
    size_t lo = c * m, hi = lo + m < n ? lo + m : n;   // carry is zero for a partial (= last) chunk
    if (lo >= n) return;
    fe_t carry = carries[c];
    if (Fr::is_zero(carry)) return;
    fe_t f = carry;
    for (size_t i = hi; i-- > lo;) {
        f = Fr::mul(f, b);
        fe_t v = y[i];
        y[i] = Fr::add(v, f);
    }
}

// ---- prefix product  z[0] = z0, z[i+1] = z[i] * p[i]  (i < n - 1 written; z has n entries) ----
// Same chunking; pass 1 computes chunk products, pass 2 the exclusive scan of chunk products
// (seeded with z0), pass 3 writes z.
ZK_D void prodscan_chunk_product(const fe_t* p, size_t n, size_t m, size_t c, fe_t* prods) {
    size_t lo = c * m, hi = lo + m < n ? lo + m : n;
    if (lo >= n) return;
    fe_t acc = p[lo];
    for (size_t i = lo + 1; i < hi; ++i) { fe_t v = p[i]; acc = Fr::mul(acc, v); }
    prods[c] = acc;
}
// One block of T threads (power of two): exclusive scan of the C chunk products seeded with z0.
// sm: 2*T field elements.
ZK_D void prodscan_carries_block(fe_t* prods, size_t C, const fe_t& z0, uint32_t T, fe_t* sm) {
    const size_t r = (C + T - 1) / T;
    ZK_PHASE_BEGIN(tid, T)
    size_t lo = (size_t)tid * r, hi = lo + r < C ? lo + r : C;
    fe_t acc = Fr::one();
    for (size_t c = lo; c < hi; ++c) { fe_t v = prods[c]; acc = Fr::mul(acc, v); }
    sm[tid] = acc;
    ZK_PHASE_END
    uint32_t cur = 0;
    for (uint32_t d = 1; d < T; d <<= 1) {              // inclusive prefix scan
        ZK_PHASE_BEGIN(tid, T)
        fe_t v = sm[cur * T + tid];
        if (tid >= d) { fe_t o = sm[cur * T + tid - d]; v = Fr::mul(o, v); }
        sm[(cur ^ 1) * T + tid] = v;
        ZK_PHASE_END
        cur ^= 1;
    }
    ZK_PHASE_BEGIN(tid, T)
    size_t lo = (size_t)tid * r, hi = lo + r < C ? lo + r : C;
    fe_t run = z0;
    if (tid > 0) { fe_t o = sm[cur * T + tid - 1]; run = Fr::mul(run, o); }
    for (size_t c = lo; c < hi; ++c) { fe_t v = prods[c]; prods[c] = run; run = Fr::mul(run, v); }
    ZK_PHASE_END
}
ZK_D void prodscan_write_chunk(const fe_t* p, fe_t* z, size_t n, size_t m, size_t c, const fe_t* prods) {
    size_t lo = c * m, hi = lo + m < n ? lo + m : n;
    if (lo >= n) return;
    fe_t run = prods[c];
    for (size_t i = lo; i < hi; ++i) {
        fe_t v = p[i];                                  // read before write: z may alias p
        z[i] = run;
        run = Fr::mul(run, v);
    }
}

}  // namespace b200zk
