// Host-side bn256 pairing for the verifier's final check (halo2_proofs v2023_02_02
// poly/kzg/strategy.rs `DualMSM::check`: e(left, [s]_2) = e(right, [1]_2), reached from
// `verify_proof` at /root/reference/src/circuits/utils.rs:56-63).  Two Miller loops and one
// final exponentiation per proof — O(1) host work like the transcript, never a kernel substitute.
//
// Tower: Fq2 = Fq[u]/(u^2 + 1); Fq12 = Fq2[w]/(w^6 - xi), xi = 9 + u, kept as six Fq2 coefficients;
// G2 is the D-type twist y^2 = x^3 + 3/xi with the untwist (x, y) -> (x w^2, y w^3).  The pairing is
// the ate pairing f_{t-1,Q}(P)^((q^12-1)/r) with t - 1 = 6 x^2 (x = 4965661367192848881): a power of
// the optimal ate pairing halo2curves computes, so every product-equals-one check has the same
// outcome, and it needs no Frobenius constants.  The final exponentiation is a plain square-and-
// multiply by the 2790-bit exponent (about 20 ms on one core; a verification runs it once).
#pragma once
#include "host_field.hpp"

namespace b200zk {
namespace host {

struct F2 { HFq a, b; };
inline F2 f2_zero() { return {HFq::zero(), HFq::zero()}; }
inline F2 f2_one() { return {HFq::one(), HFq::zero()}; }
inline bool f2_is_zero(const F2& x) { return x.a.is_zero() && x.b.is_zero(); }
inline bool f2_eq(const F2& x, const F2& y) { return x.a == y.a && x.b == y.b; }
inline F2 f2_add(const F2& x, const F2& y) { return {x.a + y.a, x.b + y.b}; }
inline F2 f2_sub(const F2& x, const F2& y) { return {x.a - y.a, x.b - y.b}; }
inline F2 f2_neg(const F2& x) { return {x.a.neg(), x.b.neg()}; }
inline F2 f2_mul(const F2& x, const F2& y) {
    HFq t0 = x.a * y.a, t1 = x.b * y.b;
    return {t0 - t1, (x.a + x.b) * (y.a + y.b) - t0 - t1};
}
inline F2 f2_sqr(const F2& x) { return {(x.a + x.b) * (x.a - x.b), (x.a * x.b).dbl()}; }
inline F2 f2_scale(const F2& x, const HFq& s) { return {x.a * s, x.b * s}; }
inline F2 f2_inv(const F2& x) {
    HFq d = (x.a.sqr() + x.b.sqr()).inv();
    return {x.a * d, (x.b * d).neg()};
}
inline F2 f2_mul_xi(const F2& x) {          // (a + b u)(9 + u)
    HFq a8 = x.a.dbl().dbl().dbl(), b8 = x.b.dbl().dbl().dbl();
    return {a8 + x.a - x.b, b8 + x.b + x.a};
}

struct F12 { F2 c[6]; };
inline F12 f12_one() { F12 r; for (int i = 0; i < 6; ++i) r.c[i] = f2_zero(); r.c[0] = f2_one(); return r; }
inline bool f12_is_one(const F12& x) {
    if (!f2_eq(x.c[0], f2_one())) return false;
    for (int i = 1; i < 6; ++i) if (!f2_is_zero(x.c[i])) return false;
    return true;
}
inline F12 f12_mul(const F12& x, const F12& y) {
    F2 t[11];
    for (int i = 0; i < 11; ++i) t[i] = f2_zero();
    for (int i = 0; i < 6; ++i) {
        if (f2_is_zero(x.c[i])) continue;
        for (int j = 0; j < 6; ++j) {
            if (f2_is_zero(y.c[j])) continue;
            t[i + j] = f2_add(t[i + j], f2_mul(x.c[i], y.c[j]));
        }
    }
    F12 r;
    for (int i = 0; i < 6; ++i) r.c[i] = i < 5 ? f2_add(t[i], f2_mul_xi(t[i + 6])) : t[i];
    return r;
}

// G2 affine on the twist; identity = (0, 0) like halo2curves' G2Affine
struct G2A { F2 x, y; };
inline bool g2_is_identity(const G2A& p) { return f2_is_zero(p.x) && f2_is_zero(p.y); }
inline G2A g2_from_limbs(const void* p128) {
    const uint64_t* p = (const uint64_t*)p128;
    return {{HFq::from_limbs(p), HFq::from_limbs(p + 4)}, {HFq::from_limbs(p + 8), HFq::from_limbs(p + 12)}};
}
inline void g2_store(const G2A& g, void* p128) {
    uint64_t* p = (uint64_t*)p128;
    g.x.a.store(p); g.x.b.store(p + 4); g.y.a.store(p + 8); g.y.b.store(p + 12);
}
inline G2A g2_generator() {                  // EIP-197 / halo2curves bn256::G2Affine::generator()
    static const uint64_t c[4][4] = {
        {0x46debd5cd992f6edULL, 0x674322d4f75edaddULL, 0x426a00665e5c4479ULL, 0x1800deef121f1e76ULL},
        {0x97e485b7aef312c2ULL, 0xf1aa493335a9e712ULL, 0x7260bfb731fb5d25ULL, 0x198e9393920d483aULL},
        {0x4ce6cc0166fa7daaULL, 0xe3d1e7690c43d37bULL, 0x4aab71808dcb408fULL, 0x12c85ea5db8c6debULL},
        {0x55acdadcd122975bULL, 0xbc4b313370b38ef3ULL, 0xec9e99ad690c3395ULL, 0x090689d0585ff075ULL}};
    return {{HFq::from_canonical(c[0]), HFq::from_canonical(c[1])}, {HFq::from_canonical(c[2]), HFq::from_canonical(c[3])}};
}
inline bool g2_on_curve(const G2A& p) {
    if (g2_is_identity(p)) return true;
    F2 b = f2_scale(f2_inv({HFq::from_u64(9), HFq::one()}), HFq::from_u64(3));
    return f2_eq(f2_sub(f2_sqr(p.y), f2_mul(f2_sqr(p.x), p.x)), b);
}
// slope of the chord / tangent; *vertical is set when the sum is the identity
inline F2 g2_slope(const G2A& t, const G2A& q, bool* vertical) {
    *vertical = false;
    if (f2_eq(t.x, q.x)) {
        if (!f2_eq(t.y, q.y) || f2_is_zero(t.y)) { *vertical = true; return f2_zero(); }
        F2 xx = f2_sqr(t.x);
        return f2_mul(f2_add(f2_add(xx, xx), xx), f2_inv(f2_add(t.y, t.y)));
    }
    return f2_mul(f2_sub(q.y, t.y), f2_inv(f2_sub(q.x, t.x)));
}
inline G2A g2_add(const G2A& t, const G2A& q) {
    if (g2_is_identity(t)) return q;
    if (g2_is_identity(q)) return t;
    bool vert;
    F2 lam = g2_slope(t, q, &vert);
    if (vert) return {f2_zero(), f2_zero()};
    F2 x3 = f2_sub(f2_sub(f2_sqr(lam), t.x), q.x);
    return {x3, f2_sub(f2_mul(lam, f2_sub(t.x, x3)), t.y)};
}
inline G2A g2_mul(const G2A& p, const HFr& s) {
    uint64_t e[4]; s.to_canonical(e);
    G2A acc = {f2_zero(), f2_zero()};
    for (int i = 255; i >= 0; --i) {
        acc = g2_add(acc, acc);
        if ((e[i >> 6] >> (i & 63)) & 1) acc = g2_add(acc, p);
    }
    return acc;
}

// one Miller step: multiply f by the line through t and q (tangent when t == q) evaluated at the
// G1 point (px, py), and replace t by t + q:   l = py - lam px w + (lam xT - yT) w^3
inline void miller_step(F12& f, G2A& t, const G2A& q, const HFq& px, const HFq& py) {
    bool vert;
    F2 lam = g2_slope(t, q, &vert);
    if (vert) {                                      // x - xT w^2 (cannot occur for points of order r inside the loop)
        F12 l = f12_one();
        l.c[0] = {px, HFq::zero()}; l.c[2] = f2_neg(t.x);
        f = f12_mul(f, l);
        t = {f2_zero(), f2_zero()};
        return;
    }
    F12 l;
    for (int i = 0; i < 6; ++i) l.c[i] = f2_zero();
    l.c[0] = {py, HFq::zero()};
    l.c[1] = f2_neg(f2_scale(lam, px));
    l.c[3] = f2_sub(f2_mul(lam, t.x), t.y);
    f = f12_mul(f, l);
    F2 x3 = f2_sub(f2_sub(f2_sqr(lam), t.x), q.x);
    t = {x3, f2_sub(f2_mul(lam, f2_sub(t.x, x3)), t.y)};
}
inline F12 miller_loop(const G2A& q, const HAffine& p) {
    static const uint64_t T[2] = {0xf83e9682e87cfd46ULL, 0x6f4d8248eeb859fbULL};      // t - 1 = 6 x^2, 127 bits
    F12 f = f12_one();
    G2A t = q;
    for (int i = 125; i >= 0; --i) {
        f = f12_mul(f, f);
        G2A tt = t;
        miller_step(f, t, tt, p.x, p.y);
        if ((T[i >> 6] >> (i & 63)) & 1) miller_step(f, t, q, p.x, p.y);
    }
    return f;
}
inline F12 final_exponentiation(const F12& f) {
    static const uint64_t E[44] = {                                                   // (q^12 - 1) / r
#include "pairing_exponent.inc"
    };
    F12 acc = f12_one();
    bool started = false;
    for (int i = 44 * 64 - 1; i >= 0; --i) {
        if (started) acc = f12_mul(acc, acc);
        if ((E[i >> 6] >> (i & 63)) & 1) { acc = started ? f12_mul(acc, f) : f; started = true; }
    }
    return acc;
}
// prod e(p_i, q_i) == 1 ?   pairs with an identity on either side contribute 1
inline bool pairing_product_is_one(const HAffine* ps, const G2A* qs, size_t count) {
    F12 f = f12_one();
    for (size_t i = 0; i < count; ++i) {
        if ((ps[i].x.is_zero() && ps[i].y.is_zero()) || g2_is_identity(qs[i])) continue;
        f = f12_mul(f, miller_loop(qs[i], ps[i]));
    }
    return f12_is_one(final_exponentiation(f));
}

}  // namespace host
}  // namespace b200zk
