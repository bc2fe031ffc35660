// Host-side bn256 arithmetic for the product's own control path: domain constants,
// transcript challenges, the last few group operations of an MSM and the affine
// normalisation of commitments.  O(1)..O(W) work per call — never a substitute for
// a kernel.  4 x u64 Montgomery limbs, the same memory format as the device's
// 8 x u32 (little-endian), i.e. halo2curves' Fr / Fq.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace b200zk {
namespace host {

typedef unsigned __int128 u128;

struct FrTag {
    static constexpr uint64_t P[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    static constexpr uint64_t INV = 0xc2e1f593efffffffULL;
    static constexpr uint64_t R1[4] = {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL};
    static constexpr uint64_t R2[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL};
    static constexpr uint64_t R3[4] = {0x5e94d8e1b4bf0040ULL, 0x2a489cbe1cfbb6b8ULL, 0x893cc664a19fcfedULL, 0x0cf8594b7fcc657cULL};
};
struct FqTag {
    static constexpr uint64_t P[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    static constexpr uint64_t INV = 0x87d20782e4866389ULL;
    static constexpr uint64_t R1[4] = {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL};
    static constexpr uint64_t R2[4] = {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL};
    static constexpr uint64_t R3[4] = {0xb1cd6dafda1530dfULL, 0x62f210e6a7283db6ULL, 0xef7f0b0c0ada0afbULL, 0x20fd6e902d592544ULL};
};

template <class Tg> struct HF {
    uint64_t v[4];

    static HF zero() { HF r; r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0; return r; }
    static HF one() { HF r; memcpy(r.v, Tg::R1, 32); return r; }
    static HF from_limbs(const void* p) { HF r; memcpy(r.v, p, 32); return r; }
    void store(void* p) const { memcpy(p, v, 32); }
    bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
    bool operator==(const HF& o) const { return memcmp(v, o.v, 32) == 0; }
    bool operator!=(const HF& o) const { return !(*this == o); }

    static bool ge_p(const uint64_t a[4]) {
        for (int i = 3; i >= 0; --i) { if (a[i] != Tg::P[i]) return a[i] > Tg::P[i]; }
        return true;
    }
    static void sub_p(uint64_t a[4]) {
        uint64_t br = 0;
        for (int i = 0; i < 4; ++i) { u128 d = (u128)a[i] - Tg::P[i] - br; a[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
    }
    HF operator+(const HF& o) const {
        HF r; u128 c = 0;
        for (int i = 0; i < 4; ++i) { c += (u128)v[i] + o.v[i]; r.v[i] = (uint64_t)c; c >>= 64; }
        if (ge_p(r.v)) sub_p(r.v);
        return r;
    }
    HF operator-(const HF& o) const {
        HF r; uint64_t br = 0;
        for (int i = 0; i < 4; ++i) { u128 d = (u128)v[i] - o.v[i] - br; r.v[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
        if (br) { u128 c = 0; for (int i = 0; i < 4; ++i) { c += (u128)r.v[i] + Tg::P[i]; r.v[i] = (uint64_t)c; c >>= 64; } }
        return r;
    }
    HF neg() const { return zero() - *this; }
    HF dbl() const { return *this + *this; }
    // separated-operand-scanning Montgomery product: full 512-bit product, then 4 reduction rounds
    HF operator*(const HF& o) const {
        uint64_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) {
            u128 c = 0;
            for (int j = 0; j < 4; ++j) { c += (u128)v[i] * o.v[j] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; }
            t[i + 4] = (uint64_t)c;
        }
        uint64_t top = 0;
        for (int i = 0; i < 4; ++i) {
            uint64_t m = t[i] * Tg::INV;
            u128 c = 0;
            for (int j = 0; j < 4; ++j) { c += (u128)m * Tg::P[j] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; }
            for (int j = i + 4; j < 8 && c; ++j) { c += t[j]; t[j] = (uint64_t)c; c >>= 64; }
            top += (uint64_t)c;
        }
        HF r; memcpy(r.v, t + 4, 32);
        if (top || ge_p(r.v)) sub_p(r.v);
        return r;
    }
    HF sqr() const { return *this * *this; }
    HF pow(const uint64_t e[4]) const {
        HF acc = one();
        for (int i = 255; i >= 0; --i) { acc = acc.sqr(); if ((e[i >> 6] >> (i & 63)) & 1) acc = acc * *this; }
        return acc;
    }
    HF pow_u64(uint64_t e) const { uint64_t t[4] = {e, 0, 0, 0}; return pow(t); }
    HF inv() const {                                    // Fermat, inv(0) = 0
        uint64_t e[4] = {Tg::P[0] - 2, Tg::P[1], Tg::P[2], Tg::P[3]};
        return pow(e);
    }
    static HF from_u64(uint64_t x) { HF a = zero(); a.v[0] = x; HF r2; memcpy(r2.v, Tg::R2, 32); return a * r2; }
    static HF from_canonical(const uint64_t x[4]) { HF a; memcpy(a.v, x, 32); HF r2; memcpy(r2.v, Tg::R2, 32); return a * r2; }
    void to_canonical(uint64_t out[4]) const { HF o = zero(); o.v[0] = 1; HF r = *this * o; memcpy(out, r.v, 32); }
    // 512-bit little-endian integer mod p (halo2curves from_bytes_wide / from_u512)
    static HF from_u512(const uint64_t x[8]) {
        HF lo, hi, r2, r3; memcpy(lo.v, x, 32); memcpy(hi.v, x + 4, 32); memcpy(r2.v, Tg::R2, 32); memcpy(r3.v, Tg::R3, 32);
        return lo * r2 + hi * r3;
    }
};

typedef HF<FrTag> HFr;
typedef HF<FqTag> HFq;

// Fr constants (halo2curves bn256/fr.rs): two-adicity 28, generator 7
inline HFr fr_root_of_unity() {
    static const uint64_t c[4] = {0xd34f1ed960c37c9cULL, 0x3215cf6dd39329c8ULL, 0x98865ea93dd31f74ULL, 0x03ddb9f5166d18b7ULL};
    return HFr::from_canonical(c);
}
inline HFr fr_zeta() {
    static const uint64_t c[4] = {0xb8ca0b2d36636f23ULL, 0xcc37a73fec2bc5e9ULL, 0x048b6e193fd84104ULL, 0x30644e72e131a029ULL};
    return HFr::from_canonical(c);
}
inline HFr fr_delta() {
    static const uint64_t c[4] = {0x870e56bbe533e9a2ULL, 0x5b5f898e5e963f25ULL, 0x64ec26aad4c86e71ULL, 0x09226b6e22c6f0caULL};
    return HFr::from_canonical(c);
}
static const unsigned FR_TWO_ADICITY = 28;

// ---- G1 on the host (XYZZ in, affine / Jacobian out) ----
struct HXyzz { HFq x, y, zz, zzz; };
struct HAffine { HFq x, y; };

inline HXyzz hx_identity() { return {HFq::zero(), HFq::zero(), HFq::zero(), HFq::zero()}; }
inline bool hx_is_identity(const HXyzz& p) { return p.zz.is_zero(); }
inline HXyzz hx_dbl(const HXyzz& p) {
    if (hx_is_identity(p)) return p;
    HFq u = p.y.dbl(), v = u.sqr(), w = u * v, s = p.x * v, xx = p.x.sqr(), m = xx.dbl() + xx;
    HXyzz r;
    r.x = m.sqr() - s.dbl();
    r.y = m * (s - r.x) - w * p.y;
    r.zz = v * p.zz; r.zzz = w * p.zzz;
    return r;
}
inline HXyzz hx_add(const HXyzz& a, const HXyzz& b) {
    if (hx_is_identity(b)) return a;
    if (hx_is_identity(a)) return b;
    HFq u1 = a.x * b.zz, u2 = b.x * a.zz, s1 = a.y * b.zzz, s2 = b.y * a.zzz;
    HFq p = u2 - u1, r = s2 - s1;
    if (p.is_zero()) return r.is_zero() ? hx_dbl(a) : hx_identity();
    HFq pp = p.sqr(), ppp = p * pp, q = u1 * pp;
    HXyzz o;
    o.x = r.sqr() - ppp - q.dbl();
    o.y = r * (q - o.x) - s1 * ppp;
    o.zz = a.zz * b.zz * pp; o.zzz = a.zzz * b.zzz * ppp;
    return o;
}
inline HAffine hx_to_affine(const HXyzz& p) {
    if (hx_is_identity(p)) return {HFq::zero(), HFq::zero()};
    HFq zi3 = p.zzz.inv();                              // 1/Z^3
    HFq zi = zi3 * p.zz, zi2 = zi.sqr();                // ZZ/ZZZ = 1/Z ; 1/Z^2 = 1/ZZ
    return {p.x * zi2, p.y * zi3};
}
// Curve::batch_normalize: one field inversion for the whole batch (Montgomery's trick)
inline void hx_batch_to_affine(const HXyzz* in, size_t n, HAffine* out) {
    std::vector<HFq> pre(n);
    HFq acc = HFq::one();
    for (size_t i = 0; i < n; ++i) { pre[i] = acc; if (!hx_is_identity(in[i])) acc = acc * in[i].zzz; }
    acc = acc.inv();
    for (size_t i = n; i-- > 0;) {
        if (hx_is_identity(in[i])) { out[i] = {HFq::zero(), HFq::zero()}; continue; }
        HFq zi3 = acc * pre[i];
        acc = acc * in[i].zzz;
        HFq zi = zi3 * in[i].zz, zi2 = zi.sqr();
        out[i] = {in[i].x * zi2, in[i].y * zi3};
    }
}
inline HXyzz hx_from_affine(const HAffine& a) {
    if (a.x.is_zero() && a.y.is_zero()) return hx_identity();
    return {a.x, a.y, HFq::one(), HFq::one()};
}

}  // namespace host
}  // namespace b200zk
