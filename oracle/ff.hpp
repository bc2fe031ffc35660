// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
// PARITY UNPINNED against the Rust reference: neither rustc nor the halo2 /
// halo2curves crates exist in this environment, so this file restates the
// published algorithm of
//   halo2curves 0.3.1  src/bn256/{fr,fq}.rs + src/derive/field.rs
// (4 x u64 little-endian limbs, Montgomery form with R = 2^256), which the
// reference reaches through `halo2_proofs::halo2curves::bn256::{Fr as Fp, ...}`
// at /root/reference/src/circuits/utils.rs:2.  It is pinned instead against
// Python big-int KATs (oracle/pyref.py, tests/golden/*.json).
#pragma once
#include <cstdint>
#include <cstring>
#include <array>
#include <vector>

namespace orc {

typedef unsigned __int128 u128;

struct FieldParams {
    uint64_t p[4];      // modulus, LE limbs
    uint64_t inv;       // -p^{-1} mod 2^64
    uint64_t r[4];      // R   mod p
    uint64_t r2[4];     // R^2 mod p
    uint64_t r3[4];     // R^3 mod p
};

// ---- raw 256-bit helpers --------------------------------------------------
static inline bool geq4(const uint64_t a[4], const uint64_t b[4]) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > b[i]) return true;
        if (a[i] < b[i]) return false;
    }
    return true;
}
static inline uint64_t add4(uint64_t o[4], const uint64_t a[4], const uint64_t b[4]) {
    u128 c = 0;
    for (int i = 0; i < 4; ++i) { c += (u128)a[i] + b[i]; o[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
static inline uint64_t sub4(uint64_t o[4], const uint64_t a[4], const uint64_t b[4]) {
    uint64_t borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - b[i] - borrow;
        o[i] = (uint64_t)d; borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
}

FieldParams make_params(const uint64_t p[4]);

template <int TAG> struct Fp {
    uint64_t l[4];
    static FieldParams P;      // filled by init_fields()

    static Fp zero() { Fp z; memset(z.l, 0, 32); return z; }
    static Fp one() { Fp z; memcpy(z.l, P.r, 32); return z; }
    static Fp from_raw(const uint64_t v[4]) {   // canonical integer -> Montgomery
        Fp a, r2; memcpy(a.l, v, 32); memcpy(r2.l, P.r2, 32); return a * r2;
    }
    static Fp from_u64(uint64_t v) { uint64_t t[4] = {v, 0, 0, 0}; return from_raw(t); }
    // halo2curves `from_bytes_wide` / `from_u512`: 512-bit LE integer mod p
    // = d0*R^2*R^-1 + d1*R^3*R^-1  (Montgomery products).
    static Fp from_u512(const uint64_t v[8]) {
        Fp d0, d1, r2, r3;
        memcpy(d0.l, v, 32); memcpy(d1.l, v + 4, 32);
        memcpy(r2.l, P.r2, 32); memcpy(r3.l, P.r3, 32);
        return d0 * r2 + d1 * r3;
    }
    void to_raw(uint64_t out[4]) const {        // Montgomery -> canonical
        Fp o; memset(o.l, 0, 32); o.l[0] = 1; Fp t = (*this).mont(o); memcpy(out, t.l, 32);
    }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
    bool operator==(const Fp& b) const { return memcmp(l, b.l, 32) == 0; }
    bool operator!=(const Fp& b) const { return !(*this == b); }

    Fp operator+(const Fp& b) const {
        Fp o; add4(o.l, l, b.l);                // p < 2^254, no carry out
        if (geq4(o.l, P.p)) sub4(o.l, o.l, P.p);
        return o;
    }
    Fp operator-(const Fp& b) const {
        Fp o; if (sub4(o.l, l, b.l)) add4(o.l, o.l, P.p); return o;
    }
    Fp neg() const { if (is_zero()) return *this; Fp o; sub4(o.l, P.p, l); return o; }
    Fp dbl() const { return *this + *this; }

    // CIOS Montgomery product a*b*R^-1 mod p
    Fp mont(const Fp& b) const {
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) {
            u128 c = 0;
            for (int j = 0; j < 4; ++j) {
                c += (u128)l[j] * b.l[i] + t[j];
                t[j] = (uint64_t)c; c >>= 64;
            }
            c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * P.inv;
            c = (u128)m * P.p[0] + t[0]; c >>= 64;
            for (int j = 1; j < 4; ++j) {
                c += (u128)m * P.p[j] + t[j];
                t[j - 1] = (uint64_t)c; c >>= 64;
            }
            c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
        }
        Fp o; memcpy(o.l, t, 32);
        if (t[4] || geq4(o.l, P.p)) sub4(o.l, o.l, P.p);
        return o;
    }
    Fp operator*(const Fp& b) const { return mont(b); }
    Fp sqr() const { return mont(*this); }
    Fp& operator+=(const Fp& b) { *this = *this + b; return *this; }
    Fp& operator-=(const Fp& b) { *this = *this - b; return *this; }
    Fp& operator*=(const Fp& b) { *this = *this * b; return *this; }

    Fp pow(const uint64_t e[4]) const {
        Fp acc = one();
        for (int i = 255; i >= 0; --i) {
            acc = acc.sqr();
            if ((e[i >> 6] >> (i & 63)) & 1) acc = acc * *this;
        }
        return acc;
    }
    Fp pow_u64(uint64_t e) const { uint64_t t[4] = {e, 0, 0, 0}; return pow(t); }
    // Fermat inverse; inv(0) = 0 (callers that need halo2's CtOption check is_zero first)
    Fp inv() const {
        uint64_t e[4]; uint64_t two[4] = {2, 0, 0, 0}; sub4(e, P.p, two); return pow(e);
    }
    // numeric order of canonical value (halo2curves `Ord for Fr`)
    static bool less(const Fp& a, const Fp& b) {
        uint64_t x[4], y[4]; a.to_raw(x); b.to_raw(y);
        for (int i = 3; i >= 0; --i) { if (x[i] != y[i]) return x[i] < y[i]; }
        return false;
    }
};

template <int TAG> FieldParams Fp<TAG>::P;

typedef Fp<0> Fr;   // scalar field  r
typedef Fp<1> Fq;   // base field    q

void init_fields();             // idempotent

// Fr two-adicity constants (halo2curves bn256/fr.rs)
extern Fr FR_ROOT_OF_UNITY;     // 7^((r-1)/2^28)
extern Fr FR_DELTA;             // 7^(2^28)
extern Fr FR_ZETA;              // primitive cube root of unity used as coset shift
static const unsigned FR_S = 28;

// in-place Montgomery batch inversion (zeros stay zero, as halo2's BatchInvert)
template <class F> void batch_invert(F* v, size_t n) {
    std::vector<F> pre(n);
    F acc = F::one();
    for (size_t i = 0; i < n; ++i) { pre[i] = acc; if (!v[i].is_zero()) acc = acc * v[i]; }
    acc = acc.inv();
    for (size_t i = n; i-- > 0;) {
        if (v[i].is_zero()) continue;
        F t = acc * pre[i]; acc = acc * v[i]; v[i] = t;
    }
}

}  // namespace orc
