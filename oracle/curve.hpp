// ORACLE — TEST INFRASTRUCTURE ONLY (see ff.hpp header; PARITY UNPINNED vs Rust).
// Restates halo2curves 0.3.1 src/bn256/curve.rs + src/derive/curve.rs:
// G1Affine {x,y} (identity = (0,0)), G1 Jacobian {x,y,z} (identity z = 0),
// y^2 = x^3 + 3, generator (1,2); compressed form = x LE | sign(y)<<7 in byte 31.
// Reference call sites: G1Affine is named at /root/reference/src/circuits/utils.rs:2,43,45.
#pragma once
#include "ff.hpp"

namespace orc {

struct G1Affine {
    Fq x, y;
    static G1Affine identity() { return {Fq::zero(), Fq::zero()}; }
    bool is_identity() const { return x.is_zero() && y.is_zero(); }
    bool operator==(const G1Affine& o) const { return x == o.x && y == o.y; }
    bool on_curve() const {
        if (is_identity()) return true;
        return y.sqr() == x.sqr() * x + Fq::from_u64(3);
    }
    G1Affine neg() const { return {x, y.neg()}; }
    void to_bytes(uint8_t out[32]) const {
        if (is_identity()) { memset(out, 0, 32); return; }
        uint64_t xr[4], yr[4]; x.to_raw(xr); y.to_raw(yr);
        memcpy(out, xr, 32);
        out[31] |= (uint8_t)((yr[0] & 1) << 7);
    }
};

struct G1 {
    Fq x, y, z;
    static G1 identity() { return {Fq::zero(), Fq::one(), Fq::zero()}; }
    static G1 from_affine(const G1Affine& a) {
        if (a.is_identity()) return identity();
        return {a.x, a.y, Fq::one()};
    }
    static G1 generator() { return {Fq::from_u64(1), Fq::from_u64(2), Fq::one()}; }
    bool is_identity() const { return z.is_zero(); }

    G1 dbl() const {            // dbl-2009-l (a = 0)
        if (is_identity()) return *this;
        Fq a = x.sqr(), b = y.sqr(), c = b.sqr();
        Fq d = ((x + b).sqr() - a - c).dbl();
        Fq e = a.dbl() + a, f = e.sqr();
        Fq z3 = (z * y).dbl();
        Fq x3 = f - d.dbl();
        Fq y3 = e * (d - x3) - c.dbl().dbl().dbl();
        return {x3, y3, z3};
    }
    G1 add(const G1& o) const { // add-2007-bl
        if (is_identity()) return o;
        if (o.is_identity()) return *this;
        Fq z1z1 = z.sqr(), z2z2 = o.z.sqr();
        Fq u1 = x * z2z2, u2 = o.x * z1z1;
        Fq s1 = y * z2z2 * o.z, s2 = o.y * z1z1 * z;
        if (u1 == u2) { if (s1 == s2) return dbl(); return identity(); }
        Fq h = u2 - u1, i = h.dbl().sqr(), j = h * i, r = (s2 - s1).dbl(), v = u1 * i;
        Fq x3 = r.sqr() - j - v.dbl();
        Fq y3 = r * (v - x3) - (s1 * j).dbl();
        Fq z3 = ((z + o.z).sqr() - z1z1 - z2z2) * h;
        return {x3, y3, z3};
    }
    G1 add_affine(const G1Affine& o) const {   // madd-2007-bl
        if (o.is_identity()) return *this;
        if (is_identity()) return from_affine(o);
        Fq z1z1 = z.sqr(), u2 = o.x * z1z1, s2 = o.y * z1z1 * z;
        if (x == u2) { if (y == s2) return dbl(); return identity(); }
        Fq h = u2 - x, hh = h.sqr(), i = hh.dbl().dbl(), j = h * i, r = (s2 - y).dbl(), v = x * i;
        Fq x3 = r.sqr() - j - v.dbl();
        Fq y3 = r * (v - x3) - (y * j).dbl();
        Fq z3 = (z + h).sqr() - z1z1 - hh;
        return {x3, y3, z3};
    }
    G1 neg() const { return {x, y.neg(), z}; }
    G1Affine to_affine() const {
        if (is_identity()) return G1Affine::identity();
        Fq zi = z.inv(), zi2 = zi.sqr();
        return {x * zi2, y * zi2 * zi};
    }
    // scalar given as canonical (non-Montgomery) 256-bit integer
    G1 mul_raw(const uint64_t e[4]) const {
        G1 acc = identity();
        for (int i = 255; i >= 0; --i) {
            acc = acc.dbl();
            if ((e[i >> 6] >> (i & 63)) & 1) acc = acc.add(*this);
        }
        return acc;
    }
    G1 mul(const Fr& s) const { uint64_t e[4]; s.to_raw(e); return mul_raw(e); }
};

// Curve::batch_normalize
static inline void batch_normalize(const G1* in, G1Affine* out, size_t n) {
    std::vector<Fq> zs(n);
    for (size_t i = 0; i < n; ++i) zs[i] = in[i].z;
    batch_invert(zs.data(), n);
    for (size_t i = 0; i < n; ++i) {
        if (in[i].is_identity()) { out[i] = G1Affine::identity(); continue; }
        Fq zi2 = zs[i].sqr();
        out[i] = {in[i].x * zi2, in[i].y * zi2 * zs[i]};
    }
}

}  // namespace orc
