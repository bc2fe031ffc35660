// bn256 Fr / Fq on the sm_100a integer pipe: 8 x u32 little-endian limbs in
// Montgomery form (R = 2^256) — bit-identical in memory to halo2curves'
// `Fr([u64; 4])` / `Fq([u64; 4])` (halo2curves 0.3.1 src/bn256/{fr,fq}.rs, reached
// from /root/reference/src/circuits/utils.rs:2), so Rust `Vec<Fr>` buffers are
// consumed zero-copy.  All results are fully reduced to [0, p): every value that
// leaves a kernel is the canonical Montgomery representative, which is what
// makes device results bit-exact against the CPU path.
//
// The multiplier is an even/odd-column CIOS: products a[j]*b_i with even j land
// on 64-bit aligned limb pairs of one accumulator, odd j on the other, so each
// (lo,hi) pair is one IMAD.WIDE.U32 with carry-in/out and the two carry chains
// never collide.  8 rounds x (16 + 16 wide halves + 1 mul.lo) = 264 mul32 halves
// = 132 IMAD.WIDE + 8 IMAD per field multiplication.
#pragma once
#include "ptx_chain.cuh"

namespace b200zk {

struct alignas(32) fe_t { uint32_t l[8]; };   // 32-byte aligned -> LDG.E.256 / STG.E.256 on sm_100a

struct FrCfg {
    static constexpr uint32_t INV = 0xefffffffu;         // -r^-1 mod 2^32
    ZK_D static constexpr uint32_t p(int i) {
        constexpr uint32_t t[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return t[i];
    }
    ZK_D static constexpr uint32_t one(int i) {          // R mod r
        constexpr uint32_t t[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                                   0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return t[i];
    }
    ZK_D static constexpr uint32_t r2(int i) {           // R^2 mod r
        constexpr uint32_t t[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                                   0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return t[i];
    }
    ZK_D static constexpr uint32_t r3(int i) {           // R^3 mod r (from_u512 / Fr::random)
        constexpr uint32_t t[8] = {0xb4bf0040u, 0x5e94d8e1u, 0x1cfbb6b8u, 0x2a489cbeu,
                                   0xa19fcfedu, 0x893cc664u, 0x7fcc657cu, 0x0cf8594bu};
        return t[i];
    }
};

struct FqCfg {
    static constexpr uint32_t INV = 0xe4866389u;         // -q^-1 mod 2^32
    ZK_D static constexpr uint32_t p(int i) {
        constexpr uint32_t t[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return t[i];
    }
    ZK_D static constexpr uint32_t one(int i) {          // R mod q
        constexpr uint32_t t[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                                   0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return t[i];
    }
    ZK_D static constexpr uint32_t r2(int i) {           // R^2 mod q
        constexpr uint32_t t[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                                   0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return t[i];
    }
};

template <class C> struct Field {
    typedef fe_t T;

    ZK_D static T zero() { T r; for (int i = 0; i < 8; ++i) r.l[i] = 0; return r; }
    ZK_D static T one() { T r; for (int i = 0; i < 8; ++i) r.l[i] = C::one(i); return r; }
    ZK_D static T r2() { T r; for (int i = 0; i < 8; ++i) r.l[i] = C::r2(i); return r; }
    ZK_D static bool is_zero(const T& a) {
        uint32_t o = 0; for (int i = 0; i < 8; ++i) o |= a.l[i]; return o == 0;
    }
    ZK_D static bool eq(const T& a, const T& b) {
        uint32_t o = 0; for (int i = 0; i < 8; ++i) o |= a.l[i] ^ b.l[i]; return o == 0;
    }

    // r = a - p if a >= p else a          (a < 2p)
    ZK_D static void reduce_once(T& a) {
        uint32_t t[8];
        t[0] = ptx::sub_cc(a.l[0], C::p(0));
#pragma unroll
        for (int i = 1; i < 8; ++i) t[i] = ptx::subc_cc(a.l[i], C::p(i));
        uint32_t borrow = ptx::subc(0, 0);               // 0xffffffff if a < p
#pragma unroll
        for (int i = 0; i < 8; ++i) a.l[i] = borrow ? a.l[i] : t[i];
    }

    ZK_D static T add(const T& a, const T& b) {
        T r;
        r.l[0] = ptx::add_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < 7; ++i) r.l[i] = ptx::addc_cc(a.l[i], b.l[i]);
        r.l[7] = ptx::addc(a.l[7], b.l[7]);              // p < 2^254: no carry out
        reduce_once(r);
        return r;
    }
    ZK_D static T sub(const T& a, const T& b) {
        T r;
        r.l[0] = ptx::sub_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < 8; ++i) r.l[i] = ptx::subc_cc(a.l[i], b.l[i]);
        uint32_t borrow = ptx::subc(0, 0);               // all-ones if a < b
        r.l[0] = ptx::add_cc(r.l[0], C::p(0) & borrow);
#pragma unroll
        for (int i = 1; i < 7; ++i) r.l[i] = ptx::addc_cc(r.l[i], C::p(i) & borrow);
        r.l[7] = ptx::addc(r.l[7], C::p(7) & borrow);
        return r;
    }
    ZK_D static T neg(const T& a) { return sub(zero(), a); }
    ZK_D static T dbl(const T& a) { return add(a, a); }

    // acc[0..7] += a[even limbs 0,2,4,6 starting at `a`] * b as one carry chain; the
    // carry out of acc[7] is left in CC.
    ZK_D static void cmad_row(uint32_t* acc, const uint32_t* a, uint32_t b) {
        acc[0] = ptx::mad_lo_cc(a[0], b, acc[0]);
        acc[1] = ptx::madc_hi_cc(a[0], b, acc[1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            acc[j] = ptx::madc_lo_cc(a[j], b, acc[j]);
            acc[j + 1] = ptx::madc_hi_cc(a[j], b, acc[j + 1]);
        }
    }
    // same, consuming the incoming CC carry and reading the addend two limbs up:
    // acc[j] = acc[j+2] + a[j]*b  (j < 6),  acc[6..7] = a[6]*b  — the "shift right by 64 bits"
    // that re-aligns the accumulator which just lost its low limb.
    ZK_D static void madc_row_rshift(uint32_t* acc, const uint32_t* a, uint32_t b) {
#pragma unroll
        for (int j = 0; j < 6; j += 2) {
            acc[j] = ptx::madc_lo_cc(a[j], b, acc[j + 2]);
            acc[j + 1] = ptx::madc_hi_cc(a[j], b, acc[j + 3]);
        }
        acc[6] = ptx::madc_lo_cc(a[6], b, 0);
        acc[7] = ptx::madc_hi(a[6], b, 0);
    }
    ZK_D static void mul_row(uint32_t* acc, const uint32_t* a, uint32_t b) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) { acc[j] = ptx::mul_lo(a[j], b); acc[j + 1] = ptx::mul_hi(a[j], b); }
    }

    // One CIOS round.  `ev` holds limb positions 0..7, `od` positions 1..8 on exit;
    // on entry (not first) `od` is the previous round's even accumulator, i.e. it is
    // positioned at -1..6 with od[0] == 0 already consumed.
    ZK_D static void round(uint32_t* ev, uint32_t* od, const uint32_t* a, uint32_t bi, bool first) {
        const uint32_t P[8] = {C::p(0), C::p(1), C::p(2), C::p(3), C::p(4), C::p(5), C::p(6), C::p(7)};
        if (first) {
            mul_row(od, a + 1, bi);
            mul_row(ev, a, bi);
        } else {
            ev[0] = ptx::add_cc(ev[0], od[1]);           // position 0 of the shifted accumulator
            madc_row_rshift(od, a + 1, bi);              // carry continues into position 1
            cmad_row(ev, a, bi);
            od[7] = ptx::addc(od[7], 0);                 // carry out of position 7 -> position 8
        }
        uint32_t m = ptx::mul_lo(ev[0], C::INV);
        cmad_row(od, P + 1, m);                          // total < 2^288, so no carry out of od[7]
        cmad_row(ev, P, m);
        od[7] = ptx::addc(od[7], 0);
        // now ev[0] == 0; dividing by 2^32 swaps the roles of the two accumulators
    }

    ZK_D static T mul(const T& a, const T& b) {
        uint32_t ev[8], od[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            round(ev, od, a.l, b.l[i], i == 0);
            round(od, ev, a.l, b.l[i + 1], false);
        }
        // result = ev (positions 0..7) + od>>32 (od[1..7] at positions 0..6)
        T r;
        r.l[0] = ptx::add_cc(ev[0], od[1]);
#pragma unroll
        for (int i = 1; i < 7; ++i) r.l[i] = ptx::addc_cc(ev[i], od[i + 1]);
        r.l[7] = ptx::addc(ev[7], 0);
        reduce_once(r);
        return r;
    }
    ZK_D static T sqr(const T& a) { return mul(a, a); }

    // ---- multiplication by a constant with a precomputed quotient (Shoup / Barrett with a fixed operand) --------------
    // For a constant w < p stored with wq = floor(w * 2^256 / p):   x * w mod p  =  x*w - q*p  (mod 2^256),  q = floor(x * wq / 2^256),
    // which lies in [0, 2p); with q taken from the columns >= 6 of the 16-limb product only (the dropped low columns sum to
    // less than 2^256, so the estimate is q or q - 1) it lies in [0, 3p) < 2^256, and two conditional subtractions finish.
    // When x is in Montgomery form and w is the PLAIN value of the constant, the result is the Montgomery form of the
    // product — no domain change — so this replaces mul(x, w_mont) wherever the second operand comes from a table (every
    // multiplication of an NTT: butterfly roots and inter-pass twiddles).  Cost: 43 (high part of x*wq) + 28 (low half of
    // x*w) + 28 (low half of q*p) wide multiply-adds + 16 low-half multiplies, against 128 + 8 for the CIOS product.
    // Same even/odd accumulator scheme as `round`: products landing on even limb positions go to one accumulator, odd ones
    // to the other, each row is one carry chain per accumulator, so every (lo, hi) pair is one IMAD.WIDE.U32 with carry.
    //
    // chain: acc[off + j], acc[off + j + 1] += a[j] * b for j = j0, j0 + 2, ... < jn;  lo_last: the last product contributes
    // its low half only (its high half falls outside the kept window);  carry_out: the carry leaving the last pair is added
    // to the limb above it (which no earlier row of the same accumulator has touched).
    ZK_D static void mad_chain(uint32_t* acc, int off, const uint32_t* a, uint32_t b, int j0, int jn, bool lo_last, bool carry_out) {
        if (j0 >= jn) return;
        int jl = j0;
#pragma unroll
        for (int j = j0; j < jn; j += 2) {
            const bool first = j == j0, last = j + 2 >= jn;
            uint32_t* q = acc + off + j;
            if (last && lo_last) {
                q[0] = first ? ptx::mad_lo(a[j], b, q[0]) : ptx::madc_lo(a[j], b, q[0]);
            } else {
                q[0] = first ? ptx::mad_lo_cc(a[j], b, q[0]) : ptx::madc_lo_cc(a[j], b, q[0]);
                q[1] = (last && !carry_out) ? ptx::madc_hi(a[j], b, q[1]) : ptx::madc_hi_cc(a[j], b, q[1]);
            }
            jl = j;
        }
        if (carry_out && !lo_last) acc[off + jl + 2] = ptx::addc(acc[off + jl + 2], 0);
    }
    // low 8 limbs of a * b (a, b: 8 limbs)
    ZK_D static void mul_low(uint32_t* r, const uint32_t* a, const uint32_t* b) {
        uint32_t E[8], O[9];                                   // E: positions 0..7 ; O[k]: position k (1..8; O[0] unused)
#pragma unroll
        for (int i = 0; i < 8; ++i) E[i] = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) O[i] = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            // product (i, j) sits at position i + j <= 7;  even position -> E, odd -> O
            const int je = i & 1, jo = (i & 1) ^ 1;            // first j of each parity class
            mad_chain(E, i, a, b[i], je, 8 - i, false, false);              // ends at pair (6, 7): carry out of the window dropped
            mad_chain(O, i, a, b[i], jo, 8 - i, true, false);               // ends with the low half at position 7
        }
        r[0] = E[0];
        r[1] = ptx::add_cc(E[1], O[1]);
#pragma unroll
        for (int k = 2; k < 7; ++k) r[k] = ptx::addc_cc(E[k], O[k]);
        r[7] = ptx::addc(E[7], O[7]);
    }
    // limbs 8..15 of the part of x * b made of the products at positions >= 6 (see above): floor(x*b / 2^256) or one less
    ZK_D static void mul_high_approx(uint32_t* r, const uint32_t* x, const uint32_t* b) {
        uint32_t E[12], O[12];                                 // E[k]: position 6 + k (pairs (6,7) .. (14,15)); O[k]: position 7 + k
#pragma unroll
        for (int i = 0; i < 12; ++i) { E[i] = 0; O[i] = 0; }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int jmin = i >= 6 ? 0 : 6 - i;               // products with i + j >= 6
            const int je = ((jmin + i) & 1) ? jmin + 1 : jmin; // first j with i + j even
            const int jo = ((jmin + i) & 1) ? jmin : jmin + 1; // first j with i + j odd
            mad_chain(E, i - 6, x, b[i], je, 8, false, true);
            mad_chain(O, i - 7, x, b[i], jo, 8, false, true);
        }
        (void)ptx::add_cc(E[1], O[0]);                         // position 7: only its carry matters
        r[0] = ptx::addc_cc(E[2], O[1]);
#pragma unroll
        for (int k = 1; k < 7; ++k) r[k] = ptx::addc_cc(E[2 + k], O[1 + k]);
        r[7] = ptx::addc(E[9], O[8]);
    }
    // low 8 limbs of a * b + c * d: the rows of both products go into the same pair of accumulators
    ZK_D static void mul_low2(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d) {
        uint32_t E[8], O[9];
#pragma unroll
        for (int i = 0; i < 8; ++i) E[i] = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) O[i] = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int je = i & 1, jo = (i & 1) ^ 1;
            mad_chain(E, i, a, b[i], je, 8 - i, false, false);
            mad_chain(O, i, a, b[i], jo, 8 - i, true, false);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int je = i & 1, jo = (i & 1) ^ 1;
            mad_chain(E, i, c, d[i], je, 8 - i, false, false);
            mad_chain(O, i, c, d[i], jo, 8 - i, true, false);
        }
        r[0] = E[0];
        r[1] = ptx::add_cc(E[1], O[1]);
#pragma unroll
        for (int k = 2; k < 7; ++k) r[k] = ptx::addc_cc(E[k], O[k]);
        r[7] = ptx::addc(E[7], O[7]);
    }
    ZK_D static constexpr uint32_t neg_p(int i) {             // limb i of 2^256 - p
        return i == 0 ? 0u - C::p(0) : ~C::p(i);             // p(0) != 0: the borrow never propagates
    }
    ZK_D static constexpr uint32_t two_p(int i) {             // limb i of 2p (< 2^255)
        return i == 0 ? C::p(0) << 1 : (C::p(i) << 1) | (C::p(i - 1) >> 31);
    }
    // x * w - q * p  (mod 2^256), in [0, 3p):  the subtraction is folded into the product as + q * (2^256 - p)
    ZK_D static T mul_shoup_raw(const T& x, const T& w, const T& wq) {
        const uint32_t NP[8] = {neg_p(0), neg_p(1), neg_p(2), neg_p(3), neg_p(4), neg_p(5), neg_p(6), neg_p(7)};
        uint32_t q[8];
        mul_high_approx(q, x.l, wq.l);
        T r;
        mul_low2(r.l, x.l, w.l, NP, q);                        // (2^256 - p)'s limbs are immediates
        return r;
    }
    ZK_D static T mul_shoup(const T& x, const T& w, const T& wq) {
        T r = mul_shoup_raw(x, w, wq);
        reduce_once(r);
        reduce_once(r);
        return r;
    }
    // ---- the same with values kept in [0, 2p) between operations (the butterflies of a transform pass) ----------------
    // Any x < 2^256 is a valid first operand of mul_shoup, and 4p < 2^256, so a pass can carry its values in [0, 2p): a
    // product needs one conditional subtraction instead of two, a sum one (of 2p), a difference that goes straight into a
    // product none; the pass that writes the final result reduces once more on the way out.
    ZK_D static void reduce_2p(T& a) {                        // a - 2p if a >= 2p else a     (a < 4p)
        uint32_t t[8];
        t[0] = ptx::sub_cc(a.l[0], two_p(0));
#pragma unroll
        for (int i = 1; i < 8; ++i) t[i] = ptx::subc_cc(a.l[i], two_p(i));
        uint32_t borrow = ptx::subc(0, 0);
#pragma unroll
        for (int i = 0; i < 8; ++i) a.l[i] = borrow ? a.l[i] : t[i];
    }
    ZK_D static T mul_shoup_lazy(const T& x, const T& w, const T& wq) {   // x < 2^256 -> [0, 2p)
        T r = mul_shoup_raw(x, w, wq);
        reduce_once(r);
        return r;
    }
    ZK_D static T add_lazy(const T& a, const T& b) {         // a, b in [0, 2p) -> [0, 2p)
        T r;
        r.l[0] = ptx::add_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < 7; ++i) r.l[i] = ptx::addc_cc(a.l[i], b.l[i]);
        r.l[7] = ptx::addc(a.l[7], b.l[7]);
        reduce_2p(r);
        return r;
    }
    ZK_D static T sub_raw(const T& a, const T& b) {          // a, b in [0, 2p) -> a - b + 2p in (0, 4p)
        T r;
        r.l[0] = ptx::sub_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < 7; ++i) r.l[i] = ptx::subc_cc(a.l[i], b.l[i]);
        r.l[7] = ptx::subc(a.l[7], b.l[7]);
        r.l[0] = ptx::add_cc(r.l[0], two_p(0));
#pragma unroll
        for (int i = 1; i < 7; ++i) r.l[i] = ptx::addc_cc(r.l[i], two_p(i));
        r.l[7] = ptx::addc(r.l[7], two_p(7));
        return r;
    }

    ZK_D static T from_mont(const T& a) {                // a * 1 * R^-1 : canonical integer
        T o = zero(); o.l[0] = 1; return mul(a, o);
    }
    ZK_D static T to_mont(const T& a) { return mul(a, r2()); }

    // a^e, e given as canonical 256-bit LE limbs (variable time)
    ZK_D static T pow(const T& a, const uint32_t e[8]) {
        T acc = one();
        for (int i = 255; i >= 0; --i) {
            acc = sqr(acc);
            if ((e[i >> 5] >> (i & 31)) & 1) acc = mul(acc, a);
        }
        return acc;
    }
    ZK_D static T pow_u64(const T& a, unsigned long long e) {
        T acc = one();
        bool started = false;
        for (int i = 63; i >= 0; --i) {
            if (started) acc = sqr(acc);
            if ((e >> i) & 1) { acc = started ? mul(acc, a) : a; started = true; }
        }
        return acc;
    }
    ZK_D static T inv(const T& a) {                      // Fermat; inv(0) = 0
        uint32_t e[8];
        e[0] = C::p(0) - 2;                              // p is odd and p(0) >= 2 for both fields
#pragma unroll
        for (int i = 1; i < 8; ++i) e[i] = C::p(i);
        return pow(a, e);
    }

    // ---- inversion by division steps ---------------------------------------------------------------
    // Bernstein–Yang "safegcd" (the delta = 1/2 variant, 30 division steps per round on the low words, then one 2 x 2
    // integer matrix applied to the 256-bit state): about 20 rounds of ~1100 simple integer instructions against the 384
    // dependent field multiplications of Fermat — a sixth of the latency of a lone inversion, which is what a batch
    // inversion of a short column waits for (profiles/r02_launches_v3_k14_summary.txt: 0.23 ms per launch).
    //   state   f = p, g = x, d = 0, e = 1 with  f = d x,  g = e x  (mod p);  at g = 0, f = +-1 and x^-1 = +-d
    //   limbs   9 signed limbs of 30 bits; d, e stay inside (-2p, p)
    // Same result as inv(): the canonical Montgomery representative, inv_gcd(0) = 0.
    static constexpr int32_t M30 = (int32_t)((1u << 30) - 1u);
    ZK_D static constexpr int32_t p30(int i) {                // limb i of p in base 2^30
        return (int32_t)((i == 8 ? (C::p(7) >> 16)
                                 : ((C::p((30 * i) >> 5) >> ((30 * i) & 31)) |
                                    (((30 * i) & 31) > 2 ? (C::p(((30 * i) >> 5) + 1) << (32 - ((30 * i) & 31))) : 0u))) & (uint32_t)M30);
    }
    ZK_D static void to30(const uint32_t* w, int32_t* o) {
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int bit = 30 * i, wd = bit >> 5, sh = bit & 31;
            uint32_t v = w[wd] >> sh;
            if (sh > 2 && wd + 1 < 8) v |= w[wd + 1] << (32 - sh);
            o[i] = (int32_t)(v & (uint32_t)M30);
        }
    }
    // 30 division steps on the low words; t = {u, v, q, r} with 2^30 (f', g') = t (f, g)
    ZK_D static int32_t divsteps30(int32_t zeta, uint32_t f, uint32_t g, int32_t* t) {
        uint32_t u = 1, v = 0, q = 0, r = 1;
#pragma unroll 6
        for (int i = 0; i < 30; ++i) {
            uint32_t c1 = (uint32_t)(zeta >> 31), c2 = 0u - (g & 1u);
            const uint32_t x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;
            g += x & c2; q += y & c2; r += z & c2;
            c1 &= c2;
            zeta = (int32_t)((uint32_t)zeta ^ c1) - 1;
            f += g & c1; u += q & c1; v += r & c1;
            g >>= 1; u <<= 1; v <<= 1;
        }
        t[0] = (int32_t)u; t[1] = (int32_t)v; t[2] = (int32_t)q; t[3] = (int32_t)r;
        return zeta;
    }
    ZK_D static T inv_gcd(const T& a) {
        int32_t f[9], g[9], d[9], e[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) { f[i] = p30(i); d[i] = 0; e[i] = 0; }
        e[0] = 1;
        to30(a.l, g);
        const uint32_t pinv = (0u - C::INV) & (uint32_t)M30;      // p^-1 mod 2^30
        int32_t zeta = -1;
        for (int round = 0; round < 24; ++round) {                // 590 steps suffice for 256-bit inputs: 20 rounds
            int32_t nz = 0;
#pragma unroll
            for (int i = 0; i < 9; ++i) nz |= g[i];
            if (nz == 0) break;
            int32_t t[4];
            zeta = divsteps30(zeta, (uint32_t)f[0] | ((uint32_t)f[1] << 30), (uint32_t)g[0] | ((uint32_t)g[1] << 30), t);
            const int64_t u = t[0], v = t[1], q = t[2], r = t[3];
            {   // (d, e) <- t (d, e) / 2^30 mod p
                const int32_t sd = d[8] >> 31, se = e[8] >> 31;
                int32_t md = (t[0] & sd) + (t[1] & se), me = (t[2] & sd) + (t[3] & se);
                int64_t cd = u * d[0] + v * e[0], ce = q * d[0] + r * e[0];
                md -= (int32_t)((pinv * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
                me -= (int32_t)((pinv * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
                cd += (int64_t)p30(0) * md; ce += (int64_t)p30(0) * me;
                cd >>= 30; ce >>= 30;
#pragma unroll
                for (int i = 1; i < 9; ++i) {
                    cd += u * d[i] + v * e[i] + (int64_t)p30(i) * md;
                    ce += q * d[i] + r * e[i] + (int64_t)p30(i) * me;
                    d[i - 1] = (int32_t)cd & M30; cd >>= 30;
                    e[i - 1] = (int32_t)ce & M30; ce >>= 30;
                }
                d[8] = (int32_t)cd; e[8] = (int32_t)ce;
            }
            {   // (f, g) <- t (f, g) / 2^30
                int64_t cf = u * f[0] + v * g[0], cg = q * f[0] + r * g[0];
                cf >>= 30; cg >>= 30;
#pragma unroll
                for (int i = 1; i < 9; ++i) {
                    cf += u * f[i] + v * g[i];
                    cg += q * f[i] + r * g[i];
                    f[i - 1] = (int32_t)cf & M30; cf >>= 30;
                    g[i - 1] = (int32_t)cg & M30; cg >>= 30;
                }
                f[8] = (int32_t)cf; g[8] = (int32_t)cg;
            }
        }
        // x^-1 = sign(f) d, brought from (-2p, 2p) into [0, p)
        const int32_t sf = f[8] >> 31;
        int32_t w[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) w[i] = (d[i] ^ sf) - sf;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {                    // + p while negative (twice), then - p if >= p
            int32_t carry = 0, add[9];
            const int32_t neg = w[8] >> 31;
            if (pass == 2) {                                      // trial subtraction
                int32_t c2 = 0, tr[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) { const int32_t s = w[i] - p30(i) + c2; tr[i] = s & M30; c2 = s >> 30; if (i == 8) tr[i] = s; }
                const int32_t keep = tr[8] >> 31;                 // negative: w < p, keep w
#pragma unroll
                for (int i = 0; i < 9; ++i) w[i] = (w[i] & keep) | (tr[i] & ~keep);
                break;
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) add[i] = p30(i) & neg;
#pragma unroll
            for (int i = 0; i < 9; ++i) { const int32_t s = w[i] + add[i] + carry; carry = s >> 30; w[i] = i == 8 ? s : (s & M30); }
        }
        T y;
#pragma unroll
        for (int i = 0; i < 8; ++i) {                             // 9 x 30 -> 8 x 32
            const int bit = 32 * i, lb = bit / 30, sh = bit % 30;
            uint32_t v = (uint32_t)w[lb] >> sh;
            v |= (uint32_t)w[lb + 1] << (30 - sh);              // sh <= 14: two limbs always cover the word
            y.l[i] = v;
        }
        // x was a R (Montgomery form): y = a^-1 R^-1, and y * R^3 * R^-1 = a^-1 R
        return mul(y, mul(r2(), r2()));
    }
};

typedef Field<FrCfg> Fr;
typedef Field<FqCfg> Fq;

}  // namespace b200zk
