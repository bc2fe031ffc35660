// Host-side Fiat-Shamir transcript: halo2_proofs v2023_02_02 src/transcript.rs
// `Blake2bWrite<Vec<u8>, G1Affine, Challenge255<G1Affine>>`, the type `full_prover` instantiates at
// /root/reference/src/circuits/utils.rs:39-45.  BLAKE2b-512 (RFC 7693) with the 16-byte
// personalisation "Halo2-Transcript".
//   common_point : 0x01 || x (32 B LE canonical) || y (32 B LE canonical)
//   common_scalar: 0x02 || scalar (32 B LE canonical)
//   write_*      : common_* + append compressed point / canonical scalar to the proof
//   squeeze      : absorb 0x00, finalize a CLONE of the state, 64-byte digest -> from_bytes_wide
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include "host_field.hpp"

namespace b200zk {
namespace host {

class Blake2b {
public:
    explicit Blake2b(const char personal[16]) {
        static const uint64_t IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                       0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
        uint8_t param[64];
        memset(param, 0, 64);
        param[0] = 64;  // digest length
        param[2] = 1;   // fanout
        param[3] = 1;   // depth
        memcpy(param + 48, personal, 16);
        for (int i = 0; i < 8; ++i) { uint64_t w; memcpy(&w, param + 8 * i, 8); h_[i] = IV[i] ^ w; }
        t_ = 0; buflen_ = 0;
    }
    void update(const void* data, size_t len) {
        const uint8_t* p = (const uint8_t*)data;
        while (len) {
            if (buflen_ == 128) { t_ += 128; compress(buf_, false); buflen_ = 0; }
            size_t take = 128 - buflen_ < len ? 128 - buflen_ : len;
            memcpy(buf_ + buflen_, p, take);
            buflen_ += take; p += take; len -= take;
        }
    }
    // digest of the current state without disturbing it (upstream clones the hasher)
    void finalize_clone(uint8_t out[64]) const {
        Blake2b c = *this;
        c.t_ += c.buflen_;
        memset(c.buf_ + c.buflen_, 0, 128 - c.buflen_);
        c.compress(c.buf_, true);
        memcpy(out, c.h_, 64);
    }

private:
    static uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
    void compress(const uint8_t block[128], bool last) {
        static const uint64_t IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                       0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
        static const uint8_t S[12][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
        uint64_t m[16], v[16];
        memcpy(m, block, 128);
        for (int i = 0; i < 8; ++i) { v[i] = h_[i]; v[i + 8] = IV[i]; }
        v[12] ^= t_;                    // low 64 bits of the byte counter (messages here are far below 2^64 bytes)
        if (last) v[14] = ~v[14];
        auto G = [&](int a, int b, int c, int d, uint64_t x, uint64_t y) {
            v[a] = v[a] + v[b] + x; v[d] = rotr(v[d] ^ v[a], 32);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 24);
            v[a] = v[a] + v[b] + y; v[d] = rotr(v[d] ^ v[a], 16);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 63);
        };
        for (int r = 0; r < 12; ++r) {
            const uint8_t* s = S[r];
            G(0, 4, 8, 12, m[s[0]], m[s[1]]);   G(1, 5, 9, 13, m[s[2]], m[s[3]]);
            G(2, 6, 10, 14, m[s[4]], m[s[5]]);  G(3, 7, 11, 15, m[s[6]], m[s[7]]);
            G(0, 5, 10, 15, m[s[8]], m[s[9]]);  G(1, 6, 11, 12, m[s[10]], m[s[11]]);
            G(2, 7, 8, 13, m[s[12]], m[s[13]]); G(3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
        for (int i = 0; i < 8; ++i) h_[i] ^= v[i] ^ v[i + 8];
    }
    uint64_t h_[8];
    uint64_t t_;
    uint8_t buf_[128];
    size_t buflen_;
};

class Transcript {
public:
    Transcript() : hash_("Halo2-Transcript") {}
    void common_scalar(const HFr& s) {
        uint8_t b[33]; b[0] = 0x02;
        uint64_t c[4]; s.to_canonical(c); memcpy(b + 1, c, 32);
        hash_.update(b, 33);
    }
    void common_point(const HAffine& p) {
        uint8_t b[65]; b[0] = 0x01;
        uint64_t c[4];
        p.x.to_canonical(c); memcpy(b + 1, c, 32);
        p.y.to_canonical(c); memcpy(b + 33, c, 32);
        hash_.update(b, 65);
    }
    void write_scalar(const HFr& s) {
        common_scalar(s);
        uint64_t c[4]; s.to_canonical(c);
        const uint8_t* p = (const uint8_t*)c;
        proof_.insert(proof_.end(), p, p + 32);
    }
    // G1Affine::to_bytes: x LE with the parity of y in bit 7 of byte 31; identity = all zero
    void write_point(const HAffine& pt) {
        common_point(pt);
        uint8_t out[32];
        if (pt.x.is_zero() && pt.y.is_zero()) memset(out, 0, 32);
        else {
            uint64_t xc[4], yc[4]; pt.x.to_canonical(xc); pt.y.to_canonical(yc);
            memcpy(out, xc, 32);
            out[31] |= (uint8_t)((yc[0] & 1) << 7);
        }
        proof_.insert(proof_.end(), out, out + 32);
    }
    HFr squeeze_challenge() {
        uint8_t z = 0x00;
        hash_.update(&z, 1);
        uint8_t d[64];
        hash_.finalize_clone(d);
        uint64_t w[8]; memcpy(w, d, 64);
        return HFr::from_u512(w);
    }
    const std::vector<uint8_t>& proof() const { return proof_; }

private:
    Blake2b hash_;
    std::vector<uint8_t> proof_;
};

}  // namespace host
}  // namespace b200zk
