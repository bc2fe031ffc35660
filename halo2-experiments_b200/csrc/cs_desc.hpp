// Constraint-system description shared by the prover (prover.cu) and the verifier (verifier.cu):
// the parts of halo2_proofs v2023_02_02 plonk::ConstraintSystem that create_proof / verify_proof read
// (src/plonk/circuit.rs), parsed from the blob documented in halo2-experiments_b200/circuit.py.
#pragma once
#include <algorithm>
#include <map>
#include <vector>
#include "expr.cuh"
#include "host_field.hpp"

namespace b200zk {
using host::HFr;

// ------------------------------------------------------------------ constraint system
struct CsDesc {
    uint32_t k = 0, A = 0, F = 0, I = 0, bf = 0, degree = 0;
    std::vector<int32_t> adv_q, fix_q, inst_q;                 // (col, rot) flattened
    std::vector<std::pair<uint32_t, uint32_t>> perm;            // (type, index)
    std::vector<std::pair<uint32_t, uint32_t>> gates;           // (off, len) into prog
    struct Lk { std::vector<std::pair<uint32_t, uint32_t>> ins, tabs; };
    std::vector<Lk> lookups;
    std::vector<HFr> consts;                                    // Montgomery
    std::vector<uint32_t> prog;
};

inline uint32_t expr_degree(const CsDesc& cs, uint32_t off, uint32_t len) {
    std::vector<uint32_t> st;
    for (uint32_t i = off; i < off + len; ++i) {
        uint32_t op = cs.prog[i] & 0xff;
        if (op == EX_CONST) st.push_back(0);
        else if (op == EX_FIXED || op == EX_ADVICE || op == EX_INSTANCE) st.push_back(1);
        else if (op == EX_ADD) { uint32_t b = st.back(); st.pop_back(); st.back() = std::max(st.back(), b); }
        else if (op == EX_MUL) { uint32_t b = st.back(); st.pop_back(); st.back() += b; }
    }
    return st.empty() ? 0 : st.back();
}

// circuit.rs: blinding_factors() and degree()
inline void derive_cs(CsDesc& cs) {
    std::map<int32_t, uint32_t> per_col;
    for (size_t i = 0; i < cs.adv_q.size(); i += 2) per_col[cs.adv_q[i]]++;
    uint32_t factors = 0;
    for (auto& kv : per_col) factors = std::max(factors, kv.second);
    cs.bf = std::max(3u, factors) + 2;
    uint32_t degree = 3;
    for (auto& lk : cs.lookups) {
        uint32_t di = 1, dt = 1;
        for (auto& e : lk.ins) di = std::max(di, expr_degree(cs, e.first, e.second));
        for (auto& e : lk.tabs) dt = std::max(dt, expr_degree(cs, e.first, e.second));
        degree = std::max(degree, std::max(4u, 2 + di + dt));
    }
    for (auto& g : cs.gates) degree = std::max(degree, expr_degree(cs, g.first, g.second));
    cs.degree = degree;
}

inline bool parse_cs(const uint32_t* w, size_t nw, CsDesc& cs) {
    if (nw < 16 || w[0] != 0x324B5A42u || w[1] != 1) return false;
    cs.k = w[2]; cs.A = w[3]; cs.F = w[4]; cs.I = w[5];
    uint32_t naq = w[6], nfq = w[7], niq = w[8], ng = w[9], nl = w[10], np = w[11], nc = w[12], nprog = w[13];
    size_t p = 16;
    auto need = [&](size_t c) { return p + c <= nw; };
    auto rd_q = [&](std::vector<int32_t>& q, uint32_t cnt) {
        if (!need(2 * (size_t)cnt)) return false;
        for (uint32_t i = 0; i < 2 * cnt; ++i) q.push_back((int32_t)w[p++]);
        return true;
    };
    if (!rd_q(cs.adv_q, naq) || !rd_q(cs.fix_q, nfq) || !rd_q(cs.inst_q, niq)) return false;
    if (!need(2 * (size_t)np)) return false;
    for (uint32_t i = 0; i < np; ++i) { cs.perm.push_back({w[p], w[p + 1]}); p += 2; }
    if (!need(2 * (size_t)ng)) return false;
    for (uint32_t i = 0; i < ng; ++i) { cs.gates.push_back({w[p], w[p + 1]}); p += 2; }
    for (uint32_t i = 0; i < nl; ++i) {
        if (!need(1)) return false;
        uint32_t m = w[p++];
        if (!need(4 * (size_t)m)) return false;
        CsDesc::Lk lk;
        for (uint32_t j = 0; j < m; ++j) { lk.ins.push_back({w[p], w[p + 1]}); p += 2; }
        for (uint32_t j = 0; j < m; ++j) { lk.tabs.push_back({w[p], w[p + 1]}); p += 2; }
        cs.lookups.push_back(lk);
    }
    if (!need(8 * (size_t)nc + nprog)) return false;
    for (uint32_t i = 0; i < nc; ++i) {
        uint64_t c[4];
        for (int j = 0; j < 4; ++j) c[j] = (uint64_t)w[p + 2 * j] | ((uint64_t)w[p + 2 * j + 1] << 32);
        cs.consts.push_back(HFr::from_canonical(c));
        p += 8;
    }
    cs.prog.assign(w + p, w + p + nprog);
    // validate indices
    for (size_t i = 0; i < cs.adv_q.size(); i += 2) if ((uint32_t)cs.adv_q[i] >= cs.A) return false;
    for (size_t i = 0; i < cs.fix_q.size(); i += 2) if ((uint32_t)cs.fix_q[i] >= cs.F) return false;
    for (size_t i = 0; i < cs.inst_q.size(); i += 2) if ((uint32_t)cs.inst_q[i] >= cs.I) return false;
    for (uint32_t word : cs.prog) {
        uint32_t op = word & 0xff, arg = word >> 8;
        if (op > EX_SCALE) return false;
        if ((op == EX_CONST || op == EX_SCALE) && arg >= nc) return false;
        if (op == EX_FIXED && arg >= nfq) return false;
        if (op == EX_ADVICE && arg >= naq) return false;
        if (op == EX_INSTANCE && arg >= niq) return false;
    }
    derive_cs(cs);
    if (cs.bf != w[14] || cs.degree != w[15]) return false;      // frontend and library must agree
    return true;
}

}  // namespace b200zk
