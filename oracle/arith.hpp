// ORACLE — TEST INFRASTRUCTURE ONLY (see ff.hpp header; PARITY UNPINNED vs Rust).
// Restates halo2_proofs (PSE tag v2023_02_02) src/arithmetic.rs and src/poly/domain.rs.
// The reference reaches these through keygen_vk / keygen_pk / create_proof at
// /root/reference/src/circuits/utils.rs:31,35,40-48.
#pragma once
#include "curve.hpp"
#include <functional>
#include <thread>

namespace orc {

// multicore::current_num_threads(); overridable for tests / the bench
int num_threads();
void set_num_threads(int t);

// arithmetic::parallelize — chunk = n / threads (upstream splits into equal chunks)
void parallelize(size_t n, const std::function<void(size_t, size_t)>& f);

// arithmetic::best_multiexp / multiexp_serial
void multiexp_serial(const Fr* coeffs, const G1Affine* bases, size_t n, G1& acc);
G1 best_multiexp(const Fr* coeffs, const G1Affine* bases, size_t n);

// arithmetic::best_fft / recursive_butterfly_arithmetic  (G = Fr)
void best_fft(Fr* a, const Fr& omega, unsigned log_n);

// arithmetic::eval_polynomial (Horner), kate_division
Fr eval_polynomial(const Fr* poly, size_t n, const Fr& x);
void kate_division(const Fr* a, size_t n, const Fr& b, Fr* q /* n-1 */);

// poly::EvaluationDomain<Fr>
struct Domain {
    unsigned k, extended_k;
    uint64_t n, quotient_poly_degree;
    Fr omega, omega_inv, extended_omega, extended_omega_inv;
    Fr g_coset, g_coset_inv, ifft_divisor, extended_ifft_divisor, barycentric_weight;
    std::vector<Fr> t_evaluations;   // already inverted

    Domain(unsigned j, unsigned k);
    size_t extended_len() const { return (size_t)1 << extended_k; }
    void lagrange_to_coeff(Fr* a) const;                         // in place, n
    void coeff_to_extended(const Fr* a, Fr* out) const;          // n -> ext
    void extended_to_coeff(Fr* a) const;                         // in place, ext (first n*q meaningful)
    void divide_by_vanishing_poly(Fr* a) const;                  // in place, ext
    Fr rotate_omega(const Fr& x, int rot) const;
    void distribute_powers_zeta(Fr* a, size_t len, bool into_coset) const;
};

}  // namespace orc
