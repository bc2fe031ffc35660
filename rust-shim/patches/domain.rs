// halo2_proofs/src/poly/domain.rs (tag v2023_02_02) — EvaluationDomain<Fr> gains a device handle created in `new`
// (b200zk_domain_create(j, k) recomputes omega, extended_omega, t_evaluations ...; b200zk_domain_constant exposes them
// for a debug assertion against the Rust-side fields), and the four O(n log n) methods call it.
use b200zk_shim::{check, ctx, sys};

// in EvaluationDomain::new, after the fields are computed:
//     let mut dev = std::ptr::null_mut();
//     check(unsafe { sys::b200zk_domain_create(ctx(), j, k, &mut dev) });

pub fn lagrange_to_coeff(&self, mut a: Polynomial<G, LagrangeCoeff>) -> Polynomial<G, Coeff> {
    assert_eq!(a.values.len(), 1 << self.k);
    check(unsafe { sys::b200zk_lagrange_to_coeff(self.dev, a.values.as_mut_ptr() as _) });
    Polynomial { values: a.values, _marker: PhantomData }
}

pub fn coeff_to_extended(&self, a: Polynomial<G, Coeff>) -> Polynomial<G, ExtendedLagrangeCoeff> {
    assert_eq!(a.values.len(), 1 << self.k);
    let mut out = vec![G::group_zero(); self.extended_len()];
    check(unsafe { sys::b200zk_coeff_to_extended(self.dev, a.values.as_ptr() as _, out.as_mut_ptr() as _) });
    Polynomial { values: out, _marker: PhantomData }
}

pub fn extended_to_coeff(&self, a: Polynomial<G, ExtendedLagrangeCoeff>) -> Vec<G> {
    assert_eq!(a.values.len(), self.extended_len());
    let mut out = vec![G::group_zero(); (self.n * self.quotient_poly_degree) as usize];
    check(unsafe { sys::b200zk_extended_to_coeff(self.dev, a.values.as_ptr() as _, out.as_mut_ptr() as _) });
    out
}

pub fn divide_by_vanishing_poly(&self, mut a: Polynomial<G, ExtendedLagrangeCoeff>) -> Polynomial<G, ExtendedLagrangeCoeff> {
    assert_eq!(a.values.len(), self.extended_len());
    check(unsafe { sys::b200zk_divide_by_vanishing_poly(self.dev, a.values.as_mut_ptr() as _) });
    Polynomial { values: a.values, _marker: PhantomData }
}
// rotate_omega / l_i_range stay on the CPU (O(range) scalar work); b200zk_domain_rotate_omega / _l_i_range exist for
// hosts without a field implementation.
