//! Polynomial (reference univariate/*.rs): coefficient vector low -> high.  mul / eval_domain / interpolate_domain /
//! scale / zerofier forward to the GPU; add / sub / neg / deg stay on the host (O(n), not on the hot path).
#![allow(dead_code)]
use crate::ff::{FieldElement, FiniteField};
use crate::ffi;

#[derive(Debug, Clone)]
pub struct Polynomial { pub coeffs: Vec<FieldElement>, pub field: FiniteField }

fn raw(v: &[FieldElement]) -> Vec<u64> { v.iter().map(|e| e.value).collect() }

impl PartialEq for Polynomial {
    fn eq(&self, o: &Self) -> bool {
        let d = self.deg();
        d == o.deg() && (d < 0 || (0..=d as usize).all(|i| self.coeffs[i].value == o.coeffs[i].value))
    }
}

impl Polynomial {
    pub fn new(coeffs: Vec<FieldElement>, field: FiniteField) -> Polynomial { Polynomial { coeffs, field } }
    fn from_raw(v: Vec<u64>, field: FiniteField) -> Polynomial { Polynomial { coeffs: v.into_iter().map(|x| field.new_element(x)).collect(), field } }
    pub fn deg(&self) -> i128 { self.coeffs.iter().rposition(|c| c.value != 0).map_or(-1, |i| i as i128) }
    pub fn is_zero(&self) -> bool { self.deg() == -1 }
    pub fn leading_coeff(&self) -> FieldElement {
        if self.is_zero() { panic!("Zero polynomial has no leading coefficient"); }
        self.coeffs[self.deg() as usize]
    }
    pub fn zero_poly(field: &FiniteField) -> Polynomial { Polynomial::new(vec![], *field) }
    pub fn constant_poly(field: &FiniteField, v: u64) -> Polynomial { Polynomial::new(vec![field.new_element(v)], *field) }
    pub fn linear_poly(field: &FiniteField, a: u64, b: u64) -> Polynomial { Polynomial::new(vec![field.new_element(a), field.new_element(b)], *field) }
    pub fn neg(p: &Polynomial) -> Polynomial { Polynomial::new(p.coeffs.iter().map(|&c| -c).collect(), p.field) }
    fn zip_with(l: &Polynomial, r: &Polynomial, f: impl Fn(FieldElement, FieldElement) -> FieldElement) -> Polynomial {
        let z = l.field.zero();
        let n = l.coeffs.len().max(r.coeffs.len());
        Polynomial::new((0..n).map(|i| f(*l.coeffs.get(i).unwrap_or(&z), *r.coeffs.get(i).unwrap_or(&z))).collect(), l.field)
    }
    pub fn add(l: &Polynomial, r: &Polynomial) -> Polynomial {
        if l.is_zero() { return r.clone(); }
        if r.is_zero() { return l.clone(); }
        Self::zip_with(l, r, |a, b| a + b)
    }
    pub fn sub(l: &Polynomial, r: &Polynomial) -> Polynomial {
        if l.is_zero() { return Self::neg(r); }
        if r.is_zero() { return l.clone(); }
        Self::zip_with(l, r, |a, b| a - b)
    }
    /// mul.rs:6-29 -> stark_poly_mul ([] if either side is zero, else len_l + len_r - 1)
    pub fn mul(l: &Polynomial, r: &Polynomial) -> Polynomial {
        let (a, b) = (raw(&l.coeffs), raw(&r.coeffs));
        let (mut out, mut n) = (vec![0u64; (a.len() + b.len()).max(1)], 0usize);
        ffi::check(unsafe { ffi::stark_poly_mul(ffi::ctx(), a.as_ptr(), a.len(), b.as_ptr(), b.len(), out.as_mut_ptr(), &mut n) });
        out.truncate(n);
        Self::from_raw(out, l.field)
    }
    /// exp.rs:6-33: square and multiply over Polynomial::mul (every product is one stark_poly_mul on the device); the same
    /// shape rules: exp 0 -> [1], zero base -> []
    pub fn exp(base: &Polynomial, exp: u64) -> Polynomial {
        if exp == 0 { return Polynomial::new(vec![base.field.one()], base.field); }
        if base.is_zero() { return Polynomial::new(vec![], base.field); }
        let (mut e, mut result, mut bpower) = (exp, Polynomial::new(vec![base.field.one()], base.field), base.clone());
        while e != 0 {
            if e & 1 == 1 { result = Self::mul(&result, &bpower); }
            bpower = Self::mul(&bpower, &bpower);
            e >>= 1;
        }
        result
    }
    pub fn eval(&self, x: &FieldElement) -> FieldElement { self.coeffs.iter().rev().fold(x.field.zero(), |acc, c| acc * *x + *c) }
    /// (offset, log_n) if `domain` is offset * w_N^i in natural order
    fn as_coset(field: &FiniteField, domain: &[FieldElement]) -> Option<(u64, u32)> {
        let n = domain.len();
        if !n.is_power_of_two() || n > (1 << 23) || domain[0].value == 0 { return None; }
        let w = field.prim_nth_root(n as u64);
        let mut x = domain[0];
        for d in domain { if d.value != x.value { return None; } x = x * w; }
        Some((domain[0].value, n.trailing_zeros()))
    }
    /// eval.rs:16-21
    pub fn eval_domain(&self, domain: &Vec<FieldElement>) -> Vec<FieldElement> {
        let c = raw(&self.coeffs);
        let mut out = vec![0u64; domain.len()];
        match Self::as_coset(&self.field, domain) {
            Some((off, lg)) if c.len() <= domain.len() =>
                ffi::check(unsafe { ffi::stark_poly_eval_coset(ffi::ctx(), c.as_ptr(), c.len(), off, lg, out.as_mut_ptr()) }),
            _ => { let d = raw(domain);
                   ffi::check(unsafe { ffi::stark_poly_eval_domain(ffi::ctx(), c.as_ptr(), c.len(), d.as_ptr(), d.len(), out.as_mut_ptr()) }) }
        }
        out.into_iter().map(|v| self.field.new_element(v)).collect()
    }
    /// interpolate.rs:6-44 (coeffs.len() rule of the reference kept by the back end)
    pub fn interpolate_domain(domain: &Vec<FieldElement>, values: &Vec<FieldElement>) -> Polynomial {
        assert!(domain.len() == values.len());
        assert!(domain.len() > 0);
        let field = domain[0].field;
        let (v, mut out, mut n) = (raw(values), vec![0u64; domain.len()], 0usize);
        match Self::as_coset(&field, domain) {
            Some((off, lg)) => ffi::check(unsafe { ffi::stark_poly_interpolate_coset(ffi::ctx(), v.as_ptr(), off, lg, out.as_mut_ptr(), &mut n) }),
            None => { let d = raw(domain);
                      ffi::check(unsafe { ffi::stark_poly_interpolate_domain(ffi::ctx(), d.as_ptr(), v.as_ptr(), d.len(), out.as_mut_ptr(), &mut n) }) }
        }
        out.truncate(n);
        Self::from_raw(out, field)
    }
    pub fn zerofier(domain: &Vec<FieldElement>) -> Polynomial {
        let field = domain[0].field;
        let (d, mut out) = (raw(domain), vec![0u64; domain.len() + 1]);
        ffi::check(unsafe { ffi::stark_poly_zerofier_domain(ffi::ctx(), d.as_ptr(), d.len(), out.as_mut_ptr()) });
        Self::from_raw(out, field)
    }
    /// div.rs:6-53 (quotient, remainder); panics "No division by zero"
    pub fn div(numer: &Polynomial, denom: &Polynomial) -> (Polynomial, Polynomial) {
        let (a, b) = (raw(&numer.coeffs), raw(&denom.coeffs));
        let (mut q, mut r) = (vec![0u64; a.len() + 1], vec![0u64; a.len() + b.len() + 1]);
        let (mut nq, mut nr) = (0usize, 0usize);
        ffi::check(unsafe { ffi::stark_poly_div(ffi::ctx(), a.as_ptr(), a.len(), b.as_ptr(), b.len(), q.as_mut_ptr(), &mut nq, r.as_mut_ptr(), &mut nr) });
        q.truncate(nq);
        r.truncate(nr);
        (Self::from_raw(q, numer.field), Self::from_raw(r, numer.field))
    }
    pub fn intdiv(numer: &Polynomial, denom: &Polynomial) -> Polynomial { let (q, r) = Self::div(numer, denom); assert!(r.is_zero()); q }
    pub fn modulo(numer: &Polynomial, denom: &Polynomial) -> Polynomial { Self::div(numer, denom).1 }
    pub fn scale(&self, factor: &FieldElement) -> Polynomial {
        let (c, mut out) = (raw(&self.coeffs), vec![0u64; self.coeffs.len()]);
        ffi::check(unsafe { ffi::stark_poly_scale(ffi::ctx(), c.as_ptr(), c.len(), factor.value, out.as_mut_ptr()) });
        Self::from_raw(out, self.field)
    }
    pub fn test_colinearity(points: &Vec<(FieldElement, FieldElement)>) -> bool {
        assert!(points.len() >= 2, "At least 2 points to test colinearity");
        Self::interpolate_domain(&points.iter().map(|p| p.0).collect(), &points.iter().map(|p| p.1).collect()).deg() <= 1
    }
}
impl std::ops::Add<&Polynomial> for &Polynomial { type Output = Polynomial; fn add(self, r: &Polynomial) -> Polynomial { Polynomial::add(self, r) } }
impl std::ops::Sub<&Polynomial> for &Polynomial { type Output = Polynomial; fn sub(self, r: &Polynomial) -> Polynomial { Polynomial::sub(self, r) } }
impl std::ops::Mul<&Polynomial> for &Polynomial { type Output = Polynomial; fn mul(self, r: &Polynomial) -> Polynomial { Polynomial::mul(self, r) } }
impl std::ops::Div<&Polynomial> for &Polynomial { type Output = (Polynomial, Polynomial); fn div(self, r: &Polynomial) -> (Polynomial, Polynomial) { Polynomial::div(self, r) } }
impl std::ops::BitXor<u64> for &Polynomial { type Output = Polynomial; fn bitxor(self, e: u64) -> Polynomial { Polynomial::exp(self, e) } }
impl std::ops::Rem<&Polynomial> for &Polynomial { type Output = Polynomial; fn rem(self, r: &Polynomial) -> Polynomial { Polynomial::modulo(self, r) } }
