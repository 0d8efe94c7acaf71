//! Scalar field arithmetic stays on the host (a GPU call per scalar is meaningless); vectors go through
//! `vec_*` -> stark_ff_vec_* (reference ff.rs:108-233 for the API surface and panic messages).
#![allow(dead_code)]
use crate::ffi;
use crate::utils::xgcd;
use std::ops::{Add, BitXor, Div, Mul, Neg, Sub};

#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub struct FiniteField { p: u64 }

#[derive(Debug, Clone, Copy, PartialEq, Eq, PartialOrd, Ord)]
pub struct FieldElement { pub value: u64, pub field: FiniteField }

impl PartialOrd for FiniteField { fn partial_cmp(&self, o: &Self) -> Option<std::cmp::Ordering> { self.p.partial_cmp(&o.p) } }
impl Ord for FiniteField { fn cmp(&self, o: &Self) -> std::cmp::Ordering { self.p.cmp(&o.p) } }

impl FiniteField {
    pub fn new(p: u64) -> Self { Self { p } }
    pub fn modulus(&self) -> u64 { self.p }
    pub fn new_element(&self, value: u64) -> FieldElement { FieldElement { value, field: *self } } // not reduced
    pub fn zero(&self) -> FieldElement { self.new_element(0) }
    pub fn one(&self) -> FieldElement { self.new_element(1) }
    fn wrap(&self, v: u128) -> FieldElement { self.new_element((v % self.p as u128) as u64) }
    pub fn mul(&self, l: &FieldElement, r: &FieldElement) -> FieldElement { self.wrap(l.value as u128 * r.value as u128) }
    pub fn add(&self, l: &FieldElement, r: &FieldElement) -> FieldElement { self.wrap(l.value as u128 + r.value as u128) }
    pub fn sub(&self, l: &FieldElement, r: &FieldElement) -> FieldElement { self.wrap(self.p as u128 + l.value as u128 - r.value as u128) }
    pub fn neg(&self, x: &FieldElement) -> FieldElement { self.new_element((self.p - x.value) % self.p) }
    pub fn inv(&self, x: &FieldElement) -> FieldElement {
        let (g, a, _) = xgcd(x.value, self.p);
        assert!(g == 1, "no inverse");
        self.new_element(a.rem_euclid(self.p as i128) as u64)
    }
    pub fn div(&self, l: &FieldElement, r: &FieldElement) -> FieldElement {
        assert!(r.value != 0, "no division by zero");
        self.mul(l, &self.inv(r))
    }
    pub fn g(&self) -> FieldElement { assert!(self.p == 998244353); self.new_element(3) }
    pub fn exp(&self, base: &FieldElement, mut e: u64) -> FieldElement {
        let (mut acc, mut b) = (self.one(), *base);
        while e > 0 {
            if e & 1 == 1 { acc = self.mul(&acc, &b); }
            b = self.mul(&b, &b);
            e >>= 1;
        }
        acc
    }
    pub fn prim_nth_root(&self, n: u64) -> FieldElement {
        assert!(self.p == 998244353);
        let mut out = 0u64;
        ffi::check(unsafe { ffi::stark_ff_prim_nth_root(n, &mut out) }); // same two panic messages as ff.rs:217-218
        self.new_element(out)
    }
    pub fn sample(&self, salt: &[u8]) -> FieldElement {
        let p = self.p as u128;
        let mut acc = 0u128;
        for &b in salt { acc = (acc << 8) % p; acc = (acc ^ b as u128) % p; }
        self.new_element(acc as u64)
    }
    /// batch forms (stark_ff_vec_*): element-wise over canonical slices
    pub fn vec_mul(&self, a: &[u64], b: &[u64]) -> Vec<u64> {
        assert_eq!(a.len(), b.len());
        let mut out = vec![0u64; a.len()];
        ffi::check(unsafe { ffi::stark_ff_vec_mul(ffi::ctx(), a.as_ptr(), b.as_ptr(), out.as_mut_ptr(), a.len()) });
        out
    }
    pub fn vec_inv(&self, a: &[u64]) -> Vec<u64> {
        let mut out = vec![0u64; a.len()];
        ffi::check(unsafe { ffi::stark_ff_vec_inv(ffi::ctx(), a.as_ptr(), out.as_mut_ptr(), a.len()) });
        out
    }
}

impl FieldElement { pub fn pow(&self, e: u64) -> FieldElement { self.field.exp(self, e) } }
macro_rules! binop { ($tr:ident, $f:ident) => {
    impl $tr for FieldElement { type Output = FieldElement; fn $f(self, r: Self) -> FieldElement { self.field.$f(&self, &r) } }
    impl $tr<&FieldElement> for &FieldElement { type Output = FieldElement; fn $f(self, r: &FieldElement) -> FieldElement { self.field.$f(self, r) } }
} }
binop!(Add, add); binop!(Sub, sub); binop!(Mul, mul); binop!(Div, div);
impl Neg for FieldElement { type Output = FieldElement; fn neg(self) -> FieldElement { self.field.neg(&self) } }
impl Neg for &FieldElement { type Output = FieldElement; fn neg(self) -> FieldElement { self.field.neg(self) } }
impl BitXor<u64> for FieldElement { type Output = FieldElement; fn bitxor(self, e: u64) -> FieldElement { self.field.exp(&self, e) } }
impl BitXor<u64> for &FieldElement { type Output = FieldElement; fn bitxor(self, e: u64) -> FieldElement { self.field.exp(self, e) } }
