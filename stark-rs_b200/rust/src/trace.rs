//! Trace container (reference trace.rs) + the low-degree extension the reference composes by hand (fri.rs:575-578).
#![allow(dead_code)]
use crate::ff::{FieldElement, FiniteField};
use crate::ffi;

#[derive(Clone)]
pub struct Trace { pub trace: Vec<Vec<i128>>, pub num_columns: usize }
impl Trace {
    pub fn new(rows: &Vec<Vec<i128>>) -> Trace { Trace { trace: rows.to_vec(), num_columns: rows[0].len() } }
    pub fn get_row(&self, i: usize) -> Option<&Vec<i128>> { self.trace.get(i) }
    pub fn get_col(&self, j: usize) -> Vec<i128> { self.trace.iter().map(|r| r[j]).collect() }
    pub fn get(&self, i: usize, j: usize) -> Option<i128> { self.trace.get(i)?.get(j).copied() }
    pub fn to_field_elements(&self, field: FiniteField) -> Vec<Vec<FieldElement>> {
        self.trace.iter().map(|r| r.iter().map(|&e| field.new_element(e as u64)).collect()).collect()
    }
    pub fn fibonacci(length: usize) -> Trace {
        let (mut a, mut b) = (1i128, 1i128);
        let rows: Vec<Vec<i128>> = (0..length).map(|_| { let r = vec![a]; (a, b) = (b, a + b); r }).collect();
        Trace::new(&rows)
    }
}

/// columns (values on w_n^i, canonical) -> values on offset * w_{n*2^log_blowup}^i, natural order; stark_lde
pub fn lde(columns: &[Vec<u64>], log_blowup: u32, offset: u64) -> Vec<Vec<u64>> {
    let n = columns[0].len();
    assert!(n.is_power_of_two() && columns.iter().all(|c| c.len() == n));
    let flat: Vec<u64> = columns.iter().flatten().copied().collect();
    let big = n << log_blowup;
    let mut out = vec![0u64; big * columns.len()];
    ffi::check(unsafe { ffi::stark_lde(ffi::ctx(), flat.as_ptr(), columns.len() as u32, n.trailing_zeros(), log_blowup, offset, out.as_mut_ptr()) });
    out.chunks(big).map(|c| c.to_vec()).collect()
}
