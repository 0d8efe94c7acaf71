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
    /// Every column as canonical residues on the device (`e as u64` like to_field_elements, then mod p as
    /// FiniteField::mul/add see it), column-major: stark_trace_to_columns.  Rows must all have num_columns entries.
    fn flat(&self) -> Vec<i128> {
        assert!(self.trace.iter().all(|r| r.len() == self.num_columns));
        self.trace.iter().flatten().copied().collect()
    }
    /// LDE (blowup 2^log_blowup, coset offset) + per-column Merkle roots + Fri::prove of column 0, straight from the
    /// row-major i128 rows: stark_prove_trace_rows.  Returns (column roots, ProofStream::serialize bytes).
    pub fn prove(&self, log_blowup: u32, offset: u64, num_colinearity_tests: u32) -> (Vec<[u8; 32]>, Vec<u8>) {
        let n = self.trace.len();
        assert!(n.is_power_of_two(), "n must be a power of two");
        let flat = self.flat();   // i128 is 16 little-endian bytes on every target this library runs on
        let mut cap = 0usize;
        ffi::check(unsafe { ffi::stark_fri_proof_size(n << log_blowup, 1 << log_blowup, num_colinearity_tests, &mut cap) });
        let (mut roots, mut proof, mut len) = (vec![[0u8; 32]; self.num_columns], vec![0u8; cap], 0usize);
        ffi::check(unsafe { ffi::stark_prove_trace_rows(ffi::ctx(), flat.as_ptr() as *const std::ffi::c_void, self.num_columns as u32,
            n.trailing_zeros(), log_blowup, offset, num_colinearity_tests, roots.as_mut_ptr() as *mut u8, proof.as_mut_ptr(), cap, &mut len) });
        proof.truncate(len);
        (roots, proof)
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
