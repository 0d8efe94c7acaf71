//! Append-only transcript (reference fiat_shamir.rs): challenge = first 8 digest bytes, LE, unreduced.
use crate::ff::{FieldElement, FiniteField};
use crate::hash::Hash;

pub struct FiatShamir { pub transcript: Vec<u8> }
impl FiatShamir {
    pub fn new() -> Self { FiatShamir { transcript: Vec::new() } }
    pub fn absorb(&mut self, data: &[u8]) { self.transcript.extend_from_slice(data); }
    pub fn challenge(&self, field: &FiniteField) -> FieldElement {
        let d = Hash::from_bytes(&self.transcript).0;
        field.new_element(u64::from_le_bytes(d[..8].try_into().unwrap()))
    }
}
