//! Tagged proof stream (reference stream.rs): 0 root | 1 value | 2 count+values | 3 count+hashes, all LE.
use crate::ff::{FieldElement, FiniteField};
use crate::hash::Hash;

#[derive(Clone, Debug)]
pub enum ProofObject { MerkleRoot(Hash), FieldElement(FieldElement), FieldElements(Vec<FieldElement>), MerklePath(Vec<Hash>) }
pub struct ProofStream { pub objects: Vec<ProofObject> }

impl ProofStream {
    pub fn new() -> Self { ProofStream { objects: Vec::new() } }
    pub fn push(&mut self, o: ProofObject) { self.objects.push(o); }
    pub fn pop(&mut self) -> Option<ProofObject> { if self.objects.is_empty() { None } else { Some(self.objects.remove(0)) } }
    pub fn serialize(&self) -> Vec<u8> {
        let mut b = Vec::new();
        for o in &self.objects {
            match o {
                ProofObject::MerkleRoot(h) => { b.push(0); b.extend_from_slice(&h.0); }
                ProofObject::FieldElement(e) => { b.push(1); b.extend_from_slice(&e.value.to_le_bytes()); }
                ProofObject::FieldElements(v) => { b.push(2); b.extend_from_slice(&(v.len() as u64).to_le_bytes()); for e in v { b.extend_from_slice(&e.value.to_le_bytes()); } }
                ProofObject::MerklePath(p) => { b.push(3); b.extend_from_slice(&(p.len() as u64).to_le_bytes()); for h in p { b.extend_from_slice(&h.0); } }
            }
        }
        b
    }
    /// lenient parser: truncated items are dropped, an unknown tag stops (stream.rs:66-168)
    pub fn deserialize(bytes: &[u8], field: FiniteField) -> Self {
        let rd = |i: usize| u64::from_le_bytes(bytes[i..i + 8].try_into().unwrap());
        let (mut objects, mut i) = (Vec::new(), 0usize);
        while i < bytes.len() {
            let tag = bytes[i]; i += 1;
            match tag {
                0 => if i + 32 <= bytes.len() { objects.push(ProofObject::MerkleRoot(Hash(bytes[i..i + 32].try_into().unwrap()))); i += 32; },
                1 => if i + 8 <= bytes.len() { objects.push(ProofObject::FieldElement(field.new_element(rd(i)))); i += 8; },
                2 | 3 => if i + 8 <= bytes.len() {
                    let (n, w) = (rd(i) as usize, if tag == 2 { 8 } else { 32 }); i += 8;
                    let (mut vs, mut hs) = (Vec::new(), Vec::new());
                    for _ in 0..n { if i + w <= bytes.len() {
                        if tag == 2 { vs.push(field.new_element(rd(i))); } else { hs.push(Hash(bytes[i..i + 32].try_into().unwrap())); }
                        i += w; } }
                    objects.push(if tag == 2 { ProofObject::FieldElements(vs) } else { ProofObject::MerklePath(hs) });
                },
                _ => break,
            }
        }
        ProofStream { objects }
    }
}
