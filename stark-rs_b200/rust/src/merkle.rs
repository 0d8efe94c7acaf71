//! MerkleTree (reference merkle.rs:4-97) with every level resident on the GPU.
#![allow(dead_code)]
use crate::ffi::{self, StarkTree};
use crate::hash::Hash;

pub struct MerkleTree { pub leaves: Vec<Hash>, pub root: Hash, handle: *mut StarkTree }

impl MerkleTree {
    pub fn new(leaves: &Vec<Hash>) -> Self {
        let mut h = std::ptr::null_mut();
        // asserts "Cannot create tree from empty leaves" / "Number of leaves must be power of 2" come back as status 1
        ffi::check(unsafe { ffi::stark_merkle_build(ffi::ctx(), leaves.as_ptr() as *const u8, leaves.len(), &mut h) });
        let mut root = Hash([0; 32]);
        ffi::check(unsafe { ffi::stark_merkle_root(h, root.0.as_mut_ptr()) });
        MerkleTree { leaves: leaves.clone(), root, handle: h }
    }
    pub fn get_root(&self) -> &Hash { &self.root }
    pub fn commit(leaves: &Vec<Hash>) -> Hash {
        let mut root = Hash([0; 32]);
        ffi::check(unsafe { ffi::stark_merkle_commit(ffi::ctx(), leaves.as_ptr() as *const u8, leaves.len(), root.0.as_mut_ptr()) });
        root
    }
    /// `nodes[level]` of the reference struct, fetched on demand
    pub fn level(&self, level: usize) -> Vec<Hash> {
        let mut out = vec![Hash([0; 32]); self.leaves.len() >> level];
        ffi::check(unsafe { ffi::stark_merkle_level(self.handle, level as u32, out.as_mut_ptr() as *mut u8) });
        out
    }
    pub fn open(&self, index: usize) -> Vec<Hash> {
        let depth = unsafe { ffi::stark_merkle_num_levels(self.handle) } as usize;
        let mut out = vec![Hash([0; 32]); depth.max(1)];
        let mut n = 0usize;
        ffi::check(unsafe { ffi::stark_merkle_open(self.handle, index, out.as_mut_ptr() as *mut u8, &mut n) }); // "Index out of bounds"
        out.truncate(n);
        out
    }
    /// MerkleTree::open for many leaves in one device gather (stark_merkle_open_batch); "Index out of bounds" as open()
    pub fn open_batch(&self, indices: &[usize]) -> Vec<Vec<Hash>> {
        let depth = (unsafe { ffi::stark_merkle_num_levels(self.handle) } as usize).saturating_sub(1);
        let idx: Vec<u64> = indices.iter().map(|&i| i as u64).collect();
        let mut out = vec![Hash([0; 32]); (depth * idx.len()).max(1)];
        ffi::check(unsafe { ffi::stark_merkle_open_batch(self.handle, idx.as_ptr(), idx.len(), out.as_mut_ptr() as *mut u8) });
        (0..idx.len()).map(|q| out[q * depth..(q + 1) * depth].to_vec()).collect()
    }
    pub fn verify(leaf: &Hash, index: usize, proof: &[Hash], root: &Hash) -> bool {
        let (mut cur, mut idx) = (*leaf, index);
        for sib in proof {
            cur = if idx & 1 == 0 { Hash::combine(&cur, sib) } else { Hash::combine(sib, &cur) };
            idx >>= 1;
        }
        cur == *root
    }
}
impl Drop for MerkleTree { fn drop(&mut self) { unsafe { ffi::stark_merkle_free(self.handle) } } }
