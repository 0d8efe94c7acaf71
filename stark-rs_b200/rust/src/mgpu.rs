//! Groups of GPUs (include/stark_b200.h, "groups of GPUs"): one `Group` per rank.  The reference has no parallel path; this
//! is the host side of the sharded prover -- the library owns NCCL and the peer-memory windows, the host only hands the
//! 128-byte id from rank 0 to the others (any channel: a file, a socket, MPI).  NOT compiled here (no rustc in the image).
#![allow(dead_code)]
use crate::ffi::{self, StarkMgpu};
use crate::hash::Hash;

pub struct Group { handle: *mut StarkMgpu, pub rank: usize, pub world: usize }

impl Group {
    /// rank 0 only: the id every rank passes to `init`
    pub fn unique_id() -> [u8; 128] {
        let mut id = [0u8; 128];
        ffi::check(unsafe { ffi::stark_mgpu_unique_id(id.as_mut_ptr()) });
        id
    }
    /// collective over the group; `max_codeword` = the longest LDE column / codeword the group will fold
    pub fn init(id: &[u8; 128], rank: usize, world: usize, max_codeword: usize) -> Group {
        let mut h = std::ptr::null_mut();
        ffi::check(unsafe { ffi::stark_mgpu_init(ffi::ctx(), id.as_ptr(), rank as i32, world as i32, max_codeword, &mut h) });
        Group { handle: h, rank, world }
    }
    /// BASELINE config 3 on the group (stark_mgpu_prove_trace): `cols` is the whole column-major trace on every rank; every
    /// rank gets all column roots and the complete ProofStream::serialize bytes (identical to the single-GPU path)
    pub fn prove_trace(&self, cols: &[u64], n_cols: usize, log_blowup: u32, offset: u64, num_colinearity_tests: usize) -> (Vec<Hash>, Vec<u8>) {
        let n = cols.len() / n_cols;
        let log_n = n.trailing_zeros();
        let mut cap = 0usize;
        ffi::check(unsafe { ffi::stark_fri_proof_size(n << log_blowup, 1u32 << log_blowup, num_colinearity_tests as u32, &mut cap) });
        let (mut roots, mut proof, mut len) = (vec![Hash([0; 32]); n_cols], vec![0u8; cap], 0usize);
        let (ranks, rp, pp) = ([self.handle], [roots.as_mut_ptr() as *mut u8], [proof.as_mut_ptr()]);
        ffi::check(unsafe {
            ffi::stark_mgpu_prove_trace(ranks.as_ptr(), 1, cols.as_ptr(), n_cols as u32, log_n, log_blowup, offset,
                                        num_colinearity_tests as u32, rp.as_ptr(), pp.as_ptr(), cap, &mut len)
        });
        proof.truncate(len);
        (roots, proof)
    }
    /// BASELINE config 4 (stark_mgpu_lde_commit): n_groups groups of group_width columns -> (group roots, commitment)
    pub fn lde_commit(&self, cols: &[u64], n_groups: usize, group_width: usize, log_n: u32, log_blowup: u32, offset: u64) -> (Vec<Hash>, Hash) {
        let (mut roots, mut com) = (vec![Hash([0; 32]); n_groups], Hash([0; 32]));
        let (ranks, rp, cp) = ([self.handle], [roots.as_mut_ptr() as *mut u8], [com.0.as_mut_ptr()]);
        ffi::check(unsafe {
            ffi::stark_mgpu_lde_commit(ranks.as_ptr(), 1, cols.as_ptr(), n_groups as u32, group_width as u32, log_n, log_blowup, offset,
                                       rp.as_ptr(), cp.as_ptr())
        });
        (roots, com)
    }
    pub fn bytes_sent(&self) -> u64 { unsafe { ffi::stark_mgpu_bytes_sent(self.handle) } }
    pub fn barrier(&self) { ffi::check(unsafe { ffi::stark_mgpu_barrier(self.handle) }); }
}
impl Drop for Group { fn drop(&mut self) { unsafe { ffi::stark_mgpu_destroy(self.handle) } } }
