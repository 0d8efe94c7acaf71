//! Fri (reference fri.rs:8-311): prover (stark_fri_prove) and verifier (stark_fri_verify) run on the GPU.
#![allow(dead_code)]
use crate::ff::{FieldElement, FiniteField};
use crate::ffi;
use crate::fiat_shamir::FiatShamir;
use crate::stream::{ProofObject, ProofStream};

pub struct Fri {
    pub offset: FieldElement, pub omega: FieldElement, pub domain_length: usize, pub field: FiniteField,
    pub expansion_factor: usize, pub num_colinearity_tests: usize,
}

impl Fri {
    pub fn new(omega: FieldElement, offset: FieldElement, domain_length: usize, expansion_factor: usize, num_colinearity_tests: usize) -> Self {
        assert!(domain_length.is_power_of_two(), "Domain length must be power of 2");
        assert!(expansion_factor.is_power_of_two(), "Expansion factor must be power of 2");
        assert!(expansion_factor >= 4, "Expansion factor must be at least 4");
        Fri { omega, offset, domain_length, field: omega.field, expansion_factor, num_colinearity_tests }
    }
    pub fn num_rounds(&self) -> u64 {
        let mut r = 0u32;
        ffi::check(unsafe { ffi::stark_fri_num_rounds(self.domain_length, self.expansion_factor as u32, self.num_colinearity_tests as u32, &mut r) });
        r as u64
    }
    /// fri.rs:57-91 for a whole codeword; alpha may be unreduced
    pub fn fold_codeword(&self, codeword: &[FieldElement], alpha: &FieldElement, offset: &FieldElement, omega: &FieldElement) -> Vec<FieldElement> {
        let vals: Vec<u64> = codeword.iter().map(|e| e.value).collect();
        let mut out = vec![0u64; vals.len() / 2];
        ffi::check(unsafe { ffi::stark_fri_fold(ffi::ctx(), vals.as_ptr(), vals.len(), alpha.value, offset.value, omega.value, out.as_mut_ptr()) });
        out.into_iter().map(|v| self.field.new_element(v)).collect()
    }
    /// fri.rs:105-156: the whole round loop runs on the device (stark_fri_commit: leaf hashes, trees, transcript, folds);
    /// the roots and the last codeword are pushed / absorbed here in the reference's order and every intermediate codeword
    /// is returned, like the reference does.
    pub fn commit(&self, initial_codeword: Vec<FieldElement>, proof_stream: &mut ProofStream, fiat_shamir: &mut FiatShamir) -> Vec<Vec<FieldElement>> {
        let vals: Vec<u64> = initial_codeword.iter().map(|e| e.value).collect();
        let mut st = std::ptr::null_mut();
        ffi::check(unsafe {
            ffi::stark_fri_commit(ffi::ctx(), vals.as_ptr(), vals.len(), self.offset.value, self.omega.value, self.expansion_factor as u32,
                                  self.num_colinearity_tests as u32, fiat_shamir.transcript.as_ptr(), fiat_shamir.transcript.len(), &mut st)
        });
        let rounds = unsafe { ffi::stark_fri_rounds(st) } as usize;
        let mut roots = vec![0u8; 32 * rounds.max(1)];
        ffi::check(unsafe { ffi::stark_fri_roots(st, roots.as_mut_ptr()) });
        for r in 0..rounds {
            let root = crate::hash::Hash(roots[32 * r..32 * r + 32].try_into().unwrap());
            proof_stream.push(ProofObject::MerkleRoot(root));     // fri.rs:129-131
            fiat_shamir.absorb(&root.0);
        }
        let mut codewords = Vec::new();
        for r in 0..rounds.max(1) {
            let mut len = 0usize;
            ffi::check(unsafe { ffi::stark_fri_codeword_len(st, r as u32, &mut len) });
            let mut cw = vec![0u64; len];
            ffi::check(unsafe { ffi::stark_fri_codeword(st, r as u32, cw.as_mut_ptr()) });
            codewords.push(cw.into_iter().map(|v| self.field.new_element(v)).collect::<Vec<_>>());
        }
        proof_stream.push(ProofObject::FieldElements(codewords.last().unwrap().clone()));   // fri.rs:151
        unsafe { ffi::stark_fri_free(st) };
        codewords
    }
    /// fri.rs:176-213 (same asserts, same text)
    pub fn sample_indices(&self, seed: &[u8], size: usize, reduced_size: usize, number: usize) -> Vec<usize> {
        let mut out = vec![0u64; number.max(1)];
        ffi::check(unsafe { ffi::stark_fri_sample_indices(seed.as_ptr(), seed.len(), size, reduced_size, number, out.as_mut_ptr()) });
        out.truncate(number);
        out.into_iter().map(|i| i as usize).collect()
    }
    /// fri.rs:215-248: reveal the triples and their authentication paths (the trees live on the device: one batched
    /// gather per tree through MerkleTree::open_batch)
    pub fn query(&self, current_codeword: &[FieldElement], next_codeword: &[FieldElement], c_indices: &[usize], proof_stream: &mut ProofStream,
                 current_tree: &crate::merkle::MerkleTree, next_tree: &crate::merkle::MerkleTree) -> Vec<usize> {
        let half = current_codeword.len() / 2;
        let a_indices: Vec<usize> = c_indices.to_vec();
        let b_indices: Vec<usize> = a_indices.iter().map(|&i| i + half).collect();
        for s in 0..self.num_colinearity_tests {
            proof_stream.push(ProofObject::FieldElements(vec![current_codeword[a_indices[s]], current_codeword[b_indices[s]], next_codeword[c_indices[s]]]));
        }
        let (pa, pb, pc) = (current_tree.open_batch(&a_indices), current_tree.open_batch(&b_indices), next_tree.open_batch(c_indices));
        for s in 0..self.num_colinearity_tests {
            proof_stream.push(ProofObject::MerklePath(pa[s].clone()));
            proof_stream.push(ProofObject::MerklePath(pb[s].clone()));
            proof_stream.push(ProofObject::MerklePath(pc[s].clone()));
        }
        let mut all = a_indices;
        all.extend(b_indices);
        all
    }
    /// fri.rs:250-311.  The GPU returns ProofStream::serialize's bytes; they are parsed back into `proof_stream`
    /// and the roots are absorbed into `fiat_shamir`, so both end in the state the reference leaves them in.
    pub fn prove(&self, initial_codeword: Vec<FieldElement>, fiat_shamir: &mut FiatShamir, proof_stream: &mut ProofStream) -> Vec<usize> {
        let vals: Vec<u64> = initial_codeword.iter().map(|e| e.value).collect();
        let mut cap = 0usize;
        ffi::check(unsafe { ffi::stark_fri_proof_size(self.domain_length, self.expansion_factor as u32, self.num_colinearity_tests as u32, &mut cap) });
        let (mut proof, mut len) = (vec![0u8; cap], 0usize);
        let mut top = vec![0u64; self.num_colinearity_tests.max(1)];
        ffi::check(unsafe {
            ffi::stark_fri_prove(ffi::ctx(), vals.as_ptr(), vals.len(), self.domain_length, self.offset.value, self.omega.value,
                                 self.expansion_factor as u32, self.num_colinearity_tests as u32, fiat_shamir.transcript.as_ptr(),
                                 fiat_shamir.transcript.len(), proof.as_mut_ptr(), cap, &mut len, top.as_mut_ptr())
        });
        for obj in ProofStream::deserialize(&proof[..len], self.field).objects {
            if let ProofObject::MerkleRoot(r) = &obj { fiat_shamir.absorb(&r.0); }
            proof_stream.push(obj);
        }
        top.truncate(self.num_colinearity_tests);
        top.into_iter().map(|i| i as usize).collect()
    }
    /// fri.rs:313-505 on the device.  On success the objects the reference pops are consumed, the roots are absorbed into
    /// `fiat_shamir` (fri.rs:327) and `polynomial_values` receives the top-layer pairs (fri.rs:437-441); on failure the
    /// reference's println! line is printed and false returned.
    pub fn verify(&self, proof_stream: &mut ProofStream, fiat_shamir: &mut FiatShamir, polynomial_values: &mut Vec<(usize, FieldElement)>) -> bool {
        let bytes = proof_stream.serialize();
        let (nq, rounds) = (self.num_colinearity_tests, self.num_rounds() as usize);
        let (mut ok, mut why) = (0i32, 0u32);
        let mut roots = vec![0u8; 32 * rounds.max(1)];
        let (mut top, mut pi, mut pv) = (vec![0u64; nq.max(1)], vec![0u64; 2 * nq + 1], vec![0u64; 2 * nq + 1]);
        ffi::check(unsafe {
            ffi::stark_fri_verify(ffi::ctx(), bytes.as_ptr(), bytes.len(), self.domain_length, self.offset.value, self.omega.value,
                                  self.expansion_factor as u32, nq as u32, fiat_shamir.transcript.as_ptr(), fiat_shamir.transcript.len(),
                                  &mut ok, &mut why, roots.as_mut_ptr(), top.as_mut_ptr(), pi.as_mut_ptr(), pv.as_mut_ptr())
        });
        if ok == 0 {
            println!("{}", unsafe { std::ffi::CStr::from_ptr(ffi::stark_fri_verify_reason(why)) }.to_string_lossy());
            return false;
        }
        for r in 0..rounds { fiat_shamir.absorb(&roots[32 * r..32 * r + 32]); }
        if rounds > 1 { for i in 0..2 * nq { polynomial_values.push((pi[i] as usize, self.field.new_element(pv[i]))); } }
        for _ in 0..rounds + 1 + (rounds - 1) * 4 * nq { proof_stream.pop(); }
        true
    }
}
