//! emit_golden.rs -- prints, FROM THE REAL stark-rs CRATE, the values this repository's oracle is pinned against only by
//! survey-derived vectors (tests/golden/survey_vectors.json): hash digests, Merkle roots, Fiat-Shamir challenges, FRI
//! roots / indices and the serialized proofs of the four statements of src/fri.rs:532-693.
//!
//! It cannot run in this repository's image (no rustc / cargo).  A maintainer with a Rust toolchain closes the pin with
//! three commands and WITHOUT editing any reference file (the orphan modules fri.rs / merkle.rs / hash.rs / ... are
//! pulled in with #[path], because the reference's main.rs does not declare them):
//!
//!     mkdir -p <stark-rs>/src/bin && cp tools/emit_golden.rs <stark-rs>/src/bin/emit_golden.rs
//!     (cd <stark-rs> && cargo run --release --bin emit_golden) > reference_golden.json
//!     python tools/check_golden.py reference_golden.json        # compares with the oracle and the survey vectors
//!
//! Output: one JSON object on stdout (hex strings for bytes).
#![allow(dead_code, unused_imports)]
#[path = "../ff.rs"]
mod ff;
#[path = "../fiat_shamir.rs"]
mod fiat_shamir;
#[path = "../fri.rs"]
mod fri;
#[path = "../hash.rs"]
mod hash;
#[path = "../merkle.rs"]
mod merkle;
#[path = "../stream.rs"]
mod stream;
#[path = "../trace.rs"]
mod trace;
#[path = "../univariate/mod.rs"]
pub mod univariate;
#[path = "../utils.rs"]
mod utils;

use crate::ff::{FieldElement, FiniteField};
use crate::fiat_shamir::FiatShamir;
use crate::fri::Fri;
use crate::hash::Hash;
use crate::merkle::MerkleTree;
use crate::stream::{ProofObject, ProofStream};
use crate::univariate::Polynomial;

const P: u64 = 998244353;

fn hex(b: &[u8]) -> String {
    b.iter().map(|x| format!("{:02x}", x)).collect()
}

/// one FRI statement: codeword = poly(coeffs) on offset * omega^i, proved with a fresh transcript
fn fri_case(n: usize, offset: u64, ef: usize, nq: usize, coeffs: &[u64]) -> String {
    let field = FiniteField::new(P);
    let omega = field.prim_nth_root(n as u64);
    let off = field.new_element(offset);
    let fri = Fri::new(omega, off, n, ef, nq);
    let poly = Polynomial::new(coeffs.iter().map(|&c| field.new_element(c)).collect(), field);
    let domain: Vec<FieldElement> = (0..n).map(|i| field.mul(&off, &field.exp(&omega, i as u64))).collect();
    let codeword = poly.eval_domain(&domain);
    let mut stream = ProofStream::new();
    let mut fs = FiatShamir::new();
    let top = fri.prove(codeword.clone(), &mut fs, &mut stream);
    let bytes = stream.serialize();
    // the roots and challenges as the verifier re-derives them
    let mut fs2 = FiatShamir::new();
    let mut roots = Vec::new();
    let mut alphas = Vec::new();
    for o in stream.objects.iter() {
        if let ProofObject::MerkleRoot(h) = o {
            fs2.absorb(&h.0);
            roots.push(format!("\"{}\"", hex(&h.0)));
            alphas.push(fs2.challenge(&field).value.to_string());
        }
    }
    let mut verifier_stream = ProofStream::deserialize(&bytes, field);
    let mut points = Vec::new();
    let ok = fri.verify(&mut verifier_stream, &mut FiatShamir::new(), &mut points);
    format!(
        "{{\"n\": {}, \"offset\": {}, \"ef\": {}, \"nq\": {}, \"coeffs\": {:?}, \"codeword\": {:?}, \"bytes\": {}, \"objects\": {}, \"top\": {:?}, \"roots\": [{}], \"challenges_after_each_root\": [{}], \"verify\": {}, \"proof_hex\": \"{}\"}}",
        n, offset, ef, nq, coeffs, codeword.iter().map(|e| e.value).collect::<Vec<u64>>(), bytes.len(), stream.objects.len(), top,
        roots.join(", "), alphas.join(", "), ok, hex(&bytes)
    )
}

fn main() {
    let field = FiniteField::new(P);
    let mut out: Vec<String> = Vec::new();
    // hash.rs:7-46
    let msgs: [&[u8]; 6] = [b"hello", b"world", b"", &[0u8], &[7u8; 32], &[9u8; 33]];
    let hb: Vec<String> = msgs.iter().map(|m| format!("\"{}\": \"{}\"", hex(m), hex(&Hash::from_bytes(m).0))).collect();
    out.push(format!("\"hash_from_bytes\": {{{}}}", hb.join(", ")));
    out.push(format!("\"hash_from_u64_0\": \"{}\"", hex(&Hash::from_u64(0).0)));
    out.push(format!("\"hash_from_field_elements_1\": \"{}\"", hex(&Hash::from_field_elements(&[1]).0)));
    out.push(format!("\"hash_from_field_elements_8\": \"{}\"", hex(&Hash::from_field_elements(&[1, 2, 3, 4, 5, 6, 7, 998244352]).0)));
    out.push(format!("\"combine_zero_zero\": \"{}\"", hex(&Hash::combine(&Hash([0; 32]), &Hash([0; 32])).0)));
    // merkle.rs:11-80
    for n in [4usize, 8, 16] {
        let leaves: Vec<Hash> = (0..n).map(|i| Hash::from_bytes(&[i as u8])).collect();
        let tree = MerkleTree::new(&leaves);
        out.push(format!("\"merkle_root_{}\": \"{}\"", n, hex(&tree.get_root().0)));
        let path: Vec<String> = tree.open(n - 3).iter().map(|h| format!("\"{}\"", hex(&h.0))).collect();
        out.push(format!("\"merkle_open_{}_{}\": [{}]", n, n - 3, path.join(", ")));
    }
    // fiat_shamir.rs:19-25 (raw, unreduced)
    let mut fs = FiatShamir::new();
    out.push(format!("\"challenge_empty\": {}", fs.challenge(&field).value));
    fs.absorb(b"stark-rs");
    out.push(format!("\"challenge_stark_rs\": {}", fs.challenge(&field).value));
    // ff.rs:215-223
    let roots: Vec<String> = [3u32, 16, 20, 22, 23].iter().map(|&k| format!("\"{}\": {}", k, field.prim_nth_root(1u64 << k).value)).collect();
    out.push(format!("\"roots_of_unity\": {{{}}}", roots.join(", ")));
    // fri.rs:57-91 on a small codeword with an UNREDUCED alpha
    {
        let n = 16usize;
        let omega = field.prim_nth_root(n as u64);
        let off = field.new_element(3);
        let fri = Fri::new(omega, off, n, 4, 2);
        let cw: Vec<FieldElement> = (0..n).map(|i| field.new_element((i as u64 * 1234567 + 89) % P)).collect();
        let alpha = field.new_element(15764728482632548394);
        let folded = fri.fold_codeword(&cw, &alpha, &off, &omega);
        out.push(format!("\"fold_16\": {{\"codeword\": {:?}, \"alpha_raw\": {}, \"folded\": {:?}}}",
                         cw.iter().map(|e| e.value).collect::<Vec<u64>>(), alpha.value, folded.iter().map(|e| e.value).collect::<Vec<u64>>()));
        out.push(format!("\"sample_indices_seed_abc\": {:?}", fri.sample_indices(b"abc", 64, 8, 5)));
    }
    // fri.rs:532-693: the four prove -> verify statements
    let cases = [fri_case(32, 3, 4, 2, &[5]), fri_case(64, 7, 4, 3, &[5, 3]), fri_case(128, 13, 4, 4, &[1, 3, 2]),
                 fri_case(256, 17, 8, 5, &[1, 2, 5, 3, 7, 4, 1, 2])];
    out.push(format!("\"fri_proofs\": [{}]", cases.join(", ")));
    println!("{{{}}}", out.join(",\n "));
}
