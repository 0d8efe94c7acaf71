//! stark-rs host crate on the B200 back end (source form; see INTEGRATION.md -- not compiled in this repo's image).
//! Same module list as the reference plus `ffi`; unlike the reference's main.rs every module is declared.
mod ff;
mod ffi;
mod fiat_shamir;
mod fri;
mod hash;
mod merkle;
mod mgpu;
mod stream;
mod trace;
pub mod univariate;
mod utils;

use crate::ff::FiniteField;
use crate::fiat_shamir::FiatShamir;
use crate::fri::Fri;
use crate::stream::ProofStream;
use crate::trace::Trace;

const P: u64 = 998244353; // 119 * 2^23 + 1

/// BASELINE config 1: Fibonacci trace column -> LDE (blowup 4, offset 3) -> Fri::prove -> proof bytes -> Fri::verify.
fn main() {
    let field = FiniteField::new(P);
    let rows = 64usize;
    let column: Vec<u64> = Trace::fibonacci(rows).get_col(0).iter().map(|&v| (v as u64) % P).collect();
    let offset = field.g();
    let codeword = trace::lde(&[column], 2, offset.value).remove(0);
    let omega = field.prim_nth_root((rows * 4) as u64);
    let fri = Fri::new(omega, offset, rows * 4, 4, 8);
    let (mut transcript, mut stream) = (FiatShamir::new(), ProofStream::new());
    let code: Vec<_> = codeword.iter().map(|&v| field.new_element(v)).collect();
    let top = fri.prove(code, &mut transcript, &mut stream);
    println!("proof: {} bytes, top-level indices {:?}", stream.serialize().len(), top);
    // the reference's own test flow (fri.rs:563-570): a fresh transcript, verify, and the opened points lie on the codeword
    let mut points = Vec::new();
    let ok = fri.verify(&mut stream, &mut FiatShamir::new(), &mut points);
    assert!(ok && points.iter().all(|(i, v)| codeword[*i] == v.value));
    // the same statement straight from the trace container (row-major i128 rows): identical proof bytes
    let (roots, proof) = Trace::fibonacci(rows).prove(2, offset.value, 8);
    println!("verified: {}, {} column root(s), {} proof bytes from Trace::prove", ok, roots.len(), proof.len());
}
