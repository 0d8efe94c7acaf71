//! Hash (reference hash.rs): single digests on the host, bulk digests on the GPU (stark_hash_bytes / _leaves).
#![allow(dead_code)]
use crate::ffi;

#[derive(Debug, Clone, Copy, PartialEq, Eq, Hash)]
pub struct Hash(pub [u8; 32]);

const SEED: [u8; 16] = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53];
const RC: [u8; 32] = [0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1b, 0x36, 0x6c, 0xd8, 0xab, 0x4d, 0x9a, 0x2f,
                      0x5e, 0xbc, 0x63, 0xc6, 0x97, 0x35, 0x6a, 0xd4, 0xb3, 0x7d, 0xfa, 0xef, 0xc5, 0x91, 0x39, 0x72];

fn permute(s: &mut [u8; 32]) {
    for b in s.iter_mut() { *b = b.wrapping_mul(251).rotate_left(1) ^ 0x63; }
    for q in s.chunks_exact_mut(4) {
        let x = q[0] ^ q[1] ^ q[2] ^ q[3];
        let (a, b, c, d) = (q[0], q[1], q[2], q[3]);
        q[0] = x ^ c; q[1] = x ^ b; q[2] = x ^ d; q[3] = x ^ a;
    }
    for i in 0..32 { s[i] = s[i].wrapping_add(s[(i + 1) % 32]).wrapping_add(s[(i + 31) % 32]); }
    for i in 0..32 { s[i] = s[i].wrapping_add(RC[i]); }
}

impl Hash {
    pub fn from_bytes(bytes: &[u8]) -> Self {
        let mut s = [0u8; 32];
        for i in 0..32 { s[i] = SEED[i % 16]; }
        for block in bytes.chunks(32) {
            for (i, &b) in block.iter().enumerate() {
                s[i] = s[i].wrapping_add(b).rotate_left(3);
                s[(i + 7) % 32] ^= s[i];
            }
            permute(&mut s);
        }
        for _ in 0..8 { permute(&mut s); }
        Hash(s)
    }
    pub fn from_field_elements(e: &[u64]) -> Self { Self::from_bytes(&e.iter().flat_map(|v| v.to_le_bytes()).collect::<Vec<u8>>()) }
    pub fn from_u64(v: u64) -> Self { Self::from_bytes(&v.to_le_bytes()) }
    pub fn combine(l: &Hash, r: &Hash) -> Self { Self::from_bytes(&[l.0, r.0].concat()) }
    pub fn to_hex(&self) -> String { self.0.iter().map(|b| format!("{:02x}", b)).collect() }
    /// leaf i = from_field_elements(&vals[i*width..(i+1)*width]) for all i, on the GPU
    pub fn leaves(vals: &[u64], width: usize) -> Vec<Hash> {
        let n = vals.len() / width;
        let mut out = vec![Hash([0; 32]); n];
        ffi::check(unsafe { ffi::stark_hash_leaves(ffi::ctx(), vals.as_ptr(), n, width as u32, out.as_mut_ptr() as *mut u8) });
        out
    }
}
