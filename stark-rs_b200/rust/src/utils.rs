//! Extended Euclid (reference utils.rs:3-13 has the recursive form; this is the iterative one, same outputs
//! up to the usual Bezout normalisation -- only `inv`'s reduced result is observable).
pub fn xgcd(x: u64, y: u64) -> (i128, i128, i128) {
    let (mut r0, mut r1) = (x as i128, y as i128);
    let (mut s0, mut s1, mut t0, mut t1) = (1i128, 0i128, 0i128, 1i128);
    while r1 != 0 {
        let q = r0 / r1;
        (r0, r1) = (r1, r0 - q * r1);
        (s0, s1) = (s1, s0 - q * s1);
        (t0, t1) = (t1, t0 - q * t1);
    }
    (r0, s0, t0)
}
