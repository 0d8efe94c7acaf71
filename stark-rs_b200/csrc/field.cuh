// field.cuh -- F_p arithmetic for p = 998244353 = 119*2^23 + 1 on sm_100a.
//
// Replaces FiniteField::{mul,add,sub,neg,inv,div,exp} (reference src/ff.rs:138-213), which reduce
// with a 128-bit `%` per operation.  Here an element is ONE 32-bit limb (p < 2^30) and products are
// reduced by Montgomery REDC with R = 2^32, built from IMAD.WIDE carry-chained multiply-adds:
//     ab  = a*b                  (IMAD.WIDE.U32, 64-bit)
//     m   = lo(ab) * (-p^-1)     (IMAD, mod 2^32)
//     t   = (ab + m*p) >> 32     (IMAD.WIDE.U32 with 64-bit addend; low word cancels)
// For a < 2^32 and b < p the result is in [0, 2p): no correction step on the hot path.
// Two spare bits (4p < 2^32) let butterflies run lazily in [0, 4p) (Harvey-style).
//
// Data convention: values in HBM are CANONICAL (in [0, p)), never in Montgomery form; constants
// (twiddles, scale factors) are kept in Montgomery form so that mont_mul(x, wR) = x*w needs no
// conversion of the data stream.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FF_HD __host__ __device__ __forceinline__
#else
#define FF_HD inline
#endif

namespace ff {

typedef uint32_t u32;
typedef uint64_t u64;

constexpr u32 P = 998244353u;       // ff.rs:191-197, main.rs:6
constexpr u32 P2 = 2u * P;          // 1996488706 < 2^31
constexpr u32 NPINV = 998244351u;   // -p^-1 mod 2^32  (= p - 2)
constexpr u32 R1 = 301989884u;      // 2^32 mod p      (Montgomery form of 1)
constexpr u32 R2 = 932051910u;      // 2^64 mod p
constexpr u32 INV2 = 499122177u;    // 2^-1 mod p
constexpr u32 GEN = 3u;             // ff.rs:191-197 g()
constexpr int TWO_ADICITY = 23;     // ff.rs:218

// a: any u32, b < p  ->  a*b*2^-32 mod p, in [0, 2p)
FF_HD u32 mont_mul(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
  // Exactly three FMA-pipe instructions: IMAD.WIDE.U32 (ab), IMAD (m), IMAD.HI.U32 with the 64-bit addend ab.  The C
  // form below reaches ptxas as 64-bit multiplies; it emits the same three plus an add of the zero cross term
  // (IADD3 Rd, Rhi, UR(=0)) per product on the ALU pipe -- 8 % of an NTT pass, 15 % of its ALU-pipe load.
  u32 r;
  asm("{\n\t.reg .u32 lo, hi, m, tl;\n\t"
      "mul.lo.u32 lo, %1, %2;\n\t"
      "mul.hi.u32 hi, %1, %2;\n\t"
      "mul.lo.u32 m, lo, %3;\n\t"
      "mad.lo.cc.u32 tl, m, %4, lo;\n\t"
      "madc.hi.u32 %0, m, %4, hi;\n\t}"
      : "=r"(r) : "r"(a), "r"(b), "r"(NPINV), "r"(P));
  return r;
#else
  u64 ab = (u64)a * b;
  u32 m = (u32)ab * NPINV;
  return (u32)((ab + (u64)m * P) >> 32);
#endif
}
// Multiplication by a CONSTANT w < p whose companion ws = floor(w * 2^32 / p) is known (Shoup): x any u32 ->
// x*w mod p in [0, 2p), with w in plain (not Montgomery) form.  IMAD.HI + 2 IMAD.  On B200 the 64-bit-result multiplies
// (IMAD.WIDE, IMAD.HI) issue at HALF the rate of a 32-bit IMAD (stark_bench_mul_peak: 9.1 / 9.3 T vs 18.5 T thread-
// instructions/s), so a Montgomery product occupies the FMA-heavy pipe for 2+1+2 = 5 slots and this one for 2+1+1 = 4:
// every product with a table twiddle takes this form; Montgomery stays for products of two run-time values.
FF_HD u32 shoup_mul(u32 x, u32 w, u32 ws) {
#if defined(__CUDA_ARCH__)
  const u32 q = __umulhi(x, ws);
#else
  const u32 q = (u32)(((u64)x * ws) >> 32);
#endif
  return x * w - q * P;
}
FF_HD u32 shoup_of(u32 w) { return (u32)(((u64)w << 32) / P); }
// [0, 4p) -> [0, 2p)
FF_HD u32 red2p(u32 x) {
  u32 y = x - P2;
  return y < x ? y : x;  // unsigned min trick: x - 2p wraps when x < 2p
}
// [0, 2p) -> [0, p)
FF_HD u32 canon(u32 x) {
  u32 y = x - P;
  return y < x ? y : x;
}
FF_HD u32 canon4(u32 x) { return canon(red2p(x)); }
// a + b pinned to the ALU pipe: ptxas turns a plain 32-bit add into IMAD.IADD on the FMA-heavy pipe where it believes
// that pipe has room, but it counts IMAD.WIDE / IMAD.HI as one slot.  A third addend that is zero only at RUN time
// (a kernel parameter) keeps the add a three-input IADD3, which the FMA pipe cannot execute, at no extra instruction.
FF_HD u32 add_alu(u32 a, u32 b, u32 zero) { return a + b + zero; }
// lazy add/sub on [0, 2p) operands -> [0, 4p)
FF_HD u32 add_lazy(u32 a, u32 b) { return a + b; }
FF_HD u32 sub_lazy(u32 a, u32 b) { return a + P2 - b; }
// canonical add/sub/neg (ff.rs:146-167) on canonical operands
FF_HD u32 add(u32 a, u32 b) { return canon(a + b); }
FF_HD u32 sub(u32 a, u32 b) { return canon(a + P - b); }
FF_HD u32 neg(u32 a) { return a ? P - a : 0u; }
FF_HD u32 to_mont(u32 a) { return canon(mont_mul(a, R2)); }
FF_HD u32 from_mont(u32 a) { return canon(mont_mul(a, 1u)); }
// canonical * canonical -> canonical (ff.rs:138-144)
FF_HD u32 mul(u32 a, u32 b) { return canon(mont_mul(canon(mont_mul(a, b)), R2)); }
// x * 2^-1 without a multiply: (x + (x odd ? p : 0)) / 2 ; x < 2^31
FF_HD u32 half(u32 x) { return (x + ((0u - (x & 1u)) & P)) >> 1; }
// Montgomery-form power: base_m = a*R  ->  a^e * R  (ff.rs:200-213 square-and-multiply)
FF_HD u32 mont_pow(u32 base_m, u64 e) {
  u32 r = R1;
  while (e) {
    if (e & 1) r = canon(mont_mul(r, base_m));
    base_m = canon(mont_mul(base_m, base_m));
    e >>= 1;
  }
  return r;
}
// canonical a^e (ff.rs:200-213)
FF_HD u32 pow(u32 a, u64 e) { return from_mont(mont_pow(to_mont(a), e)); }
// Fermat inverse a^(p-2); the reference uses xgcd (ff.rs:169-178, utils.rs:3-13), same value for a != 0
FF_HD u32 inv(u32 a) { return pow(a, (u64)P - 2); }
// reduce an arbitrary u64 (e.g. the raw Fiat-Shamir challenge, fiat_shamir.rs:21-24) to [0, p)
FF_HD u32 reduce64(u64 x) { return (u32)(x % P); }

}  // namespace ff
