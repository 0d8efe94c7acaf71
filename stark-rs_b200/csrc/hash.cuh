// hash.cuh -- the stark-rs 32-byte hash (reference src/hash.rs:7-99) as integer-pipe code.
//
// One hash per thread.  The 32-byte state lives in 32 registers, one byte per register, and the HIGH 24
// bits of each register are allowed to hold garbage: wrapping adds, multiplies and xors only ever
// propagate information upwards, so the low 8 bits stay exact.  A byte is cleaned only where a right
// shift needs it (the two rotations), and there the clean-up is free:
//     rotl8(z, k) = ((z * 0x0101) >> (8 - k)) & 0xff      and   z * 0x0101 = PRMT(z: byte0, byte0, 0, 0)
// i.e. one PRMT both masks the byte and duplicates it, one SHF finishes the rotate (bits >= 8 are garbage
// again, which is fine).  Per mix_state: sbox 3 ops/byte (IMAD, PRMT, SHF), linear layer 6 ops/4 bytes
// (LOP3), the serial neighbour add 1 IADD3/byte; the round-constant add is folded into the NEXT
// consumer's IMAD / IADD3.  The byte chain s[i] += s[i+1] + s[i-1] (hash.rs:77-81) is inherently serial
// (32 dependent adds); occupancy, not ILP, hides it.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define HS_HD __host__ __device__ __forceinline__
#else
#define HS_HD inline
#endif

namespace hs {
typedef uint32_t u32;
typedef uint8_t u8;

// hash.rs:53
#define HS_PRIMES \
  { 2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53 }
// hash.rs:96-99
#define HS_RC                                                                                                 \
  {                                                                                                           \
    0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1b, 0x36, 0x6c, 0xd8, 0xab, 0x4d, 0x9a, 0x2f, 0x5e,     \
        0xbc, 0x63, 0xc6, 0x97, 0x35, 0x6a, 0xd4, 0xb3, 0x7d, 0xfa, 0xef, 0xc5, 0x91, 0x39, 0x72              \
  }

HS_HD u32 prime_at(int i) {
  constexpr u32 t[16] = HS_PRIMES;
  return t[i & 15];
}
HS_HD u32 rc_at(int i) {
  constexpr u32 t[32] = HS_RC;
  return t[i];
}

// low byte of z duplicated into bytes 0 and 1, bytes 2-3 zero  (= (z & 0xff) * 0x0101)
HS_HD u32 dup_byte(u32 z) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(z, 0u, 0x4400);
#else
  return (z & 0xffu) * 0x0101u;
#endif
}
// rotate_left (hash.rs:55-57) of the low byte; result exact in bits 0-7, garbage above
HS_HD u32 rotl_lazy(u32 z, int k) { return dup_byte(z) >> (8 - k); }

struct State {
  u32 s[32];  // low 8 bits exact, high bits garbage; round constants possibly pending (see `pending`)
};

// hash.rs:10-12
HS_HD void init(State &st) {
#pragma unroll
  for (int i = 0; i < 32; i++) st.s[i] = prime_at(i);
}

// mix_state (hash.rs:59-86) WITHOUT its final round-constant add (left pending for the next consumer).
// If PENDING, the round constants of the previous mix are still owed and are folded into the sbox multiply:
// sbox(x + rc) = rotl1(251 x + 251 rc) ^ 0x63.
template <bool PENDING>
HS_HD void mix_lazy(State &st) {
  u32 *s = st.s;
  // sbox (hash.rs:88-94), without the ^0x63 (it cancels or is re-applied in the linear layer below)
#pragma unroll
  for (int i = 0; i < 32; i++) {
    const u32 c = PENDING ? (rc_at(i) * 251u) & 0xffu : 0u;
    s[i] = rotl_lazy(s[i] * 251u + c, 1);
  }
  // linear layer (hash.rs:64-75): out = X ^ t[2,1,3,0], X = t0^t1^t2^t3.  With the pending ^0x63 on all four
  // inputs X is unchanged and each output owes exactly one ^0x63.
#pragma unroll
  for (int g = 0; g < 8; g++) {
    const u32 t0 = s[4 * g], t1 = s[4 * g + 1], t2 = s[4 * g + 2], t3 = s[4 * g + 3];
    const u32 x = t0 ^ t1 ^ t2 ^ t3;
    s[4 * g] = x ^ t2 ^ 0x63u;
    s[4 * g + 1] = x ^ t1 ^ 0x63u;
    s[4 * g + 2] = x ^ t3 ^ 0x63u;
    s[4 * g + 3] = x ^ t0 ^ 0x63u;
  }
  // serial in-place neighbour add (hash.rs:77-81): prev is already updated, next is old; s[31] sees NEW s[0]
  s[0] = s[0] + s[1] + s[31];
#pragma unroll
  for (int i = 1; i < 31; i++) s[i] = s[i] + s[i + 1] + s[i - 1];
  s[31] = s[31] + s[0] + s[30];
}

// absorb one message byte at position i (hash.rs:15-20); no round constants may be pending
HS_HD void absorb_byte(State &st, int i, u32 b) {
  u32 *s = st.s;
  const u32 v = rotl_lazy(s[i] + b, 3);
  s[i] = v;
  s[(i + 7) & 31] ^= v;
}

HS_HD void settle(State &st) {
#pragma unroll
  for (int i = 0; i < 32; i++) st.s[i] += rc_at(i);
}

// 8 finalisation mixes (hash.rs:25-27) + settle.  PENDING as for mix_lazy.  The seven pending-form mixes stay a rolled
// loop: straight-line code was measured 15 % SLOWER in the latency-bound callers (instruction-cache misses).
template <bool PENDING>
HS_HD void finalize(State &st) {
  mix_lazy<PENDING>(st);
#pragma unroll 1
  for (int k = 0; k < 7; k++) mix_lazy<true>(st);
  settle(st);
}

// pack the 32 state bytes into 8 little-endian words
HS_HD void pack_words(const State &st, u32 *w) {
#pragma unroll
  for (int g = 0; g < 8; g++) {
#if defined(__CUDA_ARCH__)
    const u32 lo = __byte_perm(st.s[4 * g], st.s[4 * g + 1], 0x0040);
    const u32 hi = __byte_perm(st.s[4 * g + 2], st.s[4 * g + 3], 0x0040);
    w[g] = __byte_perm(lo, hi, 0x5410);
#else
    w[g] = (st.s[4 * g] & 0xff) | ((st.s[4 * g + 1] & 0xff) << 8) | ((st.s[4 * g + 2] & 0xff) << 16) |
           ((st.s[4 * g + 3] & 0xff) << 24);
#endif
  }
}

// absorb a full 32-byte chunk given as 8 LE words, then mix (hash.rs:14-23).  PENDING: constants owed on entry.
template <bool PENDING>
HS_HD void absorb_words_mix(State &st, const u32 *w) {
  if (PENDING) settle(st);
#pragma unroll
  for (int i = 0; i < 32; i++) absorb_byte(st, i, w[i >> 2] >> (8 * (i & 3)));
  mix_lazy<false>(st);
}

// Hash::combine (hash.rs:41-46): 64-byte message = two chunks, then 8 mixes
// Code size matters as much as instruction count for the latency-bound callers (a narrow tree level runs this code
// once per launch, from a cold instruction cache): the two chunks share ONE copy of absorb + mix and the eight
// finalisation mixes ONE copy of the pending-constants mix.
HS_HD void combine(const u32 *left, const u32 *right, u32 *out) {
  State st;
  init(st);
#pragma unroll 1
  for (int c = 0; c < 2; c++) {
    if (c) settle(st);
    u32 w[8];
#pragma unroll
    for (int g = 0; g < 8; g++) w[g] = c ? right[g] : left[g];
    absorb_words_mix<false>(st, w);
  }
#pragma unroll 1
  for (int k = 0; k < 8; k++) mix_lazy<true>(st);
  settle(st);
  pack_words(st, out);
}

// Hash::from_field_elements(&[v]) (hash.rs:32-35) for a canonical value (< 2^32): 8-byte LE message
HS_HD void leaf1(u32 v, u32 *out) {
  State st;
  init(st);
#pragma unroll
  for (int i = 0; i < 8; i++) absorb_byte(st, i, i < 4 ? (v >> (8 * i)) : 0u);
  mix_lazy<false>(st);
  finalize<true>(st);
  pack_words(st, out);
}

// generic Hash::from_bytes (hash.rs:7-30) -- byte-serial loop form, used for short host/transcript messages
// and for ragged message lengths on the device.
HS_HD void from_bytes(const u8 *msg, size_t n, u8 *out) {
  State st;
  init(st);
  bool pending = false;
  for (size_t off = 0; off < n; off += 32) {
    const int len = (n - off) < 32 ? (int)(n - off) : 32;
    if (pending) settle(st);
#pragma unroll 1
    for (int i = 0; i < len; i++) {
      // runtime-indexed variant of absorb_byte (the State lives in local memory here; not the hot path)
      u32 v = rotl_lazy(st.s[i] + msg[off + i], 3);
      st.s[i] = v;
      st.s[(i + 7) & 31] ^= v;
    }
    mix_lazy<false>(st);
    pending = true;
  }
  if (pending)
    finalize<true>(st);
  else
    finalize<false>(st);
  for (int i = 0; i < 32; i++) out[i] = (u8)st.s[i];
}

}  // namespace hs

// =====================================================================================================
// hs2 -- TWO hashes per thread.  Lane 0 lives in bits 0-7 of every state register and may carry garbage up to bit 23;
// lane 1 lives in bits 24-31 (its carries fall off the top of the register).  ncu on the one-hash-per-thread kernel
// shows the ALU pipe (LOP3/PRMT/SHF/IADD3) at 92 % with the FMA pipe at 10 %: the kernel is bound by ALU-pipe instruction
// count.  Packing two hashes per register halves every per-register instruction, and the formulation below splits the
// work evenly between the two 64-lane/clk pipes:
//   sbox multiply on the raw register:  y = s*251 + c  (one IMAD).  Lane 0 needs no mask: its value stays below 2^15
//                    (clean byte + at most 62 byte sums), so the product stays below 2^24 and never reaches lane 1;
//                    lane 1's product overflows out of the register.  (A 16-bit lane layout needed an AND per register
//                    per mix here: 16 % of all ALU-pipe instructions.)
//   rotl8(z,k) of both lanes:  d = PRMT(y: b0,b0,b3,b3)  (selects the two exact bytes, duplicates them),
//                              e = d << k as IMAD d*2^k,  r = PRMT(e: b1,0,0,b3)  -> clean lanes
//   linear layer: 6 LOP3 per 4 registers, lanes stay clean
//   neighbour chain: s[i] += s[i+1] + s[i-1]' is either one IADD3 (ALU pipe) or two IMAD adds (FMA pipe); the first
//                    CHAIN_ALU bytes take the IADD3 form so that both pipes carry ~117 instructions per mix
// Per hash pair and mix: ~117 ALU-pipe + ~116 FMA-pipe instructions (was 144 + 126 with 16-bit lanes).
namespace hs2 {
using hs::prime_at;
using hs::rc_at;
using hs::u32;

HS_HD u32 prmt(u32 a, u32 b, u32 sel) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(a, b, sel);
#else
  const uint64_t src = ((uint64_t)b << 32) | a;
  u32 r = 0;
  for (int i = 0; i < 4; i++) r |= (u32)((src >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
#endif
}
// a + b on the FMA pipe: a * one + b with `one` a RUNTIME 1 (ptxas folds a literal 1 back into IADD3, which
// lands on the ALU pipe)
HS_HD u32 add_fma(u32 a, u32 b, u32 one) { return a * one + b; }
HS_HD u32 xor3(u32 a, u32 b, u32 c) {
#if defined(__CUDA_ARCH__)
  u32 d;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
#else
  return a ^ b ^ c;
#endif
}
constexpr u32 BOTH = 0x01000001u;   // x * BOTH = the byte x in lane 0 (bits 0-7) and lane 1 (bits 24-31)
constexpr u32 M2 = 0xff0000ffu;
constexpr int CHAIN_ALU = 5;        // neighbour-chain bytes done as one IADD3 instead of two IMAD adds (pipe balance)
// rotl8 of both lanes by k (lane 0 possibly dirty above bit 7); result lanes clean.  `pow2k` = 2^k as a RUNTIME value
// (State2::one << k): with a literal the compiler strength-reduces the multiply to a shift/add on the ALU pipe, which is
// the saturated one (ncu: ALU 92 %, FMA 38 % before this change).
HS_HD u32 rotl2(u32 y, u32 pow2k) { return prmt(prmt(y, 0u, 0x3300) * pow2k, 0u, 0x3441); }

struct State2 {
  u32 s[32];
  u32 one, two, eight;  // runtime constants 1, 2, 8 (see add_fma, rotl2)
};
HS_HD void init(State2 &st, u32 one) {
  st.one = one, st.two = one << 1, st.eight = one << 3;
#pragma unroll
  for (int i = 0; i < 32; i++) st.s[i] = prime_at(i) * BOTH;
}
// mix_state (hash.rs:59-86) for both lanes, round constants left pending (see hs::mix_lazy)
template <bool PENDING>
HS_HD void mix_lazy(State2 &st) {
  u32 *s = st.s;
#pragma unroll
  for (int i = 0; i < 32; i++) {
    const u32 c = PENDING ? ((rc_at(i) * 251u) & 0xffu) * BOTH : 0u;
    s[i] = rotl2(s[i] * 251u + c, st.two);   // lane 0 < 2^15 here, so lane 0's product stays below bit 24
  }
#pragma unroll
  for (int g = 0; g < 8; g++) {
    // 6 LOP3 per group, pinned with lop3.b32 (left to itself the compiler emits 8): x = t0^t1^t2^t3 in two, then
    // each output = x ^ t_j ^ 0x63 in one three-input XOR
    const u32 t0 = s[4 * g], t1 = s[4 * g + 1], t2 = s[4 * g + 2], t3 = s[4 * g + 3];
    const u32 x = xor3(t0, t1, t2) ^ t3;
    s[4 * g] = xor3(x, t2, 0x63u * BOTH);
    s[4 * g + 1] = xor3(x, t1, 0x63u * BOTH);
    s[4 * g + 2] = xor3(x, t3, 0x63u * BOTH);
    s[4 * g + 3] = xor3(x, t0, 0x63u * BOTH);
  }
  // s[i] += s[i+1] + s[i-1]' (hash.rs:77-81).  Lane 0 grows to at most 255 + 31*510 + 1020 < 2^15, lane 1 wraps.
  s[0] = s[0] + s[1] + s[31];
#pragma unroll
  for (int i = 1; i < CHAIN_ALU; i++) s[i] = s[i] + s[i + 1] + s[i - 1];
  u32 t[32];
#pragma unroll
  for (int i = CHAIN_ALU; i < 31; i++) t[i] = add_fma(s[i], s[i + 1], st.one);   // independent pair sums (FMA pipe)
#pragma unroll
  for (int i = CHAIN_ALU; i < 31; i++) s[i] = add_fma(t[i], s[i - 1], st.one);    // the serial running add
  s[31] = s[31] + s[0] + s[30];
}
HS_HD void settle(State2 &st) {
#pragma unroll
  for (int i = 0; i < 32; i++) st.s[i] += rc_at(i) * BOTH;
}
template <bool PENDING>
HS_HD void finalize(State2 &st) {
  mix_lazy<PENDING>(st);
#pragma unroll 1
  for (int k = 0; k < 7; k++) mix_lazy<true>(st);
  settle(st);
}
// absorb the byte pair b (clean lanes) at position i (hash.rs:15-20); constants must be settled
HS_HD void absorb_pair(State2 &st, int i, u32 b) {
  const u32 v = rotl2(add_fma(st.s[i], b, st.one), st.eight);
  st.s[i] = v;
  st.s[(i + 7) & 31] ^= v;
}
// byte k (0..3) of word a -> lane 0, of word b -> lane 1, clean
HS_HD u32 pair_bytes(u32 a, u32 b, int k) { return prmt(a, b, ((4u + (u32)k) << 12) | (u32)k) & M2; }

template <bool PENDING>
HS_HD void absorb_words_mix(State2 &st, const u32 *wa, const u32 *wb) {
  if (PENDING) settle(st);
#pragma unroll
  for (int i = 0; i < 32; i++) absorb_pair(st, i, pair_bytes(wa[i >> 2], wb[i >> 2], i & 3));
  mix_lazy<false>(st);
}
HS_HD void pack_words(const State2 &st, u32 *wa, u32 *wb) {
#pragma unroll
  for (int g = 0; g < 8; g++) {
    const u32 *s = st.s + 4 * g;
    wa[g] = prmt(prmt(s[0], s[1], 0x0040), prmt(s[2], s[3], 0x0040), 0x5410);
    wb[g] = prmt(prmt(s[0], s[1], 0x0073), prmt(s[2], s[3], 0x0073), 0x5410);
  }
}
// two Hash::combine (hash.rs:41-46) at once
HS_HD void combine2(const u32 *la, const u32 *ra, const u32 *lb, const u32 *rb, u32 *oa, u32 *ob, u32 one) {
  State2 st;
  init(st, one);
#pragma unroll 1
  for (int c = 0; c < 2; c++) {   // one copy of absorb + mix for both chunks (code size, see hs::combine)
    if (c) settle(st);
    u32 wa[8], wb[8];
#pragma unroll
    for (int g = 0; g < 8; g++) wa[g] = c ? ra[g] : la[g], wb[g] = c ? rb[g] : lb[g];
    absorb_words_mix<false>(st, wa, wb);
  }
#pragma unroll 1
  for (int k = 0; k < 8; k++) mix_lazy<true>(st);
  settle(st);
  pack_words(st, oa, ob);
}
// two Hash::from_field_elements(&[v]) (hash.rs:32-35) at once
HS_HD void leaf2(u32 va, u32 vb, u32 *oa, u32 *ob, u32 one) {
  State2 st;
  init(st, one);
#pragma unroll
  for (int i = 0; i < 8; i++) absorb_pair(st, i, i < 4 ? pair_bytes(va, vb, i) : 0u);
  mix_lazy<false>(st);
  finalize<true>(st);
  pack_words(st, oa, ob);
}
}  // namespace hs2

// =====================================================================================================
// hsq -- ONE hash on FOUR adjacent lanes (lane q = lane & 3 holds state bytes 8q .. 8q+7), for the narrow tree levels
// and the FRI tail, where there are fewer hashes than lanes and the only thing that matters is the LATENCY of one hash.
// A lone warp runs the one-hash-per-thread form (2.3 k dependent-ish instructions) in ~2.7 us; split over four lanes
// every lane executes ~1.1 k instructions and the per-mix critical path is three shuffle hops + an 8-long add chain.
//   mix_state (hash.rs:59-86): sbox and the 4-byte linear groups are lane-local.  The neighbour-add chain is the closed
//   form  s'[i] = s[31] + sum_{k<=i} (s[k] + s[k+1])  (i <= 30),  s'[31] = s[31] + s'[0] + s'[30]:
//   a local 8-element prefix, the three lower lanes' totals fetched with three independent shuffles, one more hop for s'[0].
//   absorb (hash.rs:15-20): byte i xors into byte i+7, i.e. lane q's byte j depends on lane q-1's byte j+1 (and byte 7 on
//   the lane's own byte 0), so the four lanes absorb in four phases; lane 3's bytes 1..7 wrap into lane 0's bytes 0..6.
// Bytes are lazy as in hs (garbage above bit 7).  All 32 lanes of a warp must call these functions together.
namespace hsq {
using hs::u32;

// Like hso, written over a lane-group interface W { u32 q; u32 shfl(u32 v, u32 src_lane_of_the_quad); } so that
// tests/emul/hash_emul.cpp can run the four lanes as four host threads against the oracle.
struct Quad {
  u32 s[8];
  u32 rc[8];      // round constants of this lane's bytes
  u32 rc251[8];   // (rc * 251) & 0xff: the pending constants folded into the next sbox multiply (see hs::mix_lazy)
};

// lane-dependent constants without lane-indexed constant-memory loads (those serialise per distinct address):
// the four lanes' byte rows are packed into words and picked with selects
HS_HD u32 pick4(u32 q, u32 a, u32 b, u32 c, u32 d) { return q < 2 ? (q == 0 ? a : b) : (q == 2 ? c : d); }
template <class W>
HS_HD void init(const W &w, Quad &st) {
  const u32 q = w.q;
  constexpr u32 pr[16] = HS_PRIMES;
  constexpr u32 rc[32] = HS_RC;
#pragma unroll
  for (int w = 0; w < 2; w++) {
    // bytes 4w .. 4w+3 of each lane's row, packed little-endian
    u32 pk[4], rk[4], mk[4];
#pragma unroll
    for (int l = 0; l < 4; l++) {
      pk[l] = rk[l] = mk[l] = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int i = 8 * l + 4 * w + b;
        pk[l] |= pr[i & 15] << (8 * b);
        rk[l] |= rc[i] << (8 * b);
        mk[l] |= ((rc[i] * 251u) & 0xffu) << (8 * b);
      }
    }
    const u32 pw = (q & 1u) ? pk[1] : pk[0];   // PRIMES repeat with period 16: lanes 0,2 / 1,3
    const u32 rw = pick4(q, rk[0], rk[1], rk[2], rk[3]);
    const u32 mw = pick4(q, mk[0], mk[1], mk[2], mk[3]);
#pragma unroll
    for (int b = 0; b < 4; b++) {
      st.s[4 * w + b] = (pw >> (8 * b)) & 0xffu;
      st.rc[4 * w + b] = (rw >> (8 * b)) & 0xffu;
      st.rc251[4 * w + b] = (mw >> (8 * b)) & 0xffu;
    }
  }
}
HS_HD u32 mad2(u32 x, u32 c) {   // 2 x + c, opaque to the re-association passes
#if defined(__CUDA_ARCH__)
  u32 r;
  asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(r) : "r"(x), "r"(c));
  return r;
#else
  return 2u * x + c;
#endif
}
// mix_state without its final round-constant add (left pending, as hs::mix_lazy)
template <bool PENDING, class W>
HS_HD void mix_lazy(const W &w, Quad &st) {
  u32 *s = st.s;
#pragma unroll
  for (int j = 0; j < 8; j++) s[j] = hs::rotl_lazy(s[j] * 251u + (PENDING ? st.rc251[j] : 0u), 1);
#pragma unroll
  for (int g = 0; g < 2; g++) {
    const u32 t0 = s[4 * g], t1 = s[4 * g + 1], t2 = s[4 * g + 2], t3 = s[4 * g + 3];
    const u32 x = t0 ^ t1 ^ t2 ^ t3;
    s[4 * g] = x ^ t2 ^ 0x63u;
    s[4 * g + 1] = x ^ t1 ^ 0x63u;
    s[4 * g + 2] = x ^ t3 ^ 0x63u;
    s[4 * g + 3] = x ^ t0 ^ 0x63u;
  }
  // neighbour add, closed form.  With t[k] = s[k] + s[k+1]:  s'[i] = s[31] + sum_{k<=i} t[k]  (i <= 30).
  // A lane's total telescopes: sum_{k=8q}^{8q+7} t[k] = 2 T_q - s[8q] + s[8q+8] with T_q the lane's byte sum, so the
  // offset of lane q is  O_q = s[31] + 2 (T_0 + .. + T_{q-1}) - s[0] + s[8q]  and needs only T of the lower lanes,
  // s[0] and s[31]: ONE hop of independent shuffles per mix (the local prefix runs underneath it).
  // As in hso::mix_lazy the gather is arranged to leave as little as possible behind the last shuffle: lane q reads T from
  // lanes q-1 .. q-3 and, where that runs off the quad, from ITSELF (no selects; the 3 - q surplus copies of T_q come out
  // of the early constant), and s[31], s[0], the in-lane prefixes and the surplus are folded into C_j while the shuffles
  // are in flight: s'[j] = C_j + 2 (a1 + a2 + a3).
  const u32 q = w.q;
  const u32 e0 = s[0] + s[1];
  const u32 T = (s[0] + s[1] + s[2]) + (s[3] + s[4] + s[5]) + (s[6] + s[7]);
  const u32 s31 = w.shfl(s[7], 3u);                                       // old s[31]
  const u32 z0 = w.shfl(s[0], 0u);                                             // s[0]
  const u32 n0 = s31 + w.shfl(e0, 0u);                                         // s'[0] = s[31] + s[0] + s[1]
  const u32 nxt0 = w.shfl(s[0], (q + 1u) & 3u);                         // s[8q+8] (lanes 0..2)
  const u32 a1 = w.shfl(T, q >= 1u ? q - 1u : q);
  const u32 a2 = w.shfl(T, q >= 2u ? q - 2u : q);
  const u32 a3 = w.shfl(T, q >= 3u ? q - 3u : q);
  u32 P[7];
  P[0] = e0;
#pragma unroll
  for (int j = 1; j < 7; j++) P[j] = P[j - 1] + (s[j] + s[j + 1]);
  const u32 c = s31 - z0 + s[0] - 2u * (3u - q) * T;
  const u32 c6 = c + P[6];
  const u32 clast = (q == 3u) ? (s31 + n0 + c6) : (c6 + s[7] + nxt0);   // lane 3: s'[31] = s[31] + s'[0] + s'[30]
  // 2 x + C_j as ONE multiply-add behind the three-input add (plain C is re-associated into x + x, + c, + P_j: two adds
  // deeper behind the last shuffle)
  const u32 x = a1 + a2 + a3;
#pragma unroll
  for (int j = 0; j < 7; j++) s[j] = mad2(x, c + P[j]);
  s[7] = mad2(x, clast);
}
HS_HD void settle(Quad &st) {
#pragma unroll
  for (int j = 0; j < 8; j++) st.s[j] += st.rc[j];
}
template <bool PENDING, class W>
HS_HD void finalize(const W &w, Quad &st) {
  mix_lazy<PENDING>(w, st);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int k = 0; k < 7; k++) mix_lazy<true>(w, st);
  settle(st);
}
// absorb a full 32-byte chunk m[0..8) (little-endian words; EVERY lane holds the whole chunk), then mix.
// The absorb loop (hash.rs:15-20: byte i xors into byte i+7) is a serial pass over all 32 state bytes, cheap in one
// thread (7 independent chains of <= 5 links, ~130 cycles) and slow across lanes (each link would be a shuffle hop:
// measured 560-620 cycles per chunk for lane-phased and Jacobi-sweep variants).  So the quad all-gathers its state
// (two packed words per lane, 8 shuffles), every lane runs the same serial absorb on the full state, and keeps its own
// eight bytes.
template <bool PENDING, class W>
HS_HD void absorb_mix(const W &w, Quad &st, const u32 *m) {
  if (PENDING) settle(st);
  u32 *s = st.s;
  const u32 w0 = hs2::prmt(hs2::prmt(s[0], s[1], 0x0040), hs2::prmt(s[2], s[3], 0x0040), 0x5410);
  const u32 w1 = hs2::prmt(hs2::prmt(s[4], s[5], 0x0040), hs2::prmt(s[6], s[7], 0x0040), 0x5410);
  u32 S[32];
#pragma unroll
  for (int l = 0; l < 4; l++) {
    const u32 a = w.shfl(w0, (u32)l), b = w.shfl(w1, (u32)l);
#pragma unroll
    for (int j = 0; j < 4; j++) S[8 * l + j] = a >> (8 * j), S[8 * l + 4 + j] = b >> (8 * j);   // lazy bytes
  }
#pragma unroll
  for (int i = 0; i < 32; i++) {
    const u32 v = hs::rotl_lazy(S[i] + (m[i >> 2] >> (8 * (i & 3))), 3);
    S[i] = v;
    S[(i + 7) & 31] ^= v;
  }
#pragma unroll
  for (int j = 0; j < 8; j++) s[j] = pick4(w.q, S[j], S[8 + j], S[16 + j], S[24 + j]);
  mix_lazy<false>(w, st);
}
// Hash::combine (hash.rs:41-46) of the 32-byte hashes at `left` and `right`; every lane of the quad returns its own
// 8 output bytes as two little-endian words (lane q: bytes 8q .. 8q+7)
template <class W>
HS_HD void combine(const W &w, const uint8_t *left, const uint8_t *right, u32 &o0, u32 &o1) {
  Quad st;
  init(w, st);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int c = 0; c < 2; c++) {   // one copy of absorb + mix for both chunks (code size, see hs::combine)
    if (c) settle(st);
#if defined(__CUDA_ARCH__)
    const uint4 x = reinterpret_cast<const uint4 *>(c ? right : left)[0], y = reinterpret_cast<const uint4 *>(c ? right : left)[1];
    const u32 m[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
#else
    u32 m[8];
    for (int i = 0; i < 8; i++) {
      const uint8_t *b = (c ? right : left) + 4 * i;
      m[i] = (u32)b[0] | ((u32)b[1] << 8) | ((u32)b[2] << 16) | ((u32)b[3] << 24);
    }
#endif
    absorb_mix<false>(w, st, m);
  }
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int k = 0; k < 8; k++) mix_lazy<true>(w, st);
  settle(st);
  const u32 *s = st.s;
  o0 = hs2::prmt(hs2::prmt(s[0], s[1], 0x0040), hs2::prmt(s[2], s[3], 0x0040), 0x5410);
  o1 = hs2::prmt(hs2::prmt(s[4], s[5], 0x0040), hs2::prmt(s[6], s[7], 0x0040), 0x5410);
}
#if defined(__CUDACC__)
// the lane group on the device: four adjacent lanes of a warp (ALL 32 lanes of the warp must execute the calls)
struct Dev {
  u32 q, base;
  __device__ __forceinline__ Dev() {
    const u32 lane = threadIdx.x & 31u;
    q = lane & 3u, base = lane & ~3u;
  }
  __device__ __forceinline__ u32 shfl(u32 v, u32 src) const { return __shfl_sync(0xffffffffu, v, base + src); }
};
__device__ __forceinline__ void combine(const uint8_t *left, const uint8_t *right, u32 &o0, u32 &o1) {
  const Dev w;
  combine(w, left, right, o0, o1);
}
#endif
}  // namespace hsq

// =====================================================================================================
// hso -- ONE hash on EIGHT adjacent lanes (lane q holds state bytes 4q .. 4q+3), for the steps where the latency of one
// hash is all that matters and most lanes of the CTA are idle anyway: the narrow levels of a Merkle climb (<= NT/8
// parents), the Fiat-Shamir round.  Against hsq (four lanes, 2560 cycles per Hash::combine for a lone warp, which issues
// its ~85 instructions per mix at ~2 cycles each -- dependency-bound, not issue-bound) it measures 2180 cycles
// (stark_bench_hash_latency_hso): the ten mixes stay a chain of sbox -> linear layer -> byte totals -> one shuffle hop ->
// prefix, ~140 cycles each however few bytes a lane holds; the gain is the absorb:
//   * mix_state: half the bytes per lane.  sbox and the 4-byte linear group are lane-local; the neighbour add is the
//     closed form of hsq (local prefix + 2 x the byte totals of the lower lanes - s[0] + s[4q]) with the seven lower
//     totals fetched by independent shuffles, ONE hop.
//   * absorb (hash.rs:15-20: byte i xors into byte i+7): hsq gathers the whole state into every lane and runs the 32
//     serial steps redundantly (390 cycles).  The 32 steps are SEVEN independent chains (i mod 7) of at most five links:
//     lane c takes chain c -- gather its five bytes (5 shuffles), five dependent links, the wrap-around of the last link
//     into byte c+3 (1 shuffle), scatter back to the byte-block layout (8 shuffles).
// The code is host+device inline over a lane-group interface W (W::q, W::shfl) so that tests/emul/hash_emul.cpp runs it
// on eight host threads against the oracle; on the device W is hso::Dev (shuffles within the warp).
namespace hso {
using hs::u32;

struct Oct {
  u32 s[4];
  u32 rc[4];      // round constants of this lane's bytes
  u32 rc251[4];   // (rc * 251) & 0xff, folded into the next sbox multiply when the constants are still pending
};
HS_HD u32 pack4(u32 a, u32 b, u32 c, u32 d) {   // the four LOW bytes, little-endian
  return hs2::prmt(hs2::prmt(a, b, 0x0040), hs2::prmt(c, d, 0x0040), 0x5410);
}
template <class W>
HS_HD void init(const W &w, Oct &st) {
  constexpr u32 pr[16] = HS_PRIMES;
  constexpr u32 rc[32] = HS_RC;
  const u32 q = w.q;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    // lane-dependent constants by selects (a lane-indexed constant load serialises per distinct address)
    u32 p = 0, r = 0;
#pragma unroll
    for (int l = 0; l < 8; l++) {
      if (q == (u32)l) p = pr[(4 * l + j) & 15], r = rc[4 * l + j];
    }
    st.s[j] = p, st.rc[j] = r, st.rc251[j] = (r * 251u) & 0xffu;
  }
}
// mix_state (hash.rs:59-86) without its final round-constant add (left pending, as hs::mix_lazy)
template <bool PENDING, class W>
HS_HD void mix_lazy(const W &w, Oct &st) {
  u32 *s = st.s;
  const u32 q = w.q;
#pragma unroll
  for (int j = 0; j < 4; j++) s[j] = hs::rotl_lazy(s[j] * 251u + (PENDING ? st.rc251[j] : 0u), 1);
  {
    const u32 t0 = s[0], t1 = s[1], t2 = s[2], t3 = s[3];
    const u32 x = t0 ^ t1 ^ t2 ^ t3;
    s[0] = x ^ t2 ^ 0x63u, s[1] = x ^ t1 ^ 0x63u, s[2] = x ^ t3 ^ 0x63u, s[3] = x ^ t0 ^ 0x63u;
  }
  // neighbour add, closed form: with t[k] = s[k] + s[k+1], s'[i] = s[31] + sum_{k<=i} t[k] (i <= 30).  A lane's total of t
  // telescopes to 2 T_q - s[4q] + s[4q+4] (T_q = its byte sum), so lane q's offset is
  //   O_q = s[31] + 2 (T_0 + .. + T_{q-1}) - s[0] + s[4q]
  // The gather of the D_m = 2 T_m is the critical path of the whole hash (a lone warp waits for seven shuffles and then
  // for everything that depends on them), so it is arranged to leave ONE add behind the last shuffle:
  //  * lane q reads D from lanes q-1 .. q-7 and, where that runs off the octet, from ITSELF -- no selects after the
  //    shuffles; the (7 - q) surplus copies of D_q are taken out of the early constant instead;
  //  * everything else (s[31], s[0], the in-lane prefixes P_j, the surplus) is folded into C_j while the shuffles are in
  //    flight, and so are the first three arrivals; out_j = (C_j + a1 + a2 + a3) + (a4 + a5 + a6) + a7.
  const u32 e0 = s[0] + s[1];
  // ONE three-input add, not a child of e0: the compiler re-associates plain C into the chain (s0 + s1) + s2, one add
  // deeper on the critical path
#if defined(__CUDA_ARCH__)
  u32 g3;
  asm("{\n\t.reg .u32 t;\n\tadd.u32 t, %1, %3;\n\tadd.u32 %0, t, %2;\n\t}" : "=r"(g3) : "r"(s[0]), "r"(s[1]), "r"(s[2]));
#else
  const u32 g3 = (s[0] + s[2]) + s[1];
#endif
  const u32 D = g3 + g3 + (s[3] + s[3]);     // 2 T_q, two adds deep like T itself
  const u32 s31 = w.shfl(s[3], 7u);          // old s[31]
  const u32 z0 = w.shfl(s[0], 0u);           // s[0]
  const u32 n0 = s31 + w.shfl(e0, 0u);       // s'[0] = s[31] + s[0] + s[1]
  const u32 nxt0 = w.shfl(s[0], (q + 1u) & 7u);   // s[4q+4]
  u32 a[7];
#pragma unroll
  for (int d = 1; d <= 7; d++) a[d - 1] = w.shfl(D, q >= (u32)d ? q - (u32)d : q);
  const u32 P0 = e0, P1 = P0 + (s[1] + s[2]), P2 = P1 + (s[2] + s[3]);
  const u32 base = s31 - z0 + s[0] - (7u - q) * D;
  const u32 C0 = base + P0, C1 = base + P1, C2 = base + P2;
  // lane 7: s'[31] = s[31] + s'[0] + s'[30]
  const u32 C3 = q == 7u ? C2 + s31 + n0 : C2 + s[3] + nxt0;
  const u32 x = a[0] + a[1] + a[2], y = a[3] + a[4] + a[5];
  s[0] = (C0 + x) + y + a[6], s[1] = (C1 + x) + y + a[6], s[2] = (C2 + x) + y + a[6], s[3] = (C3 + x) + y + a[6];
}
template <class W>
HS_HD void settle(const W &, Oct &st) {
#pragma unroll
  for (int j = 0; j < 4; j++) st.s[j] += st.rc[j];
}
// absorb the 32-byte chunk at `msg` (hash.rs:14-21), then mix.  PENDING: round constants owed on entry.
template <bool PENDING, class W>
HS_HD void absorb_mix(const W &w, Oct &st, const uint8_t *msg) {
  if (PENDING) settle(w, st);
  u32 *s = st.s;
  const u32 c = w.q;   // this lane's chain: positions c, c+7, .. (lane 7 computes an unused duplicate of a shifted chain)
  const u32 word = pack4(s[0], s[1], s[2], s[3]);
  u32 v[5];
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const u32 p = (c + 7u * (u32)j) & 31u;                 // positions past 31 (chains 4..6, j = 4) are computed and ignored
    const u32 x = w.shfl(word, p >> 2) >> (8u * (p & 3u));   // state byte p (lazy: garbage above bit 7)
    const u32 in = j ? x ^ v[j - 1] : x;                   // byte p was xored with v[p-7] by the link before
    v[j] = hs::rotl_lazy(in + msg[p], 3);
  }
  // the links at positions 25..31 xor into positions 0..6 (after those were absorbed): position t receives v[t+25], the
  // LAST element of chain (t+4) mod 7
  const u32 lastv = c <= 3u ? v[4] : v[3];
  const u32 src = c + 4u >= 7u ? c + 4u - 7u : c + 4u;
  v[0] ^= w.shfl(lastv, c == 7u ? 0u : src);
  // back to the byte-block layout: byte p = 4q + j is element p / 7 of chain p mod 7
  const u32 A = pack4(v[0], v[1], v[2], v[3]), B = v[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const u32 p = 4u * w.q + (u32)j, idx = (p * 37u) >> 8, from = p - 7u * idx;
    const u32 a = w.shfl(A, from), b = w.shfl(B, from);
    s[j] = idx == 4u ? b : a >> (8u * idx);
  }
  mix_lazy<false>(w, st);
}
// Hash::combine (hash.rs:41-46) of the 32-byte hashes at `left` and `right`; lane q returns bytes 4q .. 4q+3 as one
// little-endian word
template <class W>
HS_HD u32 combine(const W &w, const uint8_t *left, const uint8_t *right) {
  Oct st;
  init(w, st);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int c = 0; c < 2; c++) {   // one copy of absorb + mix for both chunks (code size, see hs::combine)
    if (c) settle(w, st);
    absorb_mix<false>(w, st, c ? right : left);
  }
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int k = 0; k < 8; k++) mix_lazy<true>(w, st);
  settle(w, st);
  return pack4(st.s[0], st.s[1], st.s[2], st.s[3]);
}
#if defined(__CUDACC__)
// the lane group on the device: eight adjacent lanes of a warp (ALL 32 lanes of the warp must execute the calls)
struct Dev {
  u32 q, base;
  __device__ __forceinline__ Dev() {
    const u32 lane = threadIdx.x & 31u;
    q = lane & 7u, base = lane & ~7u;
  }
  __device__ __forceinline__ u32 shfl(u32 v, u32 src) const { return __shfl_sync(0xffffffffu, v, base + src); }
};
#endif
}  // namespace hso
