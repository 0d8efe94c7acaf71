// transcript.cuh -- the Fiat-Shamir transcript on the device (reference src/fiat_shamir.rs:4-25), shared by merkle.cu
// (the kernel that produces a Merkle root also absorbs it and draws alpha, so a FRI round needs no separate transcript
// launch) and fri.cu.
//
// Streaming form of Hash::from_bytes over an append-only transcript: `s` is the sponge after all complete 32-byte
// chunks (each followed by its mix, round constants settled), `pend` the bytes of the incomplete last chunk.
#pragma once
#include "field.cuh"
#include "hash.cuh"

struct TranscriptDev {
  uint32_t s[32];
  uint8_t pend[32];
  uint32_t npend;
};

HS_HD void tr_init(TranscriptDev &T) {
  for (int i = 0; i < 32; i++) T.s[i] = hs::prime_at(i), T.pend[i] = 0;
  T.npend = 0;
}
HS_HD void tr_absorb(TranscriptDev &T, const uint8_t *data, size_t n) {
  for (size_t k = 0; k < n; k++) {
    T.pend[T.npend++] = data[k];
    if (T.npend == 32) {
      hs::State st;
      for (int i = 0; i < 32; i++) st.s[i] = T.s[i];
      for (int i = 0; i < 32; i++) hs::absorb_byte(st, i, T.pend[i]);
      hs::mix_lazy<false>(st);
      hs::settle(st);
      for (int i = 0; i < 32; i++) T.s[i] = st.s[i] & 0xffu;
      T.npend = 0;
    }
  }
}
// FiatShamir::challenge (fiat_shamir.rs:19-25): first 8 bytes of Hash(transcript), little-endian, UNREDUCED
HS_HD uint64_t tr_challenge(const TranscriptDev &T) {
  hs::State st;
  for (int i = 0; i < 32; i++) st.s[i] = T.s[i];
  if (T.npend) {
    for (uint32_t i = 0; i < T.npend; i++) {
      const uint32_t v = hs::rotl_lazy(st.s[i] + T.pend[i], 3);
      st.s[i] = v;
      st.s[(i + 7) & 31] ^= v;
    }
    hs::mix_lazy<false>(st);
    hs::finalize<true>(st);
  } else {
    hs::finalize<false>(st);
  }
  uint64_t v = 0;
  for (int b = 0; b < 8; b++) v |= (uint64_t)(st.s[b] & 0xffu) << (8 * b);
  return v;
}

// what a root-producing kernel does with the root (fri.rs:129-138): copy it to roots_out, absorb it, and unless this
// is the last round draw alpha (raw u64 and mod-p Montgomery form).  T == nullptr: nothing.
struct TranscriptArgs {
  TranscriptDev *T;
  uint8_t *root_out;      // 32 bytes
  int draw;
  uint64_t *alpha_raw;
  uint32_t *alpha_m;
};

#if defined(__CUDACC__)
// generic (transcript length not a multiple of 32) path: rare, kept out of line so it does not bloat the callers
static __device__ __noinline__ uint64_t transcript_round_slow(TranscriptDev *T, const uint32_t *root, int draw) {
  uint8_t r[32];
  for (int i = 0; i < 32; i++) r[i] = (uint8_t)(root[i >> 2] >> (8 * (i & 3)));
  TranscriptDev t = *T;
  tr_absorb(t, r, 32);
  *T = t;
  return draw ? tr_challenge(t) : 0;
}
// The same round on ONE WARP (all 32 lanes must call it; every quad computes the same thing, lane 0 writes): the
// 4-lanes-per-hash form (hsq) is ~2x shorter in latency than one thread, and this round sits on the critical path of
// every FRI round.  root_bytes: the 32-byte root (shared or global memory).  Falls back to the one-thread form when
// the transcript is not chunk-aligned.
__device__ __forceinline__ void transcript_round(const TranscriptArgs &A, const uint32_t *root);
__device__ __forceinline__ void transcript_round_warp(const TranscriptArgs &A, const uint8_t *root_bytes) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint4 x = reinterpret_cast<const uint4 *>(root_bytes)[0], y = reinterpret_cast<const uint4 *>(root_bytes)[1];
  const uint32_t m[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
  TranscriptDev *T = A.T;
  if (T->npend != 0) {      // warp-uniform
    if (lane == 0) transcript_round(A, m);
    return;
  }
  if (lane < 8) reinterpret_cast<uint32_t *>(A.root_out)[lane] = m[lane];
  // one hash on eight lanes (hso); the four octets of the warp compute the same thing
  const hso::Dev w;
  hso::Oct st;
  hso::init(w, st);
#pragma unroll
  for (int j = 0; j < 4; j++) st.s[j] = T->s[4 * w.q + j];
  __syncwarp();             // every lane has read the sponge before lane group 0 overwrites it
  hso::absorb_mix<false>(w, st, root_bytes);          // round constants of this mix pending
  if (lane < 8) {
#pragma unroll
    for (int j = 0; j < 4; j++) T->s[4 * w.q + j] = (st.s[j] + st.rc[j]) & 0xffu;
  }
  if (!A.draw) return;
#pragma unroll 1
  for (int k = 0; k < 8; k++) hso::mix_lazy<true>(w, st);
  // FiatShamir::challenge: the first 8 bytes, little-endian (lanes 0 and 1 of an octet)
  const uint32_t word = hso::pack4(st.s[0] + st.rc[0], st.s[1] + st.rc[1], st.s[2] + st.rc[2], st.s[3] + st.rc[3]);
  const uint32_t lo = w.shfl(word, 0u), hi = w.shfl(word, 1u);
  if (lane == 0) {
    const uint64_t a = ((uint64_t)hi << 32) | lo;
    *A.alpha_raw = a;
    *A.alpha_m = ff::to_mont(ff::reduce64(a));
  }
}

// One thread.  root = 8 little-endian words.  Fast path (transcript length a multiple of 32, the FRI case): the sponge
// stays in registers with static indexing, one copy of each mix form (code size: this runs once per launch).
__device__ __forceinline__ void transcript_round(const TranscriptArgs &A, const uint32_t *root) {
#pragma unroll
  for (int g = 0; g < 8; g++) reinterpret_cast<uint32_t *>(A.root_out)[g] = root[g];
  TranscriptDev *T = A.T;
  uint64_t a;
  if (T->npend == 0) {
    hs::State st;
#pragma unroll
    for (int i = 0; i < 32; i++) st.s[i] = T->s[i];
    hs::absorb_words_mix<false>(st, root);   // round constants of this mix pending
#pragma unroll
    for (int i = 0; i < 32; i++) T->s[i] = (st.s[i] + hs::rc_at(i)) & 0xffu;
    if (!A.draw) return;
    // Hash(transcript) = 8 finalisation mixes of the settled sponge = 8 pending-form mixes of the unsettled one
#pragma unroll 1
    for (int k = 0; k < 8; k++) hs::mix_lazy<true>(st);
    a = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) a |= (uint64_t)((st.s[b] + hs::rc_at(b)) & 0xffu) << (8 * b);
  } else {
    a = transcript_round_slow(T, root, A.draw);
    if (!A.draw) return;
  }
  *A.alpha_raw = a;
  *A.alpha_m = ff::to_mont(ff::reduce64(a));
}
#endif
