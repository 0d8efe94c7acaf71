// ntt_core.cuh -- shared NTT building blocks (host+device inline) and the single-CTA transform for 8 <= N <= 4096.
//
// Replaces Polynomial::eval_domain / interpolate_domain on structured domains (reference
// src/univariate/eval.rs:16-21 -- O(n*m) Horner; interpolate.rs:6-44 -- O(n^3) Lagrange) by an
// O(n log n) transform with NATURAL-order input and output (the Merkle leaf index and the FRI
// (i, i+N/2) pairing both assume natural order, fri.rs:57-91, 118-127).  N >= 2^13 is ntt_pass.cuh.
//
// One CTA of N/8 threads transforms one array of N elements: radix-8 Stockham DIF rounds (first round radix 2 or 4
// when log2 N is not a multiple of 3) with every thread holding the 8 inputs of one butterfly in registers.  The first
// round loads straight from HBM and the last round stores straight to HBM; between rounds the data moves through
// shared memory.
//
// Stockham DIF round with s = product of previous radices, R = this radix, u = task in [0, L/R):
//     q = u mod s, p = u div s
//     inputs   l_in(j)  = u + j*(L/R)                         j < R
//     outputs  l_out(k) = q + s*(R*p + k)                      k < R
//     out_k = (sum_j in_j * w_R^(jk)) * w_L^(s*p*k)
// Everything is kept lazily in [0, 2p) between rounds (field.cuh).
#pragma once
#include "field.cuh"

namespace ntt {
using ff::u32;
using ff::u64;

struct alignas(16) q4 {
  u32 x, y, z, w;
};

// a constant multiplier in Shoup form (field.cuh shoup_mul): w plain, s = floor(w * 2^32 / p)
struct alignas(8) wpair {
  u32 w, s;
};

// w23^e tables (Montgomery form): lo[i] = w23^i (i < 4096), hi[j] = w23^(4096 j) (j < 2048),
// w23 = 3^((p-1)/2^23)  (ff.rs:215-223 prim_nth_root).
struct RootTables {
  const u32 *lo, *hi;
};
FF_HD u32 root_pow(const RootTables &T, u32 e) { return ff::canon(ff::mont_mul(T.lo[e & 4095u], T.hi[e >> 12])); }

// geometric scale c * g^i (Montgomery form): lo[i] = c*g^i (i < 4096), hi[j] = g^(4096 j)
struct GeoTables {
  const u32 *lo, *hi;
};
FF_HD u32 geo_pow(const GeoTables &T, u64 i) { return ff::canon(ff::mont_mul(T.lo[i & 4095u], T.hi[i >> 12])); }

enum ScaleMode { SCALE_NONE = 0, SCALE_CONST = 1, SCALE_GEO = 2 };

struct PassArgs {
  const u32 *in;
  u32 *out;
  int logL;                  // transform length L = 2^logL, 3 <= logL <= 12
  u64 in_batch, out_batch;   // element stride between the transforms of a batch (one CTA each)
  u64 n_valid;               // inputs at index >= n_valid are zero and are not read
  const wpair *tw;           // tw[e] = w_L^(+-e) for e < L, Shoup form (field.cuh shoup_mul)
  wpair w8[4];               // 1, w_8, w_8^2, w_8^3 in the transform direction, Shoup form
  int pre_mode;              // ScaleMode applied to loaded elements (index = coefficient index); GEO only
  GeoTables pre_geo;
  int post_mode;             // ScaleMode applied to stored elements (index = output index)
  u32 post_const;            // Montgomery form
  GeoTables post_geo;
};

// in-register radix-2^LOGR DIF; a[i] ends up holding output bitrev(i).  Inputs/outputs in [0, 2p).
template <int LOGR>
FF_HD void dif_regs(u32 *a, const wpair *w8) {
  constexpr int R = 1 << LOGR;
#pragma unroll
  for (int len = R; len >= 2; len >>= 1) {
    const int h = len >> 1;
#pragma unroll
    for (int blk = 0; blk < R; blk += len) {
#pragma unroll
      for (int j = 0; j < h; j++) {
        u32 u = a[blk + j], v = a[blk + j + h];
        a[blk + j] = ff::red2p(u + v);
        u32 d = u + ff::P2 - v;
        a[blk + j + h] = (j == 0) ? ff::red2p(d) : ff::shoup_mul(d, w8[j * (8 / len)].w, w8[j * (8 / len)].s);
      }
    }
  }
}
template <int LOGR>
FF_HD constexpr int bitrev(int i) {
  int r = 0;
  for (int b = 0; b < LOGR; b++) r |= ((i >> b) & 1) << (LOGR - 1 - b);
  return r;
}

// Phase A of a round for one thread: gather inputs (HBM if FIRST else smem), butterflies, twiddles.
// regs[i*R + k] = output k of task i ; T = 8/R tasks per thread per round.
template <int LOGR, bool FIRST>
FF_HD void round_load_compute(u32 tid, u32 nthreads, u32 b, const PassArgs &A, int logS, const u32 *smem, u32 *regs) {
  constexpr int R = 1 << LOGR, T = 8 / R;
  const int logL = A.logL;
#pragma unroll
  for (int i = 0; i < T; i++) {
    const u32 u = tid + (u32)i * nthreads;
    u32 a[R];
#pragma unroll
    for (int j = 0; j < R; j++) {
      const u32 l = u + ((u32)j << (logL - LOGR));
      u32 v;
      if (FIRST) {
        v = (u64)l < A.n_valid ? A.in[(u64)b * A.in_batch + l] : 0u;
        if (A.pre_mode == SCALE_GEO) v = ff::mont_mul(v, geo_pow(A.pre_geo, l));
      } else {
        v = smem[l];
      }
      a[j] = v;
    }
    const u32 p = u >> logS;
    dif_regs<LOGR>(a, A.w8);
#pragma unroll
    for (int pos = 0; pos < R; pos++) {
      const int kk = bitrev<LOGR>(pos);
      u32 val = a[pos];
      if (kk != 0) {
        const wpair t = A.tw[(p * (u32)kk) << logS];
        val = ff::shoup_mul(val, t.w, t.s);
      }
      regs[i * R + kk] = val;
    }
  }
}

// Phase B: scatter the register values (smem, or HBM with the fused post-scale if LAST).
template <int LOGR, bool LAST>
FF_HD void round_store(u32 tid, u32 nthreads, u32 b, const PassArgs &A, int logS, u32 *smem, const u32 *regs) {
  constexpr int R = 1 << LOGR, T = 8 / R;
#pragma unroll
  for (int i = 0; i < T; i++) {
    const u32 u = tid + (u32)i * nthreads;
    const u32 q = u & ((1u << logS) - 1u), p = u >> logS;
#pragma unroll
    for (int kk = 0; kk < R; kk++) {
      const u32 l = q + (((p << LOGR) + (u32)kk) << logS);
      u32 val = regs[i * R + kk];
      if (!LAST) {
        smem[l] = val;
      } else {
        if (A.post_mode == SCALE_CONST)
          val = ff::mont_mul(val, A.post_const);
        else if (A.post_mode == SCALE_GEO)
          val = ff::mont_mul(val, geo_pow(A.post_geo, l));
        A.out[(u64)b * A.out_batch + l] = ff::canon(val);
      }
    }
  }
}

// radix plan for L = 2^logL (logL >= 3): first round 2^(logL mod 3) if non-zero, then radix 8.
FF_HD int plan_rounds(int logL, int *logr) {
  int n = 0, rem = logL % 3;
  if (rem) logr[n++] = rem;
  for (int i = 0; i < logL / 3; i++) logr[n++] = 3;
  return n;
}

}  // namespace ntt
