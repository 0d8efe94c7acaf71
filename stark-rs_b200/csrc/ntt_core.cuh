// ntt_core.cuh -- radix-2^k Stockham NTT building blocks (host+device inline).
//
// Replaces Polynomial::eval_domain / interpolate_domain on structured domains (reference
// src/univariate/eval.rs:16-21 -- O(n*m) Horner; interpolate.rs:6-44 -- O(n^3) Lagrange) by an
// O(n log n) transform with NATURAL-order input and output (the Merkle leaf index and the FRI
// (i, i+N/2) pairing both assume natural order, fri.rs:57-91, 118-127).
//
// A "pass" transforms a tile [L rows][C columns] along L; every thread keeps 32 field elements in
// registers per round (radix-8 butterfly x 4 adjacent columns, or the equivalent for radix 4/2), does
// the butterflies there, and exchanges through shared memory between rounds.  The first round loads
// straight from HBM (128-bit, coalesced along the columns) and the last round stores straight to HBM,
// so a pass touches each element exactly once in each direction.
//
// Stockham DIF round with s = product of previous radices, R = this radix, u = task row in [0, L/R):
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
  int logL;    // sub-transform length L = 2^logL  (>= 3)
  int logC4;   // column quads (V=4) or columns (V=1) per tile = 2^logC4
  // input: element (l, c) of tile (b, ct) is in[b*in_batch + l*in_stride + ct*C + c]; its coefficient index
  // (for zero-padding and pre-scaling) is l*in_stride + ct*C + c.
  u64 in_batch, in_stride, n_valid;
  // output col mode: out[b*out_batch + l*out_stride + ct*C + c]; row mode: out[b*out_batch + (ct*C + c)*L + l]
  u64 out_batch, out_stride;
  int tiles_per_batch;
  const u32 *tw;   // tw[e] = w_L^(+-e) for e < L, Montgomery form
  u32 w8[4];       // 1, w_8, w_8^2, w_8^3 in the pass direction, Montgomery form
  // row mode (pass 1 of a four-step transform of size N): multiply by w_N^(+-(col * l))
  RootTables roots;
  int shiftN;      // 23 - log2(N)
  int inverse;
  int pre_mode;    // ScaleMode applied to loaded elements (index = coefficient index)
  GeoTables pre_geo;
  int post_mode;   // ScaleMode applied to stored elements, col mode only (index = l*out_stride + col)
  u32 post_const;  // Montgomery form
  GeoTables post_geo;
};

template <int V>
struct Slot;  // one shared-memory slot = V adjacent columns of one row
template <>
struct Slot<4> {
  typedef q4 type;
};
template <>
struct Slot<1> {
  typedef u32 type;
};

// bank-conflict swizzle of the quad index inside a row (see DESIGN.md "NTT shared-memory layout")
FF_HD u32 slot_index(u32 l, u32 c4, int logC4) {
  u32 sw;
  if (logC4 >= 3)
    sw = l & 7u;
  else if (logC4 == 2)
    sw = (l >> 1) & 3u;
  else if (logC4 == 1)
    sw = (l >> 2) & 1u;
  else
    sw = 0;
  return (l << logC4) + (c4 ^ sw);
}

// in-register radix-2^LOGR DIF; a[i] ends up holding output bitrev(i).  Inputs/outputs in [0, 2p).
template <int LOGR>
FF_HD void dif_regs(u32 *a, const u32 *w8) {
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
        a[blk + j + h] = (j == 0) ? ff::red2p(d) : ff::mont_mul(d, w8[j * (8 / len)]);
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

template <int V>
FF_HD void unpack(const typename Slot<V>::type &s, u32 *v);
template <>
FF_HD void unpack<4>(const q4 &s, u32 *v) {
  v[0] = s.x, v[1] = s.y, v[2] = s.z, v[3] = s.w;
}
template <>
FF_HD void unpack<1>(const u32 &s, u32 *v) {
  v[0] = s;
}
template <int V>
FF_HD typename Slot<V>::type pack(const u32 *v);
template <>
FF_HD q4 pack<4>(const u32 *v) {
  q4 r = {v[0], v[1], v[2], v[3]};
  return r;
}
template <>
FF_HD u32 pack<1>(const u32 *v) {
  return v[0];
}

// Task decode.  t in [0, (L/R) * C4).  c-fastest (coalesced 128-bit column access) or l-fastest
// (coalesced scalar access along l for the transposing store).
struct Task {
  u32 u, c4;
};
FF_HD Task decode(u32 t, int logL, int LOGR, int logC4, bool lfast) {
  Task k;
  if (lfast) {
    k.u = t & ((1u << (logL - LOGR)) - 1u);
    k.c4 = t >> (logL - LOGR);
  } else {
    k.c4 = t & ((1u << logC4) - 1u);
    k.u = t >> logC4;
  }
  return k;
}

// Phase A of a round for one thread: gather inputs (HBM if FIRST else smem), butterflies, twiddles.
// regs: 32 values laid out [task i][column v][k] ; T = 8/R tasks per thread per round.
template <int LOGR, int V, bool FIRST>
FF_HD void round_load_compute(u32 tid, u32 nthreads, u32 tile, const PassArgs &A, int logS,
                              const typename Slot<V>::type *smem, bool lfast_next_store, u32 *regs) {
  constexpr int R = 1 << LOGR, T = 8 / R;
  const int logL = A.logL, logC4 = A.logC4;
  const u32 b = tile / (u32)A.tiles_per_batch, ct = tile % (u32)A.tiles_per_batch;
  const u32 Ccols = (u32)V << logC4;
#pragma unroll
  for (int i = 0; i < T; i++) {
    const u32 t = tid + (u32)i * nthreads;
    const Task k = decode(t, logL, LOGR, logC4, lfast_next_store);
    u32 a[V][R];
#pragma unroll
    for (int j = 0; j < R; j++) {
      const u32 l = k.u + ((u32)j << (logL - LOGR));
      u32 v[V];
      if (FIRST) {
        const u64 cidx = (u64)l * A.in_stride + (u64)ct * Ccols + (u64)k.c4 * V;  // coefficient index
        const u32 *src = A.in + (u64)b * A.in_batch + cidx;
        if (cidx + V <= A.n_valid) {
          unpack<V>(*reinterpret_cast<const typename Slot<V>::type *>(src), v);
        } else {
#pragma unroll
          for (int x = 0; x < V; x++) v[x] = (cidx + x < A.n_valid) ? src[x] : 0u;
        }
        if (A.pre_mode == SCALE_GEO) {
#pragma unroll
          for (int x = 0; x < V; x++) v[x] = ff::mont_mul(v[x], geo_pow(A.pre_geo, cidx + x));
        }
      } else {
        unpack<V>(smem[slot_index(l, k.c4, logC4)], v);
      }
#pragma unroll
      for (int x = 0; x < V; x++) a[x][j] = v[x];
    }
    // twiddles w_L^(s*p*k), shared by the V columns
    const u32 p = k.u >> logS;
    u32 tw[R];
#pragma unroll
    for (int kk = 1; kk < R; kk++) tw[kk] = A.tw[(p * (u32)kk) << logS];
#pragma unroll
    for (int x = 0; x < V; x++) {
      dif_regs<LOGR>(a[x], A.w8);
#pragma unroll
      for (int pos = 0; pos < R; pos++) {
        const int kk = bitrev<LOGR>(pos);
        u32 val = a[x][pos];
        if (kk != 0) val = ff::mont_mul(val, tw[kk]);
        regs[(i * V + x) * R + kk] = val;
      }
    }
  }
}

// Phase B: scatter the 32 register values (smem, or HBM if LAST).
template <int LOGR, int V, bool LAST, bool ROWOUT>
FF_HD void round_store(u32 tid, u32 nthreads, u32 tile, const PassArgs &A, int logS, typename Slot<V>::type *smem,
                       const u32 *regs) {
  constexpr int R = 1 << LOGR, T = 8 / R;
  const int logL = A.logL, logC4 = A.logC4;
  const u32 b = tile / (u32)A.tiles_per_batch, ct = tile % (u32)A.tiles_per_batch;
  const u32 Ccols = (u32)V << logC4;
  const bool lfast = LAST && ROWOUT;
#pragma unroll
  for (int i = 0; i < T; i++) {
    const u32 t = tid + (u32)i * nthreads;
    const Task k = decode(t, logL, LOGR, logC4, lfast);
    const u32 q = k.u & ((1u << logS) - 1u), p = k.u >> logS;
#pragma unroll
    for (int kk = 0; kk < R; kk++) {
      const u32 l = q + (((p << LOGR) + (u32)kk) << logS);
      u32 v[V];
#pragma unroll
      for (int x = 0; x < V; x++) v[x] = regs[(i * V + x) * R + kk];
      if (!LAST) {
        smem[slot_index(l, k.c4, logC4)] = pack<V>(v);
      } else if (ROWOUT) {
        // transposing store + four-step twiddle w_N^(col*l); scalar, coalesced along l
#pragma unroll
        for (int x = 0; x < V; x++) {
          const u32 col = ct * Ccols + k.c4 * V + x;
          u32 e = (u32)(((u64)col * l) << A.shiftN) & ((1u << 23) - 1u);
          if (A.inverse) e = ((1u << 23) - e) & ((1u << 23) - 1u);
          const u32 w = root_pow(A.roots, e);
          A.out[(u64)b * A.out_batch + ((u64)col << logL) + l] = ff::canon(ff::mont_mul(v[x], w));
        }
      } else {
        const u64 oidx = (u64)l * A.out_stride + (u64)ct * Ccols + (u64)k.c4 * V;
#pragma unroll
        for (int x = 0; x < V; x++) {
          u32 val = v[x];
          if (A.post_mode == SCALE_CONST)
            val = ff::mont_mul(val, A.post_const);
          else if (A.post_mode == SCALE_GEO)
            val = ff::mont_mul(val, geo_pow(A.post_geo, oidx + x));
          v[x] = ff::canon(val);
        }
        *reinterpret_cast<typename Slot<V>::type *>(A.out + (u64)b * A.out_batch + oidx) = pack<V>(v);
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
