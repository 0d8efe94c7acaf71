// ntt_pass.cuh -- multi-pass Stockham NTT for N >= 2^13 (host+device inline, so tests/emul can run it on the CPU).
//
// Replaces Polynomial::eval_domain / interpolate_domain on structured domains (reference
// src/univariate/eval.rs:16-21, interpolate.rs:6-44), natural order in and out.
//
// The transform is a radix-R_1, R_2(, R_3) decimation-in-frequency Stockham (autosort) FFT, one HBM/L2 PASS per
// factor, R_i = 2^5 .. 2^8.  Pass i (s = R_1..R_{i-1}, M = N/R_i, u = q + s p with q < s):
//     in_j  = X[u + j M]                          j < R_i
//     out_k = Y[q + s (R_i p + k)] = (sum_j in_j w_R^(jk)) * w_N^(s p k)
// A CTA owns a TILE of R rows x C = 2^12/R adjacent columns u (4096 elements, 16 KB of shared memory, 128 threads,
// 32 elements per thread), so 8 CTAs are resident per SM and one CTA's barriers hide behind the others' work (8192-
// element / 256-thread tiles measured 4-9 % slower: coarser barriers, and 512 tiles of a 2^22 transform balance worse
// over 148 SMs than 1024).
// Inside the tile the R-point column DFTs are again Stockham rounds (radix 8, plus one radix-2/4 round when
// log2 R is not a multiple of 3) on registers: a thread holds an 8-row x 4-column block, does the butterflies there
// and exchanges through shared memory between rounds.  The first round reads HBM directly (128-bit, coalesced along
// the columns), the last round writes HBM directly.
//   FIRST  pass (s = 1):    every column has its own p = u, so the outer twiddle w_N^(u k) is per element (the
//                           "four-step" twiddle); the tile's output Y[R u + k] is one contiguous R*C block.
//   MIDDLE pass:            the C columns share p, the outer twiddle is a per-row table in shared memory.
//   LAST   pass (s = N/R):  no outer twiddle; fused post-scale (n^-1, or n^-1 * g^i for the coset); runs in place.
// All shapes are template parameters: the index arithmetic folds to shifts and constants.
#pragma once
#include "ntt_core.cuh"

namespace ntt2 {
using ff::u32;
using ff::u64;
using ntt::GeoTables;
using ntt::q4;
using ntt::RootTables;
using ntt::wpair;

enum Kind { FIRST = 0, MIDDLE = 1, LAST = 2 };
constexpr int TILE_LOG = 12;   // elements per tile
constexpr int NT = 128;        // threads per CTA
#ifndef NTT2_PIN_LAST
#define NTT2_PIN_LAST 0x05
#define NTT2_PIN_OTHER 0x77
#endif
#ifndef NTT2_PIN_LAST_BIG
#define NTT2_PIN_LAST_BIG 0x07
#endif
constexpr int PIN_LAST = NTT2_PIN_LAST, PIN_OTHER = NTT2_PIN_OTHER, PIN_LAST_BIG = NTT2_PIN_LAST_BIG;   // dif_lazy: stages whose sums go to the ALU pipe

struct PassParams {
  const u32 *in;
  u32 *out;
  u64 in_batch, out_batch;   // element stride between the transforms of a batch
  u64 n_valid;               // FIRST: inputs at coefficient index >= n_valid are zero and are not read
  int logN;                  // transform length
  int logS;                  // s = product of the previous passes' radices
  int log_tiles;             // tiles per transform = 2^(logN - 13)
  const wpair *tw_in;        // inner twiddles of this pass's radix in Shoup form, laid out per round (fill_inner_twiddles)
  wpair w8[4];               // 1, w_8, w_8^2, w_8^3 in the pass direction, Shoup form
  RootTables roots;          // w_{2^23}^e two-level table
  int inverse;
  int pre_mode;              // FIRST: ntt::ScaleMode on loaded elements (index = coefficient index)
  GeoTables pre_geo;
  int post_mode;             // LAST: ntt::ScaleMode on stored elements (index = output index)
  wpair post_const;          // Shoup form
  GeoTables post_geo;
  u32 zero;                  // 0, known only at run time (ff::add_alu)
  wpair pre_g1, pre_gj;      // FIRST, geometric pre-scale c g^i: g and g^(N / first-round radix) in Shoup form
  wpair post_g1, post_gk;    // LAST, geometric post-scale c g^i: g and g^(N/8) in Shoup form (the walks along a row / down a column block)
  const wpair *otw_tab;      // MIDDLE: w_{2^16}^(+-e), e < 2^16, Shoup form; w_N^(s e) = otw_tab[e << otw_shift]
  int otw_shift;             //         16 - (logN - logS)
  const wpair *row_tab;      // FIRST: w_N^(+-row), row < 2048, Shoup form
};

// pass radices (log2) for a transform of length 2^log_n, 13 <= log_n <= 23, largest first.  The FIRST pass stores one
// contiguous R x C block per tile through transposed scalar stores whose contiguous run is 4 * R/8 bytes per 8 lanes: with
// R = 2^8 a warp writes 128 contiguous bytes, with R = 2^6 only 32.  Giving the FIRST pass the SMALLEST radix instead
// (it is the FMA-bound pass: 757 / 685 / 599 slots per 32 elements for radix 2^8 / 2^7 / 2^6) was
// measured slower for that reason: 16 x 2^22 305 vs 292 us with {6, 8, 8}, 2^20 alone 19.8 vs 17.5 us with {6, 7, 7}.
// Radix 2^9 would leave 32-byte rows in a 4096-element tile and is not used.  Returns the pass count.
inline int pass_plan(int log_n, int *r) {
  static const int PLAN[11][3] = {{7, 6, 0}, {7, 7, 0}, {8, 7, 0}, {8, 8, 0}, {6, 6, 5}, {6, 6, 6},
                                  {7, 6, 6}, {7, 7, 6}, {7, 7, 7}, {8, 7, 7}, {8, 8, 7}};
  for (int i = 0; i < 3; i++) r[i] = PLAN[log_n - 13][i];
  return r[2] ? 3 : 2;
}

template <int LOGR>
struct Plan {
  static constexpr int B0 = LOGR % 3;
  static constexpr int NR = LOGR / 3 + (B0 ? 1 : 0);
  FF_HD static constexpr int lr(int r) { return (r == 0 && B0) ? B0 : 3; }
  FF_HD static constexpr int logs(int r) {
    int s = 0;
    for (int i = 0; i < r; i++) s += lr(i);
    return s;
  }
};

// Inner twiddles of the R-point column DFTs, one block of 512 pairs per radix R = 2^5 .. 2^8 (out + (LOGR - 5) * 512):
// round 0 at +0, round 1 at +256, each as [p'][k], k < RAD = the round's radix: entry = w_R^(+-(s' p' k)).  A task needs
// the RAD - 1 twiddles of ONE p', which are adjacent here: one address computation and RAD/2 128-bit loads instead of
// RAD - 1 indexed 64-bit loads (each of which cost an IMAD + a 64-bit IMAD.WIDE on the saturated FMA pipe).
// Radices 2^9 .. 2^11 (the 16384-element tiles of the two-pass plans) follow at inner_twiddle_offset(): three rounds of R pairs.
constexpr int INNER_TWIDDLE_PAIRS = 2048 + 3 * (512 + 1024 + 2048);
FF_HD constexpr int inner_twiddle_offset(int logr) {
  return logr <= 8 ? (logr - 5) * 512 : (logr == 9 ? 2048 : (logr == 10 ? 2048 + 1536 : 2048 + 1536 + 3072));
}
inline void fill_inner_twiddles(wpair *out, int inverse) {
  for (int i = 0; i < INNER_TWIDDLE_PAIRS; i++) out[i] = wpair{1, ff::shoup_of(1)};
  for (int logr = 5; logr <= 11; logr++) {
    u32 w = ff::pow(ff::GEN, (ff::P - 1) >> logr);   // ff.rs:215-223
    if (inverse) w = ff::inv(w);
    const int b0 = logr % 3, nr = logr / 3 + (b0 ? 1 : 0);
    wpair *blk = out + inner_twiddle_offset(logr);
    const int stride = logr <= 8 ? 256 : (1 << logr);
    int logs = 0;
    for (int round = 0; round + 1 < nr; round++) {
      const int lr = (round == 0 && b0) ? b0 : 3;
      const u32 n_pp = 1u << (logr - lr - logs);
      u32 wp = 1;                                     // w^(pp << logs)
      const u32 wstep = ff::pow(w, 1ull << logs);
      for (u32 pp = 0; pp < n_pp; pp++) {
        u32 v = 1;
        for (u32 k = 0; k < (1u << lr); k++) {
          blk[stride * round + (pp << lr) + k] = wpair{v, ff::shoup_of(v)};
          v = ff::mul(v, wp);
        }
        wp = ff::mul(wp, wstep);
      }
      logs += lr;
    }
  }
}

// shared-memory slot (16 bytes = 4 adjacent columns of one row) of (row l, column quad c4).  A 128-bit access is
// served per quarter-warp, conflict-free when its 8 lanes hit 8 distinct slots mod 8; the XOR keeps that true for the
// column-fastest (same row) and the row-fastest (8 adjacent rows, same quad) lane mappings used below.
template <int LOGC4, int TL = TILE_LOG>
FF_HD u32 slot(u32 l, u32 c4) {
  if (LOGC4 >= 3) return (l << LOGC4) + (c4 ^ (l & 7u));
  if (LOGC4 == 1) {
    // two quads per row: four rows form one 8-slot group g = l >> 2.  Quarter-warps touch four consecutive rows of one
    // group (column-fastest), rows 2^LR apart in four or two groups (the strided stores of round 0) or eight consecutive
    // rows of one quad (row-fastest): bits 1-2 rotate with the group, bit 0 (the quad) flips with its parity.
    const u32 g = l >> 2, pos = ((l & 3u) << 1) | c4;
    return (g << 3) + (pos ^ (((g & 3u) << 1) | (g & 1u)));
  }
  // LOGC4 == 2: two rows form one 8-slot group.  Quarter-warps touch two rows that differ in exactly one of the row
  // bits 0, 2, 3 (column-fastest; also bit 1 in the radix-2 first round of the 16384-element radix-2^10 tiles) or eight
  // consecutive rows (row-fastest): the low two slot bits rotate with the row pair, the half (bit 2) flips with row
  // bits 2 and 3 (and 1 for those tiles).
  const u32 sr = l >> 1, pos = ((l & 1u) << 2) | c4;
  if (TL == TILE_LOG) return (sr << 3) + (pos ^ ((sr & 3u) | ((((sr >> 1) ^ (sr >> 2)) & 1u) << 2)));   // radix 2^8: no bit-1 pattern
  return (sr << 3) + (pos ^ ((sr & 3u) | (((sr ^ (sr >> 1) ^ (sr >> 2)) & 1u) << 2)));
}

// radix-2^LR DIF on a[0 .. 2^LR) in [0, 2p); a[pos] ends up holding output bitrev(pos).  The outputs are LAZY, in
// [0, 4p): the caller either multiplies them by a canonical twiddle (-> [0, 2p)) or reduces them.
// PIN: bit i set = the sums of stage i (0 = the first, widest stage) are pinned to the ALU pipe (ff::add_alu); bit 4 + i
// = so are the two-input differences of stage i (the other differences are three-input IADD3 anyway).  Chosen per pass
// kind from the SASS pipe counts (tools/sass_hist.py).
template <int LR, int PIN>
FF_HD void dif_lazy(u32 *a, const wpair *w8, u32 zero) {
  constexpr int R = 1 << LR;
  int stage = 0;
#pragma unroll
  for (int len = R; len >= 2; len >>= 1, stage++) {
    const int h = len >> 1;
    const bool final_stage = len == 2;
    const bool pin = (PIN >> stage) & 1;
#pragma unroll
    for (int blk = 0; blk < R; blk += len) {
#pragma unroll
      for (int j = 0; j < h; j++) {
        const u32 u = a[blk + j], v = a[blk + j + h];
        const u32 s = pin ? ff::add_alu(u, v, zero) : u + v;
        // the difference is a three-input IADD3 except where ptxas folds the + 2p into the range reduction that follows
        // (j == 0) and is left with a two-input u - v, which it may move to the FMA pipe: pin that one as well
        const u32 d = (((PIN >> (4 + stage)) & 1) && j == 0 && !final_stage) ? ff::add_alu(u, 0u - v, zero) + ff::P2 : u + ff::P2 - v;
        if (final_stage) {
          a[blk + j] = s, a[blk + j + h] = d;
        } else {
          a[blk + j] = ff::red2p(s);
          a[blk + j + h] = (j == 0) ? ff::red2p(d) : ff::shoup_mul(d, w8[j * (8 / len)].w, w8[j * (8 / len)].s);
        }
      }
    }
  }
}
template <int LR>
FF_HD constexpr int bitrev(int i) {
  int r = 0;
  for (int b = 0; b < LR; b++) r |= ((i >> b) & 1) << (LR - 1 - b);
  return r;
}

struct TileCtx {
  const u32 *in;   // + batch offset
  u32 *out;        // + batch offset
  u32 col0;        // first column u of the tile
  u32 q0, p;       // col0 = q0 + s p
};

template <int LOGR, int TL = TILE_LOG>
FF_HD TileCtx tile_ctx(const PassParams &A, u32 tile) {
  constexpr int LOGC = TL - LOGR;
  TileCtx T;
  const u32 b = tile >> A.log_tiles, ct = tile & ((1u << A.log_tiles) - 1u);
  T.in = A.in + (u64)b * A.in_batch;
  T.out = A.out + (u64)b * A.out_batch;
  T.col0 = ct << LOGC;
  T.q0 = T.col0 & ((1u << A.logS) - 1u);
  T.p = T.col0 >> A.logS;
  return T;
}

// w_N^(+-e) for e < N
FF_HD u32 root_n(const PassParams &A, u32 e) {
  u32 e23 = e << (23 - A.logN);
  if (A.inverse) e23 = ((1u << 23) - e23) & ((1u << 23) - 1u);
  return ntt::root_pow(A.roots, e23);
}

// MIDDLE: otw[k] = w_N^(+-(s p k)) = w_{N/s}^(+-(p k)), k < R, in Shoup form (per-tile table in shared memory).
// N/s <= 2^16 for every plan (the first radix is >= 2^(logN - 16)), so the pairs come straight from one table.
template <int LOGR, int TL = TILE_LOG>
FF_HD void fill_outer_table(u32 tid, const PassParams &A, const TileCtx &T, wpair *otw) {
  for (u32 k = tid; k < (1u << LOGR); k += (1u << (TL - 5))) otw[k] = A.otw_tab[(T.p * k) << A.otw_shift];
}

// task -> (row group u', column quad c4).  Column-fastest: adjacent lanes touch adjacent 16-byte slots of one row.
// Row-fastest (only the last round of a FIRST pass): adjacent lanes own adjacent output rows, which are adjacent
// addresses of the transposed store.
template <int LOGR, int LR, bool ROWFAST, int TL = TILE_LOG>
FF_HD void decode(u32 t, u32 &up, u32 &c4) {
  constexpr int LOGC4 = TL - LOGR - 2;
  if (ROWFAST) {
    up = t & ((1u << (LOGR - LR)) - 1u);
    c4 = t >> (LOGR - LR);
  } else {
    c4 = t & ((1u << LOGC4) - 1u);
    up = t >> LOGC4;
  }
}

// Phase A of round ROUND: gather the inputs (HBM in round 0, shared memory afterwards), butterflies, twiddles.
// regs[(i*4 + x)*RAD + k] = output k of column x of task i.
// MODE specialises the per-element options at compile time (a run-time test per element costs more than it looks: the
// LAST pass spent 64 of its ~1000 instructions per thread on the post-scale switch):
//   FIRST:  bit 0 = some inputs are zero padding (n_valid < N), bit 1 = geometric pre-scale
//   LAST:   the ntt::ScaleMode of the post-scale
template <int LOGR, int KIND, int ROUND, int MODE, int TL = TILE_LOG>
FF_HD void round_compute(u32 tid, const PassParams &A, const TileCtx &T, const q4 *smem, const wpair *otw, u32 *regs) {
  typedef Plan<LOGR> PL;
  constexpr int LR = PL::lr(ROUND), LOGS = PL::logs(ROUND), RAD = 1 << LR, NTASK = 8 >> LR;
  constexpr int LOGC4 = TL - LOGR - 2;
  constexpr u32 NT = 1u << (TL - 5);               // 32 elements per thread
  constexpr int TWS = LOGR <= 8 ? 256 : (1 << LOGR);   // pairs per round in the inner-twiddle block (fill_inner_twiddles)
  constexpr bool LASTR = ROUND == PL::NR - 1;
  constexpr bool ROWFAST = LASTR && KIND == FIRST;
  const int logM = A.logN - LOGR;
#pragma unroll
  for (int i = 0; i < NTASK; i++) {
    u32 up, c4;
    decode<LOGR, LR, ROWFAST, TL>(tid + (u32)i * NT, up, c4);
    u32 a[4][RAD];
    // FIRST, geometric pre-scale c g^cidx: one table look-up per task, then Shoup walks by g^(N/RAD) from row to row
    // (the RAD inputs of a butterfly are N/RAD apart) and by g along the quad
    u32 Gj = 0;
    if (ROUND == 0 && KIND == FIRST && (MODE & 2)) Gj = ntt::geo_pow(A.pre_geo, ((u64)up << logM) + T.col0 + 4u * c4);
#pragma unroll
    for (int j = 0; j < RAD; j++) {
      const u32 l = up + ((u32)j << (LOGR - LR));
      q4 v;
      if (ROUND == 0) {
        const u64 cidx = ((u64)l << logM) + T.col0 + 4u * c4;   // coefficient index of the quad's first element
        if (KIND != FIRST || !(MODE & 1) || cidx + 4 <= A.n_valid) {
          v = *reinterpret_cast<const q4 *>(T.in + cidx);
        } else if (cidx >= A.n_valid) {
          v.x = v.y = v.z = v.w = 0u;   // zero padding is never read
        } else {
          v.x = T.in[cidx];
          v.y = cidx + 1 < A.n_valid ? T.in[cidx + 1] : 0u;
          v.z = cidx + 2 < A.n_valid ? T.in[cidx + 2] : 0u;
          v.w = 0u;
        }
        if (KIND == FIRST && (MODE & 2)) {
          if (!(MODE & 1) || cidx < A.n_valid) {   // zero padding needs no scaling
            u32 t = Gj;
            v.x = ff::mont_mul(v.x, t), t = ff::canon(ff::shoup_mul(t, A.pre_g1.w, A.pre_g1.s));
            v.y = ff::mont_mul(v.y, t), t = ff::canon(ff::shoup_mul(t, A.pre_g1.w, A.pre_g1.s));
            v.z = ff::mont_mul(v.z, t), t = ff::canon(ff::shoup_mul(t, A.pre_g1.w, A.pre_g1.s));
            v.w = ff::mont_mul(v.w, t);
          }
          if (j + 1 < RAD) Gj = ff::canon(ff::shoup_mul(Gj, A.pre_gj.w, A.pre_gj.s));
        }
      } else {
        v = smem[slot<LOGC4, TL>(l, c4)];
      }
      a[0][j] = v.x, a[1][j] = v.y, a[2][j] = v.z, a[3][j] = v.w;
    }
    // inner twiddles w_R^(s' p' k), shared by the four columns (the last round has p' = 0: none)
    u32 tw[RAD], step[RAD];   // !LASTR: tw = plain twiddle, step = its Shoup companion
    if (!LASTR) {
      const u32 pp = up >> LOGS;
      const q4 *tp = reinterpret_cast<const q4 *>(A.tw_in + TWS * ROUND + (pp << LR));   // pairs (2h, 2h + 1)
#pragma unroll
      for (int h = 0; h < RAD / 2; h++) {
        const q4 t = tp[h];
        if (h) tw[2 * h] = t.x, step[2 * h] = t.y;
        tw[2 * h + 1] = t.z, step[2 * h + 1] = t.w;
      }
    }
    if (LASTR && KIND == FIRST) {
      // The R-point DFTs of the four columns, then the outer ("four-step") twiddle of element (column cb + x, row_k =
      // up + S' k), S' = R/8:   w^((cb + x) row_k) = [w^(cb up) (w^(cb S'))^k] * (w^row_k)^x.
      // The bracket walks along k with a run-time ratio (Montgomery products); along x the ratio w^row_k comes from a
      // 256-entry table in Shoup form, so the walk t <- t * w^row_k is a Shoup product (4 FMA-pipe slots, not 5).
#pragma unroll
      for (int x = 0; x < 4; x++) dif_lazy<LR, PIN_OTHER>(a[x], A.w8, A.zero);
      const u32 cb = T.col0 + 4u * c4;
      const u32 wc = root_n(A, cb << LOGS);
      u32 base = root_n(A, cb * up);   // Montgomery form
#pragma unroll
      for (int k = 0; k < RAD; k++) {
        const wpair st = A.row_tab[up + ((u32)k << LOGS)];
        u32 t = base;
#pragma unroll
        for (int x = 0; x < 4; x++) {
          regs[(i * 4 + x) * RAD + k] = ff::mont_mul(a[x][bitrev<LR>(k)], t);
          if (x < 3) t = ff::canon(ff::shoup_mul(t, st.w, st.s));
        }
        if (k + 1 < RAD) base = ff::canon(ff::mont_mul(base, wc));
      }
      continue;
    }
    if (LASTR && KIND == LAST && MODE == ntt::SCALE_GEO) {
      // Fused geometric post-scale c g^oidx, oidx = O + x + k N/8 (the eight outputs of a column are N/8 apart): one
      // table look-up per task, then Shoup walks by g^(N/8) down the column block and by g along the four columns
      // (was: two table look-ups and two Montgomery products per element).
#pragma unroll
      for (int x = 0; x < 4; x++) dif_lazy<LR, PIN_OTHER>(a[x], A.w8, A.zero);
      u32 G = ntt::geo_pow(A.post_geo, (u64)T.col0 + 4u * c4 + ((u64)up << A.logS));   // Montgomery form
#pragma unroll
      for (int k = 0; k < RAD; k++) {
        u32 t = G;
#pragma unroll
        for (int x = 0; x < 4; x++) {
          regs[(i * 4 + x) * RAD + k] = ff::canon(ff::mont_mul(a[x][bitrev<LR>(k)], t));
          if (x < 3) t = ff::canon(ff::shoup_mul(t, A.post_g1.w, A.post_g1.s));
        }
        if (k + 1 < RAD) G = ff::canon(ff::shoup_mul(G, A.post_gk.w, A.post_gk.s));
      }
      continue;
    }
#pragma unroll
    for (int x = 0; x < 4; x++) {
      dif_lazy<LR, ((KIND == LAST && MODE == ntt::SCALE_NONE) ? (TL == TILE_LOG ? PIN_LAST : PIN_LAST_BIG) : PIN_OTHER)>(a[x], A.w8, A.zero);
#pragma unroll
      for (int pos = 0; pos < RAD; pos++) {
        const int k = bitrev<LR>(pos);
        u32 val = a[x][pos];
        if (!LASTR) {
          val = k ? ff::shoup_mul(val, tw[k], step[k]) : ff::red2p(val);
        } else {
          const u32 row = up + ((u32)k << LOGS);   // output row of the R-point DFT (p' = 0, q' = u')
          if (KIND == MIDDLE) {
            const wpair o = otw[row];
            val = ff::shoup_mul(val, o.w, o.s);
          } else {
            if (MODE == ntt::SCALE_CONST) {
              val = ff::canon(ff::shoup_mul(val, A.post_const.w, A.post_const.s));
            } else if (MODE == ntt::SCALE_GEO) {
              const u64 oidx = (u64)T.col0 + 4u * c4 + (u32)x + ((u64)row << A.logS);
              val = ff::canon(ff::mont_mul(val, ntt::geo_pow(A.post_geo, oidx)));
            } else {
              val = ff::canon4(val);
            }
          }
        }
        regs[(i * 4 + x) * RAD + k] = val;
      }
    }
  }
}

// Phase B of round ROUND: scatter the register block (shared memory, or HBM in the last round)
template <int LOGR, int KIND, int ROUND, int TL = TILE_LOG>
FF_HD void round_store(u32 tid, const PassParams &A, const TileCtx &T, q4 *smem, const u32 *regs) {
  typedef Plan<LOGR> PL;
  constexpr int LR = PL::lr(ROUND), LOGS = PL::logs(ROUND), RAD = 1 << LR, NTASK = 8 >> LR;
  constexpr int LOGC4 = TL - LOGR - 2;
  constexpr u32 NT = 1u << (TL - 5);
  constexpr bool LASTR = ROUND == PL::NR - 1;
  constexpr bool ROWFAST = LASTR && KIND == FIRST;
#pragma unroll
  for (int i = 0; i < NTASK; i++) {
    u32 up, c4;
    decode<LOGR, LR, ROWFAST, TL>(tid + (u32)i * NT, up, c4);
    const u32 qp = up & ((1u << LOGS) - 1u), pp = up >> LOGS;
    // MIDDLE / LAST output rows of one task are 2^(LOGS + logS) elements apart: a running 64-bit pointer (two ALU adds per
    // store) instead of an address product per store on the FMA pipe
    u32 *run = nullptr;
    if (LASTR && KIND != FIRST) run = T.out + T.q0 + 4u * c4 + ((((u64)T.p << LOGR) + qp + ((pp << LR) << LOGS)) << A.logS);
    const u64 run_step = 1ull << (LOGS + A.logS);
#pragma unroll
    for (int k = 0; k < RAD; k++) {
      const u32 l = qp + (((pp << LR) + (u32)k) << LOGS);
      const u32 *r = regs + i * 4 * RAD + k;
      if (!LASTR) {
        q4 v = {r[0], r[RAD], r[2 * RAD], r[3 * RAD]};
        smem[slot<LOGC4, TL>(l, c4)] = v;
      } else if (KIND == FIRST) {
        // Y[R u + k]: the tile's output is one contiguous block; scalar stores, coalesced along the rows
        u32 *dst = T.out + (((u64)T.col0 + 4u * c4) << LOGR) + l;
#pragma unroll
        for (int x = 0; x < 4; x++) dst[(u64)x << LOGR] = r[x * RAD];
      } else {
        // Y[q + s (R p + k)]
        q4 v = {r[0], r[RAD], r[2 * RAD], r[3 * RAD]};
        *reinterpret_cast<q4 *>(run) = v;
        run += run_step;
      }
    }
  }
}

}  // namespace ntt2
