// fri.cu -- FRI prover on the device (reference src/fri.rs:29-311, src/fiat_shamir.rs, src/stream.rs:35-64).
//
// Fri::prove runs as ONE stream of kernel launches with no host round trip: per round leaf hashes + tree
// (merkle.cu), a one-thread transcript kernel that absorbs the root and draws alpha (fiat_shamir.rs:15-25),
// and the fused fold kernel; then index sampling (fri.rs:176-213), and kernels that gather the revealed
// values and authentication paths straight into ProofStream::serialize's byte layout (stream.rs:35-64).  The
// host makes a single D2H copy of the finished proof.  Every tree is built once and kept (the reference
// rebuilds each tree 2-3x, fri.rs:127, 288-298); results are identical because a tree is a pure function of
// its codeword.
//
// fold_codeword (fri.rs:57-91) in closed form:
//   out[i] = (c[i] + c[i+h]) / 2  +  (alpha mod p) / (2 offset) * w^-i * (c[i] - c[i+h])
// The per-element exp + two xgcd inversions of the reference become one geometric twiddle: w^-i comes from a
// two-level table of (w0^-1)^e shared by all rounds (round r uses e = i * 2^r), i.e. the inverted domain
// constants are produced by ONE host inversion, not one per element.  12 bytes of HBM traffic per output.
#include <vector>

#include "common.cuh"
#include "hash.cuh"
#include "merkle.h"
#include "merkle_dev.cuh"
#include "mgpu.h"

using hs::State;
using ntt::GeoTables;

// ---------------------------------------------------------------------------------------- transcript


// ---------------------------------------------------------------------------------------------- fold

// V outputs per thread (128-bit accesses for V = 4).  tw(i) = K * g_r^i, g_r^i = g0^(i << r).
// Outputs [i0, i1) of the fold of a codeword of length 2h (the whole fold is i0 = 0, i1 = h; a rank of the sharded
// prover folds only its own output range); out is indexed by the GLOBAL output index i.
template <int V>
__global__ void __launch_bounds__(256) k_fri_fold(const u32 *__restrict__ cw, u32 *__restrict__ out, size_t h, size_t i0,
                                                  size_t i1, int r, GeoTables G, u32 g_r_m,
                                                  const u32 *__restrict__ alpha_m, u32 alpha_val, u32 inv2off_m) {
  pdl_entry();
  // alpha (Montgomery form) comes from device memory inside Fri::commit (written by the transcript step) and by value
  // from the stand-alone entry points
  const u32 K = ff::canon(ff::mont_mul(alpha_m ? *alpha_m : alpha_val, inv2off_m));
  const size_t stride = (size_t)gridDim.x * blockDim.x * V;
  for (size_t i = i0 + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * V; i < i1; i += stride) {
    u32 a[V], b[V], o[V];
    if constexpr (V == 4) {
      const uint4 x = *reinterpret_cast<const uint4 *>(cw + i), y = *reinterpret_cast<const uint4 *>(cw + h + i);
      a[0] = x.x, a[1] = x.y, a[2] = x.z, a[3] = x.w;
      b[0] = y.x, b[1] = y.y, b[2] = y.z, b[3] = y.w;
    } else {
      a[0] = cw[i], b[0] = cw[h + i];
    }
    u32 tw = ff::canon(ff::mont_mul(ntt::geo_pow(G, (u64)i << r), K));
#pragma unroll
    for (int k = 0; k < V; k++) {
      const u32 s = a[k] + b[k];           // < 2p
      const u32 d = a[k] + ff::P - b[k];   // < 2p
      o[k] = ff::canon4(ff::half(s) + ff::mont_mul(d, tw));
      if (k + 1 < V) tw = ff::canon(ff::mont_mul(tw, g_r_m));
    }
    if constexpr (V == 4)
      *reinterpret_cast<uint4 *>(out + i) = make_uint4(o[0], o[1], o[2], o[3]);
    else
      out[i] = o[0];
  }
}

// fold fused with its all-gather (SURVEY 8(e), K8 "fused variant"): the rank computes its output range and stores every
// result straight into ALL ranks' replicas of the next codeword -- one multimem.st through the NVSwitch multicast
// address when there is one (the switch replicates the 16-byte store to every GPU), else one store per peer over
// NVLink P2P -- so the transfer overlaps the arithmetic and no separate collective runs.  The caller brackets the
// launch with the symmetric-memory barrier.
struct FoldPeers {
  u32 *out[8];   // every rank's replica of the next codeword (own included), indexed by the global output index
  u32 *mc;       // multicast address of the same buffer, or nullptr
  int n;
};
__global__ void __launch_bounds__(256) k_fri_fold_bcast(const u32 *__restrict__ cw, const __grid_constant__ FoldPeers P, size_t h, size_t i0,
                                                        size_t i1, int r, GeoTables G, u32 g_r_m,
                                                        const u32 *__restrict__ alpha_m, u32 alpha_val, u32 inv2off_m) {
  const u32 K = ff::canon(ff::mont_mul(alpha_m ? *alpha_m : alpha_val, inv2off_m));
  const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (size_t i = i0 + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < i1; i += stride) {
    const uint4 x = *reinterpret_cast<const uint4 *>(cw + i), y = *reinterpret_cast<const uint4 *>(cw + h + i);
    const u32 a[4] = {x.x, x.y, x.z, x.w}, b[4] = {y.x, y.y, y.z, y.w};
    u32 o[4];
    u32 tw = ff::canon(ff::mont_mul(ntt::geo_pow(G, (u64)i << r), K));
#pragma unroll
    for (int k = 0; k < 4; k++) {
      o[k] = ff::canon4(ff::half(a[k] + b[k]) + ff::mont_mul(a[k] + ff::P - b[k], tw));
      if (k < 3) tw = ff::canon(ff::mont_mul(tw, g_r_m));
    }
    if (P.mc != nullptr) {
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(P.mc + i), "r"(o[0]), "r"(o[1]),
                   "r"(o[2]), "r"(o[3])
                   : "memory");
    } else {
      const uint4 v = make_uint4(o[0], o[1], o[2], o[3]);
#pragma unroll 1
      for (int g = 0; g < P.n; g++) *reinterpret_cast<uint4 *>(P.out[g] + i) = v;
    }
  }
}

// -------------------------------------------------------------------------------------------- FRI tail

// The last rounds of Fri::commit (fri.rs:116-147), codeword <= 2^10, as ONE single-CTA kernel: per round leaf hashes,
// the whole tree, the transcript round and the fold, back to back through shared memory.  These rounds are a pure
// dependency chain (root -> alpha -> fold -> next leaves) of ~log2(n)+3 hash latencies each; run as separate launches
// they cost a launch + drain per link.
constexpr int TAIL_LOG = 10, TAIL_NT = 512, TAIL_MAX_ROUNDS = 12;
struct TailArgs {
  u32 n_rounds, first_round, len0;
  u32 *cw[TAIL_MAX_ROUNDS + 1];   // cw[i] = codeword of tail round i; cw[i + 1] receives its fold
  u8 *nodes[TAIL_MAX_ROUNDS];     // tree of cw[i]
  u32 inv2off_m[TAIL_MAX_ROUNDS]; // (2 offset_r)^-1, Montgomery
  GeoTables G;                    // (w0^-1)^e table shared by all rounds (round r uses e = i << r)
  TranscriptDev *T;
  u8 *roots;                      // entries of the first tail round onwards
  u64 *alpha_raw;
  u32 *alpha_m;
};
__global__ void __launch_bounds__(TAIL_NT, 1) k_fri_tail(const __grid_constant__ TailArgs A) {
  __shared__ __align__(16) u8 sm[(1 << TAIL_LOG) * 16];
  pdl_entry();
  const u32 t = threadIdx.x, one = blockDim.y;
  u32 len = A.len0;
  for (u32 i = 0; i < A.n_rounds; i++) {
    const u32 *cw = A.cw[i];
    u8 *nodes = A.nodes[i];
    // leaves (fri.rs:118-121), two per thread
    if (2 * t < len) {
      u32 wa[8], wb[8];
      if (2 * t + 1 < len) {
        const uint2 v = *reinterpret_cast<const uint2 *>(cw + 2 * t);
        hs2::leaf2(v.x, v.y, wa, wb, one);
        store_hash(nodes + 64 * (size_t)t + 32, wb);
      } else {
        hs2::leaf2(cw[2 * t], 0u, wa, wb, one);
      }
      store_hash(nodes + 64 * (size_t)t, wa);
    }
    __syncthreads();
    u32 levels = 0;
    for (u32 c = len; c > 1; c >>= 1) levels++;
    cta_climb<TAIL_NT>(nodes, len, 0, 0, len, levels, nodes, sm, one);
    const bool last = i + 1 == A.n_rounds;
    if (t < 32) {
      const TranscriptArgs tr = {A.T, A.roots + 32 * i, last ? 0 : 1, A.alpha_raw + i, A.alpha_m + i};
      transcript_round_warp(tr, len > 1 ? sm : nodes);   // fri.rs:129-138
    }
    __syncthreads();
    if (last) break;   // fri.rs:133-135
    // fold_codeword (fri.rs:57-91), same closed form as k_fri_fold
    const u32 K = ff::canon(ff::mont_mul(A.alpha_m[i], A.inv2off_m[i]));
    const u32 h = len >> 1, r = A.first_round + i;
    u32 *out = A.cw[i + 1];
    for (u32 j = t; j < h; j += TAIL_NT) {
      const u32 a = cw[j], b = cw[h + j];
      const u32 tw = ff::canon(ff::mont_mul(ntt::geo_pow(A.G, (u64)j << r), K));
      out[j] = ff::canon4(ff::half(a + b) + ff::mont_mul(a + ff::P - b, tw));
    }
    __syncthreads();
    len = h;
  }
}

// fold_codeword (fri.rs:57-91) fused with the leaf hashing of the NEXT round (fri.rs:118-121): writes the folded
// codeword and its leaf hashes in one pass, two outputs per thread (hs2).  h even.
__global__ void __launch_bounds__(64) k_fold_leaf1(const u32 *__restrict__ cw, u32 *__restrict__ out, size_t h, int r,
                                                    GeoTables G, u32 g_r_m, const u32 *__restrict__ alpha_m,
                                                    u32 inv2off_m, u8 *__restrict__ leaves) {
  pdl_entry();
  const size_t i = 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= h) return;
  const u32 K = ff::canon(ff::mont_mul(*alpha_m, inv2off_m));
  const uint2 x = *reinterpret_cast<const uint2 *>(cw + i), y = *reinterpret_cast<const uint2 *>(cw + h + i);
  const u32 tw0 = ff::canon(ff::mont_mul(ntt::geo_pow(G, (u64)i << r), K));
  const u32 tw1 = ff::canon(ff::mont_mul(tw0, g_r_m));
  const u32 o0 = ff::canon4(ff::half(x.x + y.x) + ff::mont_mul(x.x + ff::P - y.x, tw0));
  const u32 o1 = ff::canon4(ff::half(x.y + y.y) + ff::mont_mul(x.y + ff::P - y.y, tw1));
  *reinterpret_cast<uint2 *>(out + i) = make_uint2(o0, o1);
  u32 wa[8], wb[8];
  hs2::leaf2(o0, o1, wa, wb, blockDim.y);
  store_hash(leaves + 32 * i, wa);
  store_hash(leaves + 32 * i + 32, wb);
}

// ------------------------------------------------------------------------------------ index sampling

// Fri::sample_indices (fri.rs:176-213) with seed = Hash::from_u64(FiatShamir::challenge) (fri.rs:272, hash.rs:37-39).
// Thread 0 draws the challenge from the transcript and hashes it into the seed; every candidate message is
// seed || counter (36 bytes, fri.rs:198-200), so the sponge after the first chunk (the seed) is computed once and shared;
// candidates for a batch of 256 counters are then hashed in parallel (4 absorbed bytes + 9 mixes each, registers only)
// and thread 0 applies the sequential reject rule.
__global__ void __launch_bounds__(256) k_sample_indices(const TranscriptDev *T, u64 *challenge_out, u64 size, u64 reduced,
                                                        u32 number, u64 *out) {
  __shared__ u64 cand[256];
  __shared__ u64 seen_red[256];   // idx % reduced of the accepted indices
  __shared__ u32 pre[32];
  __shared__ u32 got;
  pdl_entry();
  if (threadIdx.x == 0) {
    u64 c;
    if (T->npend == 0) {
      // FiatShamir::challenge on a chunk-aligned transcript: 8 finalisation mixes of the stored sponge
      State st;
#pragma unroll
      for (int i = 0; i < 32; i++) st.s[i] = T->s[i];
      hs::mix_lazy<false>(st);
#pragma unroll 1
      for (int k = 0; k < 7; k++) hs::mix_lazy<true>(st);
      c = 0;
#pragma unroll
      for (int b = 0; b < 8; b++) c |= (u64)((st.s[b] + hs::rc_at(b)) & 0xffu) << (8 * b);
    } else {
      c = tr_challenge(*T);
    }
    *challenge_out = c;
    State st;
    hs::init(st);
#pragma unroll
    for (int i = 0; i < 8; i++) hs::absorb_byte(st, i, (u32)(c >> (8 * i)) & 0xffu);
    hs::mix_lazy<false>(st);
#pragma unroll 1
    for (int k = 0; k < 8; k++) hs::mix_lazy<true>(st);
    hs::settle(st);
    u32 seed[8];
    hs::pack_words(st, seed);
    State s2;
    hs::init(s2);
    hs::absorb_words_mix<false>(s2, seed);   // round constants of this mix still pending
#pragma unroll
    for (int i = 0; i < 32; i++) pre[i] = s2.s[i];
    got = 0;
  }
  __syncthreads();
  if (number == 0) return;
  // size and reduced are powers of two for every Fri::new-accepted domain (fri.rs:37-44): % is a mask then
  const bool pow2 = (size & (size - 1)) == 0 && (reduced & (reduced - 1)) == 0;
  for (u32 base = 0;; base += 256) {
    const u32 counter = base + threadIdx.x;
    State st;
#pragma unroll
    for (int i = 0; i < 32; i++) st.s[i] = pre[i];
    hs::settle(st);
#pragma unroll
    for (int i = 0; i < 4; i++) hs::absorb_byte(st, i, (counter >> (8 * i)) & 0xffu);   // u32 LE, fri.rs:199-200
    hs::mix_lazy<false>(st);
#pragma unroll 1
    for (int k = 0; k < 8; k++) hs::mix_lazy<true>(st);
    hs::settle(st);
    u64 v = 0;
#pragma unroll
    for (int b = 24; b < 32; b++) v = (v << 8) | (u64)(st.s[b] & 0xffu);   // fri.rs:168-174: low 64 bits, big-endian
    cand[threadIdx.x] = pow2 ? (v & (size - 1)) : (v % size);
    __syncthreads();
    if (threadIdx.x == 0) {
      u32 g = got;
      for (u32 k = 0; k < 256 && g < number; k++) {
        const u64 idx = cand[k], ri = pow2 ? (idx & (reduced - 1)) : (idx % reduced);
        bool seen = false;
        for (u32 j = 0; j < g; j++) seen |= seen_red[j & 255u] == ri && j < 256;
        if (g >= 256)   // more accepted indices than the cache holds: fall back to the stored list
          for (u32 j = 256; j < g; j++) seen |= (out[j] % reduced) == ri;
        if (!seen) {
          if (g < 256) seen_red[g] = ri;
          out[g++] = idx;
        }
      }
      got = g;
    }
    __syncthreads();
    if (got >= number) break;
  }
}

// host wrapper for the verifier (verify.cu): the same index sampling on a transcript the verifier rebuilt
int fri_sample_indices_dev(stark_ctx *ctx, const TranscriptDev *T, u64 *d_challenge, u64 size, u64 reduced, u32 number,
                           u64 *d_out) {
  LAUNCH_PDL(ctx, "sample_indices", 0, k_sample_indices, 1, 256, T, d_challenge, size, reduced, number, d_out);
  return STARK_OK;
}

// ------------------------------------------------------------------------------------ proof assembly

__device__ __forceinline__ void put_u64(u8 *d, u64 v) {
#pragma unroll
  for (int b = 0; b < 8; b++) d[b] = (u8)(v >> (8 * b));
}
// R x [0x00, root]  then  [0x02, len u64, values u64...]   (fri.rs:129, 151; stream.rs:39-53)
__global__ void k_proof_header(u8 *out, const u8 *roots, u32 R, const u32 *last, u64 last_len) {
  pdl_entry();
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < R) {
    u8 *d = out + 33 * t;
    d[0] = 0;
    for (int i = 0; i < 32; i++) d[1 + i] = roots[32 * t + i];
  }
  u8 *base = out + 33 * (u64)R;
  if (t == 0) {
    base[0] = 2;
    put_u64(base + 1, last_len);
  }
  if (t < last_len) put_u64(base + 9 + 8 * t, last[t]);
}
// the query rounds (fri.rs:215-248, 280-307), all in one launch (blockIdx.y = round i): nq x [0x02, 3, a, b, c]  then
// nq x (path a, path b, path c), each path [0x03, depth u64, depth hashes] (stream.rs:54-60); indices folded as
// fri.rs:282-285 (top % (len_i / 2), which equals the reference's cumulative reduction because the lengths halve).
constexpr int MAX_FRI_ROUNDS = 24;
struct ProofRoundsArgs {
  const u32 *cw[MAX_FRI_ROUNDS];
  const u8 *nodes[MAX_FRI_ROUNDS];
  u64 out_off[MAX_FRI_ROUNDS];
  u64 len0;
  u32 nq, depth0;
};
__global__ void k_proof_rounds(u8 *proof, const __grid_constant__ ProofRoundsArgs A, const u64 *top) {
  pdl_entry();
  const u32 i = blockIdx.y, nq = A.nq;
  u8 *out = proof + A.out_off[i];
  const u32 *cur = A.cw[i], *nxt = A.cw[i + 1];
  const u8 *cur_nodes = A.nodes[i], *nxt_nodes = A.nodes[i + 1];
  const u64 cur_len = A.len0 >> i;
  const u32 depth_cur = A.depth0 - i;
  const u64 half = cur_len >> 1;
  const u32 depth_nxt = depth_cur - 1;
  const u64 triples = 33ull * nq;
  const u64 path_cur = 9 + 32ull * depth_cur, path_nxt = 9 + 32ull * depth_nxt;
  const u64 per_q = 2 * path_cur + path_nxt;
  const u32 hashes_per_q = 2 * depth_cur + depth_nxt;
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nq) {
    const u64 c = top[t] % half;
    u8 *d = out + 33 * t;
    d[0] = 2;
    put_u64(d + 1, 3);
    put_u64(d + 9, cur[c]);
    put_u64(d + 17, cur[c + half]);
    put_u64(d + 25, nxt[c]);
    u8 *p = out + triples + per_q * t;
    p[0] = 3, put_u64(p + 1, depth_cur);
    p += path_cur;
    p[0] = 3, put_u64(p + 1, depth_cur);
    p += path_cur;
    p[0] = 3, put_u64(p + 1, depth_nxt);
  }
  if (t >= (u64)nq * hashes_per_q) return;
  const u32 q = (u32)(t / hashes_per_q), k = (u32)(t % hashes_per_q);
  const u64 c = top[q] % half;
  u64 leaf, n_tree;
  const u8 *nodes;
  u32 level;
  u8 *dst = out + triples + per_q * q;
  if (k < depth_cur) {
    leaf = c, nodes = cur_nodes, n_tree = cur_len, level = k, dst += 9 + 32ull * level;
  } else if (k < 2 * depth_cur) {
    leaf = c + half, nodes = cur_nodes, n_tree = cur_len, level = k - depth_cur, dst += path_cur + 9 + 32ull * level;
  } else {
    leaf = c, nodes = nxt_nodes, n_tree = half, level = k - 2 * depth_cur, dst += 2 * path_cur + 9 + 32ull * level;
  }
  const u64 sib = (leaf >> level) ^ 1;  // merkle.rs:73-77
  const uint4 *src = reinterpret_cast<const uint4 *>(nodes + 32 * ((2 * n_tree - 2 * (n_tree >> level)) + sib));
  const uint4 x = src[0], y = src[1];
  const u32 w[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
#pragma unroll
  for (int b = 0; b < 32; b++) dst[b] = (u8)(w[b >> 2] >> (8 * (b & 3)));
}

// ---------------------------------------------------------------------------------------- host logic

struct stark_fri_state {
  stark_ctx *ctx;
  u32 rounds;                 // Fri::num_rounds()
  std::vector<u32 *> cw;      // codewords[i] (fri.rs:110,140,153); cw[0] borrowed if !own0
  std::vector<size_t> len;
  std::vector<stark_tree *> trees;  // tree of codewords[i] for i < rounds (sharded rounds: this rank's SUBTREE)
  std::vector<bool> borrowed;       // cw[i] lives in the multi-GPU window (not freed here)
  // sharded prover (mgpu.h): rounds [0, sharded) have their tree split by leaf range over the ranks
  u32 sharded = 0, log_per[MG_MAX_ROUNDS] = {};   // log2 of the leaves per rank of a sharded round
  std::vector<u8 *> top_nodes;      // per sharded round: the replicated top tree over the ranks' subtree roots
  bool own0;
  u8 *d_roots;        // rounds * 32
  u64 *d_alpha_raw;   // rounds entries (last unused)
  u32 *d_alpha_m;
  TranscriptDev *d_tr;
  u32 nq, ef;
};

static int fri_check(stark_ctx *ctx, size_t n, u32 ef, u32 *rounds, u32 nq) {
  if (n == 0 || (n & (n - 1))) return stark_fail(ctx, STARK_ERR_ARG, "Domain length must be power of 2");    // fri.rs:37-40
  if (ef == 0 || (ef & (ef - 1))) return stark_fail(ctx, STARK_ERR_ARG, "Expansion factor must be power of 2");  // fri.rs:41-44
  if (ef < 4) return stark_fail(ctx, STARK_ERR_ARG, "Expansion factor must be at least 4");                  // fri.rs:45
  size_t len = n;
  u32 r = 0;
  while (len > ef && 4 * (size_t)nq < len) len >>= 1, r++;  // fri.rs:93-103
  *rounds = r;
  return STARK_OK;
}

static void fri_state_free(stark_fri_state *s) {
  if (!s) return;
  stark_ctx *ctx = s->ctx;
  for (size_t i = 0; i < s->cw.size(); i++)
    if ((i > 0 || s->own0) && !(i < s->borrowed.size() && s->borrowed[i])) dev_free(ctx, s->cw[i]);
  for (stark_tree *t : s->trees) stark_merkle_free(t);
  for (u8 *t : s->top_nodes) dev_free(ctx, t);
  dev_free(ctx, s->d_roots), dev_free(ctx, s->d_alpha_raw), dev_free(ctx, s->d_alpha_m), dev_free(ctx, s->d_tr);
  delete s;
}

// out is indexed by the global output index (out[i] for i in [i0, i1))
static int fold_launch_range(stark_ctx *ctx, const u32 *cw, u32 *out, size_t h, size_t i0, size_t i1, int r, GeoTables G,
                             u32 g_r_m, const u32 *alpha_m, u32 inv2off_m, u32 alpha_val = 0) {
  if (i1 <= i0) return STARK_OK;
  const size_t cnt = i1 - i0;
  // 128-bit path only when every address the kernel touches is 16-byte aligned (out may be an offset view)
  const bool aligned = (((uintptr_t)(cw + i0) | (uintptr_t)(cw + h + i0) | (uintptr_t)(out + i0)) & 15u) == 0;
  if (h % 4 == 0 && i0 % 4 == 0 && cnt % 4 == 0 && aligned) {
    size_t blocks = (cnt / 4 + 255) / 256, cap = (size_t)ctx->sm_count * 8;
    LAUNCH_PDL(ctx, "fri_fold", 12ull * cnt, k_fri_fold<4>, (u32)(blocks < cap ? blocks : cap), 256, cw, out, h, i0, i1, r, G,
               g_r_m, alpha_m, alpha_val, inv2off_m);
  } else {
    LAUNCH_PDL(ctx, "fri_fold", 12ull * cnt, k_fri_fold<1>, (u32)((cnt + 255) / 256), 256, cw, out, h, i0, i1, r, G, g_r_m,
               alpha_m, alpha_val, inv2off_m);
  }
  return STARK_OK;
}
static int fold_launch(stark_ctx *ctx, const u32 *cw, u32 *out, size_t h, int r, GeoTables G, u32 g_r_m,
                       const u32 *alpha_m, u32 inv2off_m) {
  return fold_launch_range(ctx, cw, out, h, 0, h, r, G, g_r_m, alpha_m, inv2off_m);
}

// The domain constants a round of Fri::commit carries (fri.rs:146-147): off_r / g_r belong to round r, the fold INTO
// round r uses those of round r-1.
struct RoundConsts {
  u32 off_r, g_r, off_prev, g_prev;
  size_t len;     // length of codewords[r]
};
static void round_consts_next(RoundConsts &c) {
  c.off_prev = c.off_r, c.g_prev = c.g_r;
  c.off_r = ff::mul(c.off_r, c.off_r), c.g_r = ff::mul(c.g_r, c.g_r);  // fri.rs:146-147
  c.len /= 2;
}

// Allocations and transcript of Fri::commit (fri.rs:105-115).  cw0 is borrowed unless copy0.
static int fri_commit_setup(stark_ctx *ctx, const u32 *cw0, size_t n, u32 offset, u32 omega, u32 ef, u32 nq,
                            const u8 *transcript, size_t transcript_len, bool copy0, stark_fri_state **out, GeoTables *G_out,
                            RoundConsts *C_out) {
  u32 R = 0;
  ST_TRY(fri_check(ctx, n, ef, &R, nq));
  // fold_codeword divides by x = offset * w^i (fri.rs:72-78 -> ff.rs:182)
  if (R > 1 && (offset == 0 || (omega == 0 && n > 2))) return stark_fail(ctx, STARK_ERR_ARG, "no division by zero");
  stark_fri_state *s = new stark_fri_state();
  s->ctx = ctx, s->rounds = R, s->own0 = copy0, s->nq = nq, s->ef = ef;
  s->d_roots = nullptr, s->d_alpha_raw = nullptr, s->d_alpha_m = nullptr, s->d_tr = nullptr;
  int rc = STARK_OK;
#define TRY_(e)                      \
  if (rc == STARK_OK) rc = (e);
  TRY_(dev_alloc(ctx, (void **)&s->d_roots, (size_t)(R ? R : 1) * 32));
  TRY_(dev_alloc(ctx, (void **)&s->d_alpha_raw, (size_t)(R ? R : 1) * 8));
  TRY_(dev_alloc(ctx, (void **)&s->d_alpha_m, (size_t)(R ? R : 1) * 4));
  TRY_(dev_alloc(ctx, (void **)&s->d_tr, sizeof(TranscriptDev)));
#undef TRY_
  if (rc == STARK_OK) {
    TranscriptDev t;
    tr_init(t);
    if (transcript_len) tr_absorb(t, transcript, transcript_len);  // host side: FiatShamir prefix
    if (cudaMemcpyAsync(s->d_tr, &t, sizeof t, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
      rc = stark_fail(ctx, STARK_ERR_CUDA, "H2D copy failed");
  }
  u32 *cur = const_cast<u32 *>(cw0);
  if (rc == STARK_OK && copy0) {
    u32 *c = nullptr;
    rc = dev_alloc(ctx, (void **)&c, n * 4);
    if (rc == STARK_OK && cudaMemcpyAsync(c, cw0, n * 4, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)
      rc = stark_fail(ctx, STARK_ERR_CUDA, "D2D copy failed");
    cur = c;
  }
  if (rc == STARK_OK) s->cw.push_back(cur), s->len.push_back(n), s->borrowed.push_back(false);
  GeoTables G = {nullptr, nullptr};
  u32 g0 = 1;
  if (rc == STARK_OK && R > 1) {
    g0 = ff::inv(omega);
    rc = geo_tables(ctx, g0, 1, n / 2, &G);
  }
  if (rc != STARK_OK) {
    fri_state_free(s);
    return rc;
  }
  *out = s, *G_out = G;
  *C_out = RoundConsts{offset, g0, offset, g0, n};
  return STARK_OK;
}

// Rounds [r0, R) of Fri::commit (fri.rs:116-147) on this device alone; codewords[r0 - 1] (if r0 > 0) and its alpha are in
// the state.  Round r:  [r > 0: fold of round r-1 fused with this round's leaf hashes]  ->  tree  ->  root + transcript.
static int fri_commit_rounds(stark_ctx *ctx, stark_fri_state *s, u32 r0, GeoTables G, RoundConsts C) {
  const u32 R = s->rounds;
  int rc = STARK_OK;
  u32 &off_r = C.off_r, &off_prev = C.off_prev, &g_prev = C.g_prev;
  size_t &len = C.len;
  for (u32 r = r0; r < R && rc == STARK_OK; r++) {
    const bool tail = len >= 2 && len <= ((size_t)1 << TAIL_LOG) && R - r <= (u32)TAIL_MAX_ROUNDS;
    if (r > 0) {
      u32 *nxt = nullptr;
      rc = dev_alloc(ctx, (void **)&nxt, len * 4);
      if (rc != STARK_OK) break;
      s->cw.push_back(nxt), s->len.push_back(len), s->borrowed.push_back(false);
    }
    if (tail) {
      // the remaining rounds in one single-CTA launch (its input codeword still has to be folded)
      if (r > 0) {
        const u32 inv2off_m = ff::to_mont(ff::inv(ff::mul(2, off_prev)));
        rc = fold_launch(ctx, s->cw[r - 1], s->cw[r], len, (int)(r - 1), G, ff::to_mont(g_prev), s->d_alpha_m + (r - 1), inv2off_m);
        if (rc != STARK_OK) break;
      }
      TailArgs A;
      memset(&A, 0, sizeof A);
      A.n_rounds = R - r, A.first_round = r, A.len0 = (u32)len;
      A.G = G, A.T = s->d_tr, A.roots = s->d_roots + 32 * r, A.alpha_raw = s->d_alpha_raw + r, A.alpha_m = s->d_alpha_m + r;
      A.cw[0] = s->cw[r];
      size_t l = len;
      u64 hashes = 0;
      for (u32 i = 0; i < A.n_rounds && rc == STARK_OK; i++) {
        stark_tree *tree = nullptr;
        rc = merkle_tree_alloc(ctx, l, &tree);
        if (rc != STARK_OK) break;
        s->trees.push_back(tree);
        A.nodes[i] = tree->nodes;
        A.inv2off_m[i] = ff::to_mont(ff::inv(ff::mul(2, off_r)));
        hashes += 2 * l - 1;
        if (i + 1 < A.n_rounds) {
          u32 *nxt = nullptr;
          rc = dev_alloc(ctx, (void **)&nxt, (l / 2) * 4);
          if (rc != STARK_OK) break;
          A.cw[i + 1] = nxt;
          s->cw.push_back(nxt), s->len.push_back(l / 2), s->borrowed.push_back(false);
          l /= 2;
          off_r = ff::mul(off_r, off_r);
        }
      }
      if (rc == STARK_OK) LAUNCH_PDL(ctx, "fri_tail", 64 * hashes, k_fri_tail, 1u, TAIL_NT, A);
      break;
    }
    stark_tree *tree = nullptr;
    rc = merkle_tree_alloc(ctx, len, &tree);
    if (rc != STARK_OK) break;
    s->trees.push_back(tree);
    if (r == 0) {
      rc = merkle_leaves_dev(ctx, s->cw[0], len, 1, 1, 0, tree->nodes);   // fri.rs:118-121
    } else if (len % 2 == 0) {
      const u32 inv2off_m = ff::to_mont(ff::inv(ff::mul(2, off_prev)));
      LAUNCH_PDL(ctx, "fold_leaf", 12ull * len + 32ull * len, k_fold_leaf1, (u32)((len / 2 + 63) / 64), 64,
                 (const u32 *)s->cw[r - 1], s->cw[r], len, (int)(r - 1), G, ff::to_mont(g_prev),
                 (const u32 *)(s->d_alpha_m + (r - 1)), inv2off_m, tree->nodes);
    } else {
      const u32 inv2off_m = ff::to_mont(ff::inv(ff::mul(2, off_prev)));
      rc = fold_launch(ctx, s->cw[r - 1], s->cw[r], len, (int)(r - 1), G, ff::to_mont(g_prev), s->d_alpha_m + (r - 1), inv2off_m);
      if (rc == STARK_OK) rc = merkle_leaves_dev(ctx, s->cw[r], len, 1, 1, 0, tree->nodes);
    }
    if (rc != STARK_OK) break;
    // tree (fri.rs:127); the kernel that produces the root also pushes it, absorbs it and, unless this is the last
    // round (fri.rs:133-135), draws alpha (fri.rs:129-138)
    const bool last = r == R - 1;
    const TranscriptArgs tr = {s->d_tr, s->d_roots + 32 * r, last ? 0 : 1, s->d_alpha_raw + r, s->d_alpha_m + r};
    rc = merkle_climb_dev(ctx, tree->nodes, len, &tr);
    if (rc != STARK_OK) break;
    round_consts_next(C);
  }
  if (rc == STARK_OK && cudaGetLastError() != cudaSuccess) rc = stark_fail(ctx, STARK_ERR_CUDA, "kernel launch failed");
  return rc;
}

// Fri::commit (fri.rs:105-156) on device.  cw0 is borrowed unless copy0.
static int fri_commit_dev(stark_ctx *ctx, const u32 *cw0, size_t n, u32 offset, u32 omega, u32 ef, u32 nq,
                          const u8 *transcript, size_t transcript_len, bool copy0, stark_fri_state **out) {
  stark_fri_state *s = nullptr;
  GeoTables G;
  RoundConsts C;
  ST_TRY(fri_commit_setup(ctx, cw0, n, offset, omega, ef, nq, transcript, transcript_len, copy0, &s, &G, &C));
  const int rc = fri_commit_rounds(ctx, s, 0, G, C);
  if (rc != STARK_OK) {
    fri_state_free(s);
    return rc;
  }
  *out = s;
  return STARK_OK;
}

struct ProofLayout {
  u32 R, n_cw;              // rounds, number of codewords (max(R,1))
  size_t last_len, header;  // bytes of roots + last codeword object
  std::vector<size_t> round_off;
  size_t total;
};
static void proof_layout(size_t n, u32 R, u32 nq, ProofLayout *L) {
  L->R = R, L->n_cw = R ? R : 1;
  L->last_len = n >> (L->n_cw - 1);
  L->header = 33 * (size_t)R + 9 + 8 * L->last_len;
  size_t off = L->header;
  L->round_off.clear();
  for (u32 i = 0; i + 1 < L->n_cw; i++) {
    L->round_off.push_back(off);
    size_t len = n >> i;
    u32 d = 0;
    for (size_t m = len; m > 1; m >>= 1) d++;
    off += (size_t)nq * (33 + 2 * (9 + 32 * (size_t)d) + (9 + 32 * (size_t)(d - 1)));
  }
  L->total = off;
}

// Fri::prove (fri.rs:250-311) + ProofStream::serialize: proof bytes to host in one D2H copy
static int fri_prove_dev(stark_ctx *ctx, const u32 *cw0, size_t n, size_t domain_length, u32 offset, u32 omega, u32 ef,
                         u32 nq, const u8 *transcript, size_t transcript_len, u8 *proof, size_t proof_cap,
                         size_t *proof_len, u64 *top_indices, bool sync = true) {
  // sync == false: everything (incl. the D2H copies of the result) is queued and the CALLER synchronises the stream --
  // the multi-column pipeline queues the latency chain first and its throughput work behind it
  if (n != domain_length)
    return stark_fail(ctx, STARK_ERR_ARG, "initial codeword length does not match domain length");  // fri.rs:256-260
  u32 R = 0;
  ST_TRY(fri_check(ctx, n, ef, &R, nq));
  ProofLayout L;
  proof_layout(n, R, nq, &L);
  if (proof_len) *proof_len = L.total;
  // fri.rs:183-192
  if ((size_t)nq > 2 * L.last_len) return stark_fail(ctx, STARK_ERR_ARG, "not enough entropy in indices wrt last codeword");
  if ((size_t)nq > L.last_len)
    return stark_fail(ctx, STARK_ERR_ARG, "cannot sample more indices than available in last codeword; requested: %u, available: %zu", nq, L.last_len);
  if (!proof || proof_cap < L.total) return stark_fail(ctx, STARK_ERR_ARG, "proof buffer too small: need %zu bytes", L.total);
  if (L.n_cw > (u32)MAX_FRI_ROUNDS) return stark_fail(ctx, STARK_ERR_ARG, "more than %d FRI rounds are not supported", MAX_FRI_ROUNDS);
  stark_fri_state *s = nullptr;
  ST_TRY(fri_commit_dev(ctx, cw0, n, offset, omega, ef, nq, transcript, transcript_len, false, &s));
  int rc = STARK_OK;
  u8 *d_proof = nullptr;
  u64 *d_top = nullptr, *d_seed = nullptr;
  rc = dev_alloc(ctx, (void **)&d_proof, L.total + 8 * (size_t)nq + 16);
  if (rc == STARK_OK) rc = dev_alloc(ctx, (void **)&d_seed, 8);
  if (rc == STARK_OK) {
    d_top = reinterpret_cast<u64 *>(d_proof + ((L.total + 7) & ~(size_t)7));
    if (ctx->prof_on) prof_begin(ctx, "query_phase", 0);
    const size_t sample_size = L.n_cw > 1 ? s->len[1] : s->len[0];       // fri.rs:266-270
    cudaError_t qe = launch_pdl(k_sample_indices, dim3(1), dim3(256), 0, ctx->stream, (const TranscriptDev *)s->d_tr, d_seed,
                                (u64)sample_size, (u64)L.last_len, nq, d_top);  // fri.rs:272-276
    const size_t hdr_threads = L.last_len > R ? L.last_len : R;
    if (qe == cudaSuccess)
      qe = launch_pdl(k_proof_header, dim3((u32)((hdr_threads + 255) / 256)), dim3(256), 0, ctx->stream, d_proof,
                      (const u8 *)s->d_roots, R, (const u32 *)s->cw[L.n_cw - 1], (u64)L.last_len);
    ctx->launches += 2;
    if (L.n_cw > 1 && nq) {
      ProofRoundsArgs PA;
      memset(&PA, 0, sizeof PA);
      u32 d0 = 0;
      for (size_t m = s->len[0]; m > 1; m >>= 1) d0++;
      PA.len0 = s->len[0], PA.nq = nq, PA.depth0 = d0;
      for (u32 i = 0; i < L.n_cw; i++) PA.cw[i] = s->cw[i], PA.nodes[i] = s->trees[i]->nodes;
      for (u32 i = 0; i + 1 < L.n_cw; i++) PA.out_off[i] = L.round_off[i];
      const size_t threads = (size_t)nq * (3 * d0 - 1);
      if (qe == cudaSuccess)
        qe = launch_pdl(k_proof_rounds, dim3((u32)((threads + 127) / 128), L.n_cw - 1), dim3(128), 0, ctx->stream, d_proof, PA,
                        (const u64 *)d_top);
      ctx->launches++;
    }
    if (qe != cudaSuccess) rc = stark_fail(ctx, STARK_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(qe));
    if (ctx->prof_on) prof_end(ctx);
    if (cudaGetLastError() != cudaSuccess) rc = stark_fail(ctx, STARK_ERR_CUDA, "kernel launch failed");
  }
  if (rc == STARK_OK && cudaMemcpyAsync(proof, d_proof, L.total, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
    rc = stark_fail(ctx, STARK_ERR_CUDA, "D2H copy failed");
  if (rc == STARK_OK && top_indices && nq &&
      cudaMemcpyAsync(top_indices, d_top, 8 * (size_t)nq, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
    rc = stark_fail(ctx, STARK_ERR_CUDA, "D2H copy failed");
  dev_free(ctx, d_proof), dev_free(ctx, d_seed);
  fri_state_free(s);
  if (rc == STARK_OK && sync) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = stark_fail(ctx, STARK_ERR_CUDA, "prove failed: %s", cudaGetErrorString(e));
  }
  return rc;
}

static int reduce_params(stark_ctx *ctx, uint64_t offset, uint64_t omega, u32 *off, u32 *om) {
  // FiniteField::mul reduces with u128 % p (ff.rs:138-144), so unreduced offset / omega act as their residues
  *off = ff::reduce64(offset), *om = ff::reduce64(omega);
  (void)ctx;
  return STARK_OK;
}

// ============================================================================ sharded prover (mgpu.h)
//
// Fri::prove on G ranks.  Every rank holds a replica of the round-0 codeword.  Rounds with at least 2^shard_log elements
// are SHARDED: rank g hashes the leaves [g n/G, (g+1) n/G) into a subtree, the subtree roots are exchanged through the
// peer windows inside the kernel that produces them (merkle_dev.cuh: mg_exchange_top), every rank climbs the G roots to
// the root and draws alpha; the fold into the next sharded round computes the rank's own output range and stores it
// into EVERY rank's replica (fused fold + all-gather over NVLink), hashing its leaves on the way.  The remaining rounds
// are latency-bound and run replicated on every rank (no exchange).  In the query phase the rank that owns a subtree
// node stores it into every rank's proof buffer, so that every rank ends up with the complete ProofStream::serialize
// bytes -- identical to the single-GPU path and to the reference for every G.

struct FoldPeersN {
  u32 *out[MG_MAX_RANKS];   // every rank's replica of the next codeword (own included), indexed by the global output index
  int n;
};
// outputs [i0, i1) of fold_codeword (fri.rs:57-91): two per thread, stored into all replicas, leaf-hashed (fri.rs:118-121)
// into this rank's subtree (leaf index i - i0).  i0, i1, h even.
__global__ void __launch_bounds__(64) k_mg_fold_leaf(const u32 *__restrict__ cw, const __grid_constant__ FoldPeersN P, size_t h,
                                                      size_t i0, size_t i1, int r, GeoTables G, u32 g_r_m,
                                                      const u32 *__restrict__ alpha_m, u32 inv2off_m, u8 *__restrict__ leaves) {
  pdl_entry();
  const size_t i = i0 + 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= i1) return;
  const u32 K = ff::canon(ff::mont_mul(*alpha_m, inv2off_m));
  // the codeword was written by the peers: read it past L1
  const uint2 x = __ldcg(reinterpret_cast<const uint2 *>(cw + i)), y = __ldcg(reinterpret_cast<const uint2 *>(cw + h + i));
  const u32 tw0 = ff::canon(ff::mont_mul(ntt::geo_pow(G, (u64)i << r), K));
  const u32 tw1 = ff::canon(ff::mont_mul(tw0, g_r_m));
  const u32 o0 = ff::canon4(ff::half(x.x + y.x) + ff::mont_mul(x.x + ff::P - y.x, tw0));
  const u32 o1 = ff::canon4(ff::half(x.y + y.y) + ff::mont_mul(x.y + ff::P - y.y, tw1));
  const uint2 o = make_uint2(o0, o1);
#pragma unroll 1
  for (int g = 0; g < P.n; g++) *reinterpret_cast<uint2 *>(P.out[g] + i) = o;
  u32 wa[8], wb[8];
  hs2::leaf2(o0, o1, wa, wb, blockDim.y);
  store_hash(leaves + 32 * (i - i0), wa);
  store_hash(leaves + 32 * (i - i0) + 32, wb);
  __threadfence_system();   // the slice is visible to the peers before this rank raises the round's root flag
}

// the query rounds (k_proof_rounds) when some trees are sharded.  Replicated data (triples, path headers, top-tree and
// unsharded-tree nodes) goes to this rank's own proof buffer; a node of a sharded level is written by the rank that
// owns it, into EVERY rank's buffer.
struct MgProofArgs {
  ProofRoundsArgs A;
  u32 log_per[MG_MAX_ROUNDS + 1];        // sharded tree: log2(leaves per rank); 0xff: whole tree in A.nodes
  const u8 *top[MG_MAX_ROUNDS + 1];      // sharded tree: replicated top tree (2G - 1 hashes)
  u8 *out[MG_MAX_RANKS];                 // every rank's proof buffer
  int world, rank;
};
__global__ void k_mg_proof_rounds(const __grid_constant__ MgProofArgs M, const u64 *top) {
  pdl_entry();
  const ProofRoundsArgs &A = M.A;
  const u32 i = blockIdx.y, nq = A.nq;
  u8 *out = M.out[M.rank] + A.out_off[i];
  const u64 cur_len = A.len0 >> i;
  const u32 depth_cur = A.depth0 - i;
  const u64 half = cur_len >> 1;
  const u32 depth_nxt = depth_cur - 1;
  const u64 triples = 33ull * nq;
  const u64 path_cur = 9 + 32ull * depth_cur, path_nxt = 9 + 32ull * depth_nxt;
  const u64 per_q = 2 * path_cur + path_nxt;
  const u32 hashes_per_q = 2 * depth_cur + depth_nxt;
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nq) {
    const u32 *cur = A.cw[i], *nxt = A.cw[i + 1];
    const u64 c = top[t] % half;
    u8 *d = out + 33 * t;
    d[0] = 2;
    put_u64(d + 1, 3);
    put_u64(d + 9, __ldcg(cur + c));
    put_u64(d + 17, __ldcg(cur + c + half));
    put_u64(d + 25, __ldcg(nxt + c));
    u8 *p = out + triples + per_q * t;
    p[0] = 3, put_u64(p + 1, depth_cur);
    p += path_cur;
    p[0] = 3, put_u64(p + 1, depth_cur);
    p += path_cur;
    p[0] = 3, put_u64(p + 1, depth_nxt);
  }
  if (t >= (u64)nq * hashes_per_q) return;
  const u32 q = (u32)(t / hashes_per_q), k = (u32)(t % hashes_per_q);
  const u64 c = top[q] % half;
  u64 leaf, n_tree;
  u32 level, tree;
  u64 dst_off = A.out_off[i] + triples + per_q * q;
  if (k < depth_cur) {
    leaf = c, tree = i, n_tree = cur_len, level = k, dst_off += 9 + 32ull * level;
  } else if (k < 2 * depth_cur) {
    leaf = c + half, tree = i, n_tree = cur_len, level = k - depth_cur, dst_off += path_cur + 9 + 32ull * level;
  } else {
    leaf = c, tree = i + 1, n_tree = half, level = k - 2 * depth_cur, dst_off += 2 * path_cur + 9 + 32ull * level;
  }
  const u64 sib = (leaf >> level) ^ 1;  // merkle.rs:73-77
  const u32 lp = M.log_per[tree];
  const uint4 *src;
  bool everywhere = false;
  if (lp == 0xffu) {
    src = reinterpret_cast<const uint4 *>(A.nodes[tree] + 32 * ((2 * n_tree - 2 * (n_tree >> level)) + sib));
  } else if (level >= lp) {
    const u64 G = (u64)M.world;
    const u32 tl = level - lp;
    src = reinterpret_cast<const uint4 *>(M.top[tree] + 32 * ((2 * G - 2 * (G >> tl)) + sib));
  } else {
    // level `level` of the sharded tree: rank g owns nodes [g per >> level, (g + 1) per >> level)
    const u64 per = 1ull << lp, owner = sib >> (lp - level);
    if (owner != (u64)M.rank) return;
    const u64 local = sib & ((per >> level) - 1);
    src = reinterpret_cast<const uint4 *>(A.nodes[tree] + 32 * ((2 * per - 2 * (per >> level)) + local));
    everywhere = true;
  }
  const uint4 x = src[0], y = src[1];
  const u32 w[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
  if (everywhere) {
#pragma unroll 1
    for (int g = 0; g < M.world; g++) {
      u8 *dst = M.out[g] + dst_off;
#pragma unroll
      for (int b = 0; b < 32; b++) dst[b] = (u8)(w[b >> 2] >> (8 * (b & 3)));
    }
    __threadfence_system();
  } else {
    u8 *dst = M.out[M.rank] + dst_off;
#pragma unroll
    for (int b = 0; b < 32; b++) dst[b] = (u8)(w[b >> 2] >> (8 * (b & 3)));
  }
}

// one rank's share of a sharded Fri::prove, cut into the phases a lock-step driver interleaves over the ranks
struct MgProve {
  stark_mgpu *m = nullptr;
  stark_ctx *ctx = nullptr;
  stark_fri_state *s = nullptr;
  GeoTables G = {nullptr, nullptr};
  RoundConsts C = {};
  ProofLayout L;
  size_t n = 0;
  u32 R = 0, Rs = 0, nq = 0;
  size_t arena_off[MG_MAX_ROUNDS + 1] = {};
  u64 *d_top = nullptr, *d_seed = nullptr;
  int rc = STARK_OK;

  u32 *arena(int g, size_t off) const { return reinterpret_cast<u32 *>(m->peer[g] + m->L.arena) + off; }
  u8 *proof_buf(int g) const { return m->peer[g] + m->L.proof; }

  MgExchange exchange(u32 r, int mode) const {
    MgExchange X;
    memset(&X, 0, sizeof X);
    X.world = m->world, X.rank = m->rank, X.mode = mode, X.epoch = mg_epoch(m, r);
    const size_t slot = m->L.slots + 32 * (size_t)MG_MAX_RANKS * r;
    for (int g = 0; g < m->world; g++) X.slot_peer[g] = m->peer[g] + slot, X.flag_peer[g] = mg_flags(m, g, 0);
    X.slot_local = m->win + slot, X.flag_local = mg_flags(m, m->rank, 0);
    X.err_local = reinterpret_cast<u32 *>(m->win + m->L.err);
    X.top_nodes = s->top_nodes[r];
    return X;
  }
  TranscriptArgs transcript(u32 r) const {
    return TranscriptArgs{s->d_tr, s->d_roots + 32 * r, r == R - 1 ? 0 : 1, s->d_alpha_raw + r, s->d_alpha_m + r};
  }

  // Fri::prove's checks (fri.rs:256-260, 183-192), allocations, and which rounds are sharded
  int begin(stark_mgpu *m_, const u32 *cw0, size_t n_, size_t domain_length, u32 offset, u32 omega, u32 ef, u32 nq_,
            const u8 *transcript_, size_t transcript_len, size_t proof_cap, size_t *proof_len) {
    m = m_, ctx = m_->ctx, n = n_, nq = nq_;
    mg_begin_op(m);
    if (n != domain_length)
      return rc = stark_fail(ctx, STARK_ERR_ARG, "initial codeword length does not match domain length");  // fri.rs:256-260
    if ((rc = fri_check(ctx, n, ef, &R, nq)) != STARK_OK) return rc;
    proof_layout(n, R, nq, &L);
    if (proof_len) *proof_len = L.total;
    if ((size_t)nq > 2 * L.last_len) return rc = stark_fail(ctx, STARK_ERR_ARG, "not enough entropy in indices wrt last codeword");
    if ((size_t)nq > L.last_len)
      return rc = stark_fail(ctx, STARK_ERR_ARG, "cannot sample more indices than available in last codeword; requested: %u, available: %zu", nq, L.last_len);
    if (proof_cap < L.total) return rc = stark_fail(ctx, STARK_ERR_ARG, "proof buffer too small: need %zu bytes", L.total);
    if (L.total + 8 * (size_t)nq + 16 > m->L.proof_cap) return rc = stark_fail(ctx, STARK_ERR_ARG, "proof larger than the group's window");
    if (L.n_cw > (u32)MAX_FRI_ROUNDS) return rc = stark_fail(ctx, STARK_ERR_ARG, "more than %d FRI rounds are not supported", MAX_FRI_ROUNDS);
    if ((rc = fri_commit_setup(ctx, cw0, n, offset, omega, ef, nq, transcript_, transcript_len, false, &s, &G, &C)) != STARK_OK) return rc;
    // sharded rounds: long enough to be throughput-bound, at least 4 leaves per rank, folded codewords fit the arena
    const size_t W = (size_t)m->world;
    size_t len = n, used = 0;
    Rs = 0;
    while (W > 1 && Rs < R && len >= ((size_t)1 << m->shard_log) && len / W >= 4) {
      if (Rs > 0) {
        if (used + len > m->L.arena_elems) break;
        arena_off[Rs] = used, used += len;
      }
      Rs++, len >>= 1;
    }
    s->sharded = Rs;
    for (u32 r = 0; r < Rs && rc == STARK_OK; r++) {
      u8 *t = nullptr;
      rc = dev_alloc(ctx, (void **)&t, 32 * (2 * W - 1));
      if (rc == STARK_OK) s->top_nodes.push_back(t);
    }
    if (rc == STARK_OK) rc = dev_alloc(ctx, (void **)&d_seed, 8);
    return rc;
  }

  // sharded round r, first half: [fold into this round, own range, stored into every replica + leaf hashes] -> subtree ->
  // subtree root into every window (+ in a fused group: wait, top levels, transcript)
  int round_signal(u32 r) {
    if (rc != STARK_OK) return rc;
    const size_t W = (size_t)m->world, len = C.len, per = len / W, lo = per * (size_t)m->rank;
    stark_tree *sub = nullptr;
    if ((rc = merkle_tree_alloc(ctx, per, &sub)) != STARK_OK) return rc;
    s->trees.push_back(sub);
    u32 lp = 0;
    for (size_t c = per; c > 1; c >>= 1) lp++;
    s->log_per[r] = lp;
    if (r == 0) {
      rc = merkle_leaves_dev(ctx, s->cw[0] + lo, per, 1, 1, 0, sub->nodes);   // fri.rs:118-121
    } else {
      s->cw.push_back(arena(m->rank, arena_off[r])), s->len.push_back(len), s->borrowed.push_back(true);
      FoldPeersN P;
      memset(&P, 0, sizeof P);
      P.n = m->world;
      for (int g = 0; g < m->world; g++) P.out[g] = arena(g, arena_off[r]);
      const u32 inv2off_m = ff::to_mont(ff::inv(ff::mul(2, C.off_prev)));
      if (ctx->prof_on) prof_begin(ctx, "mg_fold_leaf", 12ull * per + 32ull * per);
      cudaError_t e = launch_pdl(k_mg_fold_leaf, dim3((u32)((per / 2 + 63) / 64)), dim3(64), 0, ctx->stream, (const u32 *)s->cw[r - 1], P, len,
                                 lo, lo + per, (int)(r - 1), G, ff::to_mont(C.g_prev), (const u32 *)(s->d_alpha_m + (r - 1)), inv2off_m,
                                 sub->nodes);
      if (ctx->prof_on) prof_end(ctx);
      ctx->launches++;
      if (e != cudaSuccess) rc = stark_fail(ctx, STARK_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
      m->bytes_sent += 4ull * per * (W - 1);
    }
    if (rc != STARK_OK) return rc;
    const TranscriptArgs tr = transcript(r);
    const MgExchange X = exchange(r, m->lockstep ? MG_X_SIGNAL : MG_X_FUSED);
    rc = merkle_climb_dev(ctx, sub->nodes, per, &tr, &X);
    m->bytes_sent += 36ull * (W - 1);
    if (!m->lockstep) round_consts_next(C);
    return rc;
  }
  // second half for lock-step groups: wait for the peers' roots, top levels, transcript
  int round_wait(u32 r) {
    if (rc != STARK_OK || !m->lockstep) return rc;
    const TranscriptArgs tr = transcript(r);
    const MgExchange X = exchange(r, MG_X_WAIT);
    rc = merkle_mg_top_dev(ctx, &tr, &X);
    round_consts_next(C);
    return rc;
  }

  // the replicated rounds, index sampling and proof assembly; column roots (optional) ride the same final barrier
  int rest() {
    if (rc != STARK_OK) return rc;
    if ((rc = fri_commit_rounds(ctx, s, Rs, G, C)) != STARK_OK) return rc;
    u8 *d_proof = proof_buf(m->rank);
    d_top = reinterpret_cast<u64 *>(d_proof + ((L.total + 7) & ~(size_t)7));
    if (ctx->prof_on) prof_begin(ctx, "query_phase", 0);
    const size_t sample_size = L.n_cw > 1 ? s->len[1] : s->len[0];       // fri.rs:266-270
    cudaError_t qe = launch_pdl(k_sample_indices, dim3(1), dim3(256), 0, ctx->stream, (const TranscriptDev *)s->d_tr, d_seed,
                                (u64)sample_size, (u64)L.last_len, nq, d_top);  // fri.rs:272-276
    const size_t hdr_threads = L.last_len > R ? L.last_len : R;
    if (qe == cudaSuccess)
      qe = launch_pdl(k_proof_header, dim3((u32)((hdr_threads + 255) / 256)), dim3(256), 0, ctx->stream, d_proof,
                      (const u8 *)s->d_roots, R, (const u32 *)s->cw[L.n_cw - 1], (u64)L.last_len);
    ctx->launches += 2;
    if (L.n_cw > 1 && nq) {
      MgProofArgs M;
      memset(&M, 0, sizeof M);
      u32 d0 = 0;
      for (size_t c = s->len[0]; c > 1; c >>= 1) d0++;
      M.A.len0 = s->len[0], M.A.nq = nq, M.A.depth0 = d0;
      for (u32 i = 0; i < L.n_cw; i++) {
        M.A.cw[i] = s->cw[i], M.A.nodes[i] = s->trees[i]->nodes;
        M.log_per[i] = i < Rs ? s->log_per[i] : 0xffu;
        M.top[i] = i < Rs ? s->top_nodes[i] : nullptr;
      }
      for (u32 i = 0; i + 1 < L.n_cw; i++) M.A.out_off[i] = L.round_off[i];
      M.world = m->world, M.rank = m->rank;
      for (int g = 0; g < m->world; g++) M.out[g] = proof_buf(g);
      const size_t threads = (size_t)nq * (3 * d0 - 1);
      if (qe == cudaSuccess)
        qe = launch_pdl(k_mg_proof_rounds, dim3((u32)((threads + 127) / 128), L.n_cw - 1), dim3(128), 0, ctx->stream, M,
                        (const u64 *)d_top);
      ctx->launches++;
      // path nodes of the sharded levels: about a 1/G share of them goes to each of the G - 1 peers
      u64 sharded_hashes = 0;
      for (u32 i = 0; i < Rs; i++) sharded_hashes += (u64)nq * s->log_per[i] * (i == 0 ? 2 : 3);
      m->bytes_sent += 32ull * sharded_hashes * (u64)(m->world - 1) / (u64)m->world;
    }
    if (ctx->prof_on) prof_end(ctx);
    if (qe != cudaSuccess) rc = stark_fail(ctx, STARK_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(qe));
    return rc;
  }
  u32 final_epoch() const { return mg_epoch(m, MG_MAX_ROUNDS + 1); }
  int finish_signal() {
    if (rc != STARK_OK) return rc;
    return rc = m->lockstep ? mg_barrier_signal(m, 1, final_epoch()) : mg_barrier(m, 1, final_epoch());
  }
  int finish_wait() {
    if (rc != STARK_OK || !m->lockstep) return rc;
    return rc = mg_barrier_wait(m, 1, final_epoch());
  }
  // proof bytes (and optionally the top-level indices) to the host; frees the state
  int download(u8 *proof, u64 *top_indices) {
    if (rc == STARK_OK && proof &&
        cudaMemcpyAsync(proof, proof_buf(m->rank), L.total, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
      rc = stark_fail(ctx, STARK_ERR_CUDA, "D2H copy failed");
    if (rc == STARK_OK && top_indices && nq &&
        cudaMemcpyAsync(top_indices, d_top, 8 * (size_t)nq, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
      rc = stark_fail(ctx, STARK_ERR_CUDA, "D2H copy failed");
    return rc;
  }
  void release() {
    if (ctx) dev_free(ctx, d_seed);
    d_seed = nullptr;
    fri_state_free(s);
    s = nullptr;
  }
};

// Runs `world` MgProve objects through their phases.  A multi-process rank passes its own single object; a local group
// passes all of them and the phases are interleaved over the ranks (lock step when they share a device).
template <typename PerRank>
static void mg_each(MgProve *P, int count, PerRank fn) {
  for (int k = 0; k < count; k++) {
    mg_use(P[k].m);
    fn(P[k]);
  }
}
static int mg_prove_run(MgProve *P, int count) {
  u32 Rs = 0;
  for (int k = 0; k < count; k++) Rs = P[k].Rs > Rs ? P[k].Rs : Rs;
  auto failed = [&]() {
    for (int k = 0; k < count; k++)
      if (P[k].rc != STARK_OK) return P[k].rc;
    return (int)STARK_OK;
  };
  // a rank of this call that failed stops all of them (the others' queued waits give up after MG_TIMEOUT_NS)
  for (u32 r = 0; r < Rs && failed() == STARK_OK; r++) {
    mg_each(P, count, [&](MgProve &p) { p.round_signal(r); });
    if (failed() != STARK_OK) break;
    mg_each(P, count, [&](MgProve &p) { p.round_wait(r); });
  }
  if (failed() == STARK_OK) mg_each(P, count, [&](MgProve &p) { p.rest(); });
  return failed();
}
static int mg_prove_finish(MgProve *P, int count) {
  mg_each(P, count, [&](MgProve &p) { p.finish_signal(); });
  mg_each(P, count, [&](MgProve &p) { p.finish_wait(); });
  int rc = STARK_OK;
  for (int k = 0; k < count; k++)
    if (P[k].rc != STARK_OK) rc = P[k].rc;
  return rc;
}

// ---- the columns that are not FRI-proved, on their own stream (shared by the single-GPU and the sharded config-3
// pipelines).  Their LDE and trees are throughput work that does not depend on the FRI, whose rounds are a latency-bound
// chain: queued on the column stream they fill the SMs the FRI leaves idle.  Columns go through in GROUPS:
// one batched LDE per group and ONE set of tree launches for all trees of the group (merkle_build_batch_dev), so the
// latency-bound top levels of the trees overlap instead of queueing behind one another.  With host input every group is
// copied on a third stream (ctx->side[3]) while the groups before it compute: only the first copy is exposed.
// The climb kernel's last-CTA tickets are per stream (TICKET_SIDE for the column stream).
// Groups ALTERNATE between two column streams: the end of a group's tree build (the small levels and the climb, ~0.1 ms
// of latency-bound launches) then overlaps the next group's LDE and leaf hashing instead of stalling the stream
// (host-input traces go through 4 groups: 7.45 -> 7.1x ms end to end).
constexpr int COL_STREAM = 2, COPY_STREAM = 3, COL_STREAM2 = 4;
struct ColumnPipe {
  cudaStream_t main_stream = nullptr;
  bool forked = false;
  u8 *nodes = nullptr;              // the trees of all committed columns, tree_stride bytes apart
  size_t trees_reserved = 0;
  std::vector<cudaEvent_t> events;  // copy-done events (host input)
  u64 *staging = nullptr;           // host input: the uint64 trace as copied (column-major), narrowed group by group
  u8 *d_roots = nullptr;            // one root per committed column, in the order committed
  u32 n_roots = 0;
};
static size_t tree_stride_of(size_t N) { return ((2 * N - 1) * 32 + 255) & ~(size_t)255; }

// route the context's launches to the column stream / back (the launch macros use ctx->stream)
static void column_pipe_enter(stark_ctx *ctx, ColumnPipe *cp, int which = 0) {
  if (!cp->forked) return;
  if (which & 1)
    ctx->stream = ctx->side[COL_STREAM2], ctx->climb_counter = ctx->flag + TICKET_SIDE2;
  else
    ctx->stream = ctx->side[COL_STREAM], ctx->climb_counter = ctx->flag + TICKET_SIDE;
}
static void column_pipe_leave(stark_ctx *ctx, ColumnPipe *cp) {
  if (cp->forked) ctx->stream = cp->main_stream, ctx->climb_counter = ctx->flag + TICKET_MAIN;
}
// fork the column (and copy) stream from the context's stream: everything allocated or written so far is visible there
static int column_pipe_begin(stark_ctx *ctx, ColumnPipe *cp, u32 max_roots, size_t N) {
  cp->main_stream = ctx->stream;
  if (max_roots) {
    ST_TRY(dev_alloc(ctx, (void **)&cp->d_roots, 32 * (size_t)max_roots));
    ST_TRY(dev_alloc(ctx, (void **)&cp->nodes, tree_stride_of(N) * max_roots));
    cp->trees_reserved = max_roots;
  }
  if (!ctx->prof_on && !ctx->colpipe_serial && side_streams(ctx, 5) == STARK_OK) {
    cudaEventRecord(ctx->fork_ev, cp->main_stream);
    cudaStreamWaitEvent(ctx->side[COL_STREAM], ctx->fork_ev, 0);
    cudaStreamWaitEvent(ctx->side[COL_STREAM2], ctx->fork_ev, 0);
    cudaStreamWaitEvent(ctx->side[COPY_STREAM], ctx->fork_ev, 0);
    cp->forked = true;
  }
  return STARK_OK;
}
// queue the H2D copy of n_elems host values into the staging buffer at element offset dst_off (copy stream); when
// `record` the event to wait for is returned (one per group, after the group's last copy)
static int column_pipe_copy(stark_ctx *ctx, ColumnPipe *cp, const uint64_t *src, size_t dst_off, size_t n_elems, bool record,
                            cudaEvent_t *ev) {
  cudaStream_t st = cp->forked ? ctx->side[COPY_STREAM] : ctx->stream;
  CU_TRY(ctx, cudaMemcpyAsync(cp->staging + dst_off, src, 8 * n_elems, cudaMemcpyHostToDevice, st));
  if (ev) *ev = nullptr;
  if (cp->forked && record && ev) {
    CU_TRY(ctx, cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
    cp->events.push_back(*ev);
    CU_TRY(ctx, cudaEventRecord(*ev, st));
  }
  return STARK_OK;
}
// On the CURRENT stream of the context: [wait for the group's copy, narrow it] -> [LDE of the group] -> batched trees ->
// roots appended to d_roots.  cols_dev: the u32 columns (source of the LDE; written here when the input is on the host).
static int column_pipe_group(stark_ctx *ctx, ColumnPipe *cp, cudaEvent_t copied, u32 *cols_dev, u32 *lde, u32 c0, u32 cnt,
                             u32 log_n, u32 log_blowup, u32 offset, bool do_lde, bool do_trees) {
  const size_t n = (size_t)1 << log_n, N = n << log_blowup;
  if (cnt == 0) return STARK_OK;
  if (copied) CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, copied, 0));
  if (cp->staging) ST_TRY(narrow_dev(ctx, cp->staging + (size_t)c0 * n, n * cnt, cols_dev + (size_t)c0 * n));
  if (do_lde) ST_TRY(lde_dev(ctx, cols_dev + (size_t)c0 * n, cnt, log_n, log_blowup, offset, lde + (size_t)c0 * N));
  for (u32 b0 = 0; do_trees && b0 < cnt; b0 += CLIMB_TICKETS) {
    const u32 nb = cnt - b0 < (u32)CLIMB_TICKETS ? cnt - b0 : (u32)CLIMB_TICKETS;
    const size_t stride = tree_stride_of(N);
    // the trees of all groups live in ONE block reserved on the context's stream before the fork (column_pipe_begin):
    // every stream-ordered allocation and free of the pipeline's big buffers happens on the same stream
    if ((size_t)(cp->n_roots + nb) > cp->trees_reserved) return stark_fail(ctx, STARK_ERR_ARG, "column pipeline: tree block too small");
    u8 *nodes = cp->nodes + stride * cp->n_roots;
    ST_TRY(merkle_build_batch_dev(ctx, lde + (size_t)(c0 + b0) * N, N, nb, N, nodes, stride));
    CU_TRY(ctx, cudaMemcpy2DAsync(cp->d_roots + 32 * (size_t)cp->n_roots, 32, nodes + 32 * (2 * N - 2), stride, 32, nb,
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    cp->n_roots += nb;
  }
  return STARK_OK;
}
// the context's stream waits for the column stream (also on an error path: the buffers are freed afterwards)
static void column_pipe_join(stark_ctx *ctx, ColumnPipe *cp) {
  if (cp->forked) {
    cudaEventRecord(ctx->side_done[COL_STREAM], ctx->side[COL_STREAM]);
    cudaStreamWaitEvent(cp->main_stream, ctx->side_done[COL_STREAM], 0);
    cudaEventRecord(ctx->side_done[COL_STREAM2], ctx->side[COL_STREAM2]);
    cudaStreamWaitEvent(cp->main_stream, ctx->side_done[COL_STREAM2], 0);
    cudaEventRecord(ctx->side_done[COPY_STREAM], ctx->side[COPY_STREAM]);
    cudaStreamWaitEvent(cp->main_stream, ctx->side_done[COPY_STREAM], 0);
  }
  ctx->stream = cp->main_stream, ctx->climb_counter = ctx->flag + TICKET_MAIN;
  cp->forked = false;
}
static void column_pipe_free(stark_ctx *ctx, ColumnPipe *cp) {
  column_pipe_join(ctx, cp);
  dev_free(ctx, cp->nodes);
  cp->nodes = nullptr;
  for (cudaEvent_t e : cp->events) cudaEventDestroy(e);
  cp->events.clear();
  dev_free(ctx, cp->d_roots), dev_free(ctx, cp->staging);
  cp->d_roots = nullptr, cp->staging = nullptr;
}

// For the duration of an operation the context's launches go to its high-priority stream (forked from the caller's stream,
// joined back to it at the end).  The FRI rounds are a chain of short latency-bound kernels; while the column stream has
// thousands of throughput CTAs pending, a default-priority launch of the chain waits for CTA slots behind them: measured
// on 8 GPUs, 2 column trees per rank added 0.7 ms to a 1.0 ms chain that should have hidden them (profiles/r2e_bench_g8.json).
struct PrioScope {
  stark_ctx *ctx = nullptr;
  cudaStream_t user = nullptr;
  bool on = false, pdl_was_off = false;
  void enter(stark_ctx *c) {
    ctx = c, user = c->stream;
    if (c->prof_on || c->colpipe_serial || c->no_prio) return;
    if (!c->prio_stream) {
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);   // hi = numerically smallest = greatest priority
      if (cudaStreamCreateWithPriority(&c->prio_stream, cudaStreamNonBlocking, hi) != cudaSuccess ||
          cudaEventCreateWithFlags(&c->prio_ev, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        c->prio_stream = nullptr;
        return;
      }
    }
    cudaEventRecord(c->prio_ev, user);
    cudaStreamWaitEvent(c->prio_stream, c->prio_ev, 0);
    c->stream = c->prio_stream, on = true;
    pdl_was_off = t_pdl_off, t_pdl_off = !c->keep_pdl;
  }
  void leave() {
    if (!on) return;
    t_pdl_off = pdl_was_off;
    cudaEventRecord(ctx->prio_ev, ctx->prio_stream);
    cudaStreamWaitEvent(user, ctx->prio_ev, 0);
    ctx->stream = user, on = false;
  }
  ~PrioScope() { leave(); }
};

// The groups local columns [1, n_my) go through the column stream in: (first column, count).  On the device already: one
// group (the widest batches).  From the host: 1, 2, then `gs` columns -- the column stream starts as soon as ONE column has
// arrived instead of idling until a full group is there, and the compute of a group hides the copy of the next.
static std::vector<std::pair<u32, u32>> column_groups(u32 n_my, bool from_host, u32 gs) {
  std::vector<std::pair<u32, u32>> g;
  u32 c0 = 1, step = from_host ? 1u : (n_my > 1 ? n_my - 1 : 1u);
  while (c0 < n_my) {
    const u32 cnt = n_my - c0 < step ? n_my - c0 : step;
    g.push_back({c0, cnt});
    c0 += cnt;
    if (from_host) step = step * 2 < gs ? step * 2 : gs;
  }
  return g;
}

// BASELINE config 3 on one device, from device columns (cols_dev) or from the host trace (host_cols; cols_dev then is the
// buffer the narrowed columns go to): column 0 -> LDE -> Fri::prove on the context's stream, the other columns ->
// LDE + trees on the column stream.
static int prove_trace_pipeline(stark_ctx *ctx, const uint64_t *host_cols, u32 *cols_dev, uint32_t n_cols, uint32_t log_n,
                                uint32_t log_blowup, uint32_t offset, uint32_t nq, uint8_t *column_roots, uint8_t *proof,
                                size_t proof_cap, size_t *proof_len) {
  const size_t n = (size_t)1 << log_n, N = n << log_blowup;
  u32 fri_rounds = 0;
  ST_TRY(fri_check(ctx, N, 1u << log_blowup, &fri_rounds, nq));
  // With num_rounds() == 0 (N <= expansion factor or N <= 4 nq, fri.rs:93-103) Fri::commit builds no tree and the proof
  // starts with the last codeword, so column 0 is committed like the others.
  const bool tree0 = fri_rounds == 0;
  u32 *lde = nullptr;
  ColumnPipe cp;
  PrioScope prio;
  if (n_cols > 1) prio.enter(ctx);     // with one column nothing competes with the chain
  ST_TRY(dev_alloc(ctx, (void **)&lde, N * n_cols * 4));
  int rc = STARK_OK;
  if (host_cols) {
    rc = dev_alloc(ctx, (void **)&cp.staging, 8 * n * n_cols);
    if (rc == STARK_OK) rc = upload_flag_reset(ctx);
  }
  if (rc == STARK_OK) rc = column_pipe_begin(ctx, &cp, tree0 ? n_cols : n_cols - 1, N);
  // groups: column 0 alone (the FRI waits for nothing else), then the rest (column_groups)
  const auto groups = column_groups(n_cols, host_cols != nullptr, (u32)ctx->colpipe_group);
  std::vector<cudaEvent_t> ev(1 + groups.size(), nullptr);
  if (host_cols) {
    rc = column_pipe_copy(ctx, &cp, host_cols, 0, n, true, &ev[0]);
    for (size_t gi = 0; gi < groups.size() && rc == STARK_OK; gi++)
      rc = column_pipe_copy(ctx, &cp, host_cols + (size_t)groups[gi].first * n, (size_t)groups[gi].first * n,
                            n * groups[gi].second, true, &ev[1 + gi]);
  }
  // STARK_TRACE_PIPE=1: device timestamps of the pipeline's milestones (diagnosis; printed to stderr after the call)
  std::vector<std::pair<const char *, cudaEvent_t>> marks;
  auto mark = [&](const char *what, cudaStream_t st) {
    if (!ctx->trace_pipe) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    marks.push_back({what, e});
  };
  mark("start", ctx->stream);
  if (host_cols && cp.forked) mark("all copies done", ctx->side[COPY_STREAM]);
  // column 0 and the whole latency chain are QUEUED FIRST (no host synchronisation inside): the ~60 launches of the
  // chain are on the critical path, the ~25 launches per column group are not -- issued in the other order the chain
  // started half a millisecond of host launch time late (a constant 0.9 ms between the host-input and the
  // device-resident call at every group size, profiles/r2h_bench_g*.json)
  if (rc == STARK_OK) rc = column_pipe_group(ctx, &cp, ev[0], cols_dev, lde, 0, 1, log_n, log_blowup, offset, true, tree0);
  const u32 omega = ff::pow(ff::GEN, (ff::P - 1) >> (log_n + log_blowup));  // prim_nth_root(N), ff.rs:215-223
  mark("column 0 LDE done", ctx->stream);
  if (rc == STARK_OK)
    rc = fri_prove_dev(ctx, lde, N, N, offset, omega, 1u << log_blowup, nq, nullptr, 0, proof, proof_cap, proof_len, nullptr, false);
  mark("chain done", ctx->stream);
  // the other columns on the column stream
  for (size_t gi = 0; gi < groups.size() && rc == STARK_OK; gi++) {
    column_pipe_enter(ctx, &cp, (int)gi);
    rc = column_pipe_group(ctx, &cp, ev[1 + gi], cols_dev, lde, groups[gi].first, groups[gi].second, log_n, log_blowup, offset, true, true);
    mark("column group done", ctx->stream);
    column_pipe_leave(ctx, &cp);
  }
  column_pipe_join(ctx, &cp);
  mark("joined", ctx->stream);
  // roots: d_roots holds [column 0 if tree0] followed by columns 1..
  const u32 first_col = tree0 ? 0u : 1u;
  if (rc == STARK_OK && column_roots && cp.n_roots &&
      cudaMemcpyAsync(column_roots + 32 * first_col, cp.d_roots, 32 * (size_t)cp.n_roots, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
    rc = stark_fail(ctx, STARK_ERR_CUDA, "D2H copy failed");
  if (rc == STARK_OK && host_cols) rc = upload_flag_fetch(ctx);
  mark("end", ctx->stream);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && rc == STARK_OK)
    rc = stark_fail(ctx, STARK_ERR_CUDA, "prove failed: %s", cudaGetErrorString(cudaGetLastError()));
  for (size_t i = 0; i < marks.size(); i++) {
    float ms = 0;
    cudaEventSynchronize(marks[i].second);
    cudaEventElapsedTime(&ms, marks[0].second, marks[i].second);
    fprintf(stderr, "[pipe] %-22s %8.3f ms\n", marks[i].first, ms);
    if (i) cudaEventDestroy(marks[i].second);
  }
  if (!marks.empty()) cudaEventDestroy(marks[0].second);
  if (rc == STARK_OK && column_roots && !tree0) memcpy(column_roots, proof + 1, 32);  // first object = root of column 0
  if (rc == STARK_OK && host_cols) rc = upload_u64_check(ctx);
  column_pipe_free(ctx, &cp);
  dev_free(ctx, lde);
  return rc;
}

// which trace columns a rank commits in the sharded config 3: column 0 is LDE'd by every rank (its codeword is the FRI
// input, replicated without a broadcast) and committed by the sharded FRI itself; columns 1.. go round robin
static u32 mg_owned_columns(int rank, int world, u32 n_cols, u32 *out) {
  u32 k = 0;
  for (u32 c = 1; c < n_cols; c++)
    if ((int)((c - 1) % (u32)world) == rank) {
      if (out) out[k] = c;
      k++;
    }
  return k;
}

struct MgTraceRank {
  ColumnPipe cp;
  u32 *lde = nullptr;
  stark_buf *in = nullptr;       // host-input variant: this rank's columns on the device
  std::vector<u32> owned;        // global indices of the committed columns
  u32 n_my = 0;
  std::vector<std::pair<u32, u32>> groups;   // the column groups, queued after the chain
  std::vector<cudaEvent_t> ev;
  u32 *cols_dev = nullptr;
  bool bcast0 = false;             // column 0 came through the window: its canonical flag travels with it
  u32 bcast_flag = 0;
};

static int mg_prove_trace_impl(stark_mgpu *const *ranks, int n_here, const uint64_t *host_cols, const stark_buf *const *my_cols,
                               uint32_t n_cols, uint32_t log_n, uint32_t log_blowup, uint64_t offset, uint32_t nq,
                               uint8_t *const *column_roots, uint8_t *const *proofs, size_t proof_cap, size_t *proof_len) {
  if (!ranks || n_here < 1 || n_cols == 0 || (!host_cols && !my_cols)) return stark_fail(nullptr, STARK_ERR_ARG, "null argument");
  for (int k = 0; k < n_here; k++)
    if (!ranks[k] || (my_cols && !my_cols[k]) || !proofs || !proofs[k]) return stark_fail(nullptr, STARK_ERR_ARG, "null argument");
  stark_ctx *ctx0 = ranks[0]->ctx;
  if (ranks[0]->mode == MG_LOCAL && n_here != ranks[0]->world)
    return stark_fail(ctx0, STARK_ERR_ARG, "a local group is driven with all of its ranks in one call");
  if (log_n + log_blowup > (u32)ff::TWO_ADICITY) return stark_fail(ctx0, STARK_ERR_ARG, "n > 2^23 not supported by this modulus");
  if (offset == 0 || offset >= ff::P) return stark_fail(ctx0, STARK_ERR_ARG, "offset must be a non-zero canonical element");
  if (n_cols > ranks[0]->L.max_cols / 2) return stark_fail(ctx0, STARK_ERR_ARG, "more columns than the group's window holds");
  const size_t n = (size_t)1 << log_n, N = n << log_blowup;
  const u32 omega = ff::pow(ff::GEN, (ff::P - 1) >> (log_n + log_blowup));  // prim_nth_root(N), ff.rs:215-223
  u32 fri_rounds = 0;
  ST_TRY(fri_check(ctx0, N, 1u << log_blowup, &fri_rounds, nq));
  std::vector<MgTraceRank> T(n_here);
  std::vector<MgProve> P(n_here);
  std::vector<PrioScope> prio(n_here);
  int rc = STARK_OK;
  // STARK_TRACE_PIPE=1: device timestamps of the first driven rank's milestones (diagnosis)
  std::vector<std::pair<const char *, cudaEvent_t>> marks;
  auto mark = [&](int k, const char *what, cudaStream_t st) {
    if (k != 0 || !ranks[0]->ctx->trace_pipe) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    marks.push_back({what, e});
  };
  // phase 1 (no exchange): upload / LDE of column 0 and the owned columns, column trees on the side stream
  for (int k = 0; k < n_here && rc == STARK_OK; k++) {
    stark_mgpu *m = ranks[k];
    stark_ctx *ctx = m->ctx;
    mg_use(m);
    MgTraceRank &t = T[k];
    t.owned.resize(n_cols);
    t.owned.resize(mg_owned_columns(m->rank, m->world, n_cols, t.owned.data()));
    t.n_my = 1 + (u32)t.owned.size();
    // ranks that share a stream (virtual ranks in lock step) stay on it: their order IS the stream order
    if (!m->lockstep && n_cols > 1) prio[k].enter(ctx);
    mark(k, "start", ctx->stream);
    u32 *cols_dev = nullptr;
    if (host_cols) {
      rc = stark_buf_alloc(ctx, n * t.n_my, &t.in);
      if (rc == STARK_OK) cols_dev = t.in->ptr;
      if (rc == STARK_OK) rc = dev_alloc(ctx, (void **)&t.cp.staging, 8 * n * t.n_my);
      if (rc == STARK_OK) rc = upload_flag_reset(ctx);
    } else {
      if (my_cols[k]->n < n * t.n_my) rc = stark_fail(ctx, STARK_ERR_ARG, "buffer too small");
      cols_dev = my_cols[k]->ptr;
    }
    if (rc == STARK_OK) rc = dev_alloc(ctx, (void **)&t.lde, N * t.n_my * 4);
    // local column 0 = trace column 0 (LDE on the context's stream: the FRI input); local columns 1.. = the owned ones,
    // LDE + batched trees on the column stream, in groups of 4 when they are copied from the host.  With
    // num_rounds() == 0 Fri::commit builds no tree, so rank 0 commits column 0 as well.
    const bool tree0 = fri_rounds == 0 && m->rank == 0;
    if (rc == STARK_OK) rc = column_pipe_begin(ctx, &t.cp, tree0 ? t.n_my : t.n_my - 1, N);
    const auto groups = column_groups(t.n_my, host_cols != nullptr, (u32)ctx->colpipe_group);
    std::vector<cudaEvent_t> ev(1 + groups.size(), nullptr);
    // Column 0 is needed by every rank.  From the host it is copied ONCE, by rank 0, and broadcast over NVLink
    // (mg_bcast_column): with all ranks copying at the same time the host links are the scarce resource (8 ranks x 3
    // columns took 2.5x as long per byte as one rank alone), and column 0 heads everybody's critical path.
    const bool bcast0 = host_cols && m->world > 1 && n % 4 == 0 && n <= m->L.arena_elems / 2 && !ctx->no_bcast0;
    auto copy_owned = [&]() {
      for (size_t gi = 0; gi < groups.size() && rc == STARK_OK; gi++)
        for (u32 i = groups[gi].first; i < groups[gi].first + groups[gi].second && rc == STARK_OK; i++)
          rc = column_pipe_copy(ctx, &t.cp, host_cols + (size_t)t.owned[i - 1] * n, (size_t)i * n, n,
                                i + 1 == groups[gi].first + groups[gi].second, &ev[1 + gi]);
    };
    if (host_cols && rc == STARK_OK) {
      if (!bcast0 || m->rank == 0) rc = column_pipe_copy(ctx, &t.cp, host_cols, 0, n, true, &ev[0]);
      if (!bcast0) copy_owned();
    }
    if (rc == STARK_OK && bcast0) {
      // rank 0: wait for its copy, narrow, store the column into every window, raise the flag; the others: wait for it.
      // Every rank then runs the LDE of column 0 out of its own window.
      if (m->rank == 0) {
        if (ev[0] && cudaStreamWaitEvent(ctx->stream, ev[0], 0) != cudaSuccess) rc = stark_fail(ctx, STARK_ERR_CUDA, "stream wait failed");
        if (rc == STARK_OK) rc = narrow_dev(ctx, t.cp.staging, n, cols_dev);
      }
      mark(k, "column 0 narrowed", ctx->stream);
      if (rc == STARK_OK) rc = mg_bcast_column(m, 0, cols_dev, n, mg_epoch(m, MG_MAX_ROUNDS + 2));
      mark(k, "column 0 broadcast", ctx->stream);
      // The host links are shared (8 ranks copying at once get 22-35 GB/s each instead of 53): the copies of the other
      // columns start only once column 0 -- which heads every rank's critical path -- has arrived everywhere
      if (rc == STARK_OK && t.cp.forked) {
        cudaEvent_t arrived = nullptr;
        if (cudaEventCreateWithFlags(&arrived, cudaEventDisableTiming) == cudaSuccess) {
          t.cp.events.push_back(arrived);
          cudaEventRecord(arrived, ctx->stream);
          cudaStreamWaitEvent(ctx->side[COPY_STREAM], arrived, 0);
        }
      }
      copy_owned();
      u64 *staging = t.cp.staging;
      t.cp.staging = nullptr;     // column_pipe_group: no narrowing, the column is already u32
      if (rc == STARK_OK)
        rc = column_pipe_group(ctx, &t.cp, nullptr, mg_bcast_ptr(m, m->rank), t.lde, 0, 1, log_n, log_blowup, (u32)offset, true, tree0);
      t.cp.staging = staging;
      t.bcast0 = true;
    } else if (rc == STARK_OK) {
      rc = column_pipe_group(ctx, &t.cp, ev[0], cols_dev, t.lde, 0, 1, log_n, log_blowup, (u32)offset, true, tree0);
    }
    t.groups = groups, t.ev = ev, t.cols_dev = cols_dev;
    if (host_cols && t.cp.forked) mark(k, "all copies done", ctx->side[COPY_STREAM]);
    mark(k, "column 0 LDE done", ctx->stream);
    if (rc == STARK_OK)
      rc = P[k].begin(m, t.lde, N, N, (u32)offset, omega, 1u << log_blowup, nq, nullptr, 0, proof_cap, proof_len);
    else
      mg_begin_op(m);   // keep the operation count in step with the ranks that did begin
  }
  // phase 2: the sharded Fri::prove on column 0 -- queued BEFORE the column groups (see prove_trace_pipeline)
  if (rc == STARK_OK) rc = mg_prove_run(P.data(), n_here);
  mark(0, "chain done", ranks[0]->ctx->stream);
  for (int k = 0; k < n_here && rc == STARK_OK; k++) {
    stark_mgpu *m = ranks[k];
    stark_ctx *ctx = m->ctx;
    mg_use(m);
    MgTraceRank &t = T[k];
    for (size_t gi = 0; gi < t.groups.size() && rc == STARK_OK; gi++) {
      column_pipe_enter(ctx, &t.cp, (int)gi);
      rc = column_pipe_group(ctx, &t.cp, t.ev[1 + gi], t.cols_dev, t.lde, t.groups[gi].first, t.groups[gi].second, log_n, log_blowup,
                             (u32)offset, true, true);
      mark(k, "column group done", ctx->stream);
      column_pipe_leave(ctx, &t.cp);
    }
  }
  // phase 3: the column roots into every rank's table, then the barrier that ends the operation
  for (int k = 0; k < n_here && rc == STARK_OK; k++) {
    stark_mgpu *m = ranks[k];
    mg_use(m);
    MgTraceRank &t = T[k];
    column_pipe_join(m->ctx, &t.cp);
    std::vector<u32> idx;
    if (fri_rounds == 0 && m->rank == 0) idx.push_back(0);
    for (u32 c : t.owned) idx.push_back(c);
    if (!idx.empty()) rc = mg_put_roots(m, t.cp.d_roots, idx.data(), (u32)idx.size());
    if (rc == STARK_OK && host_cols) rc = upload_flag_fetch(m->ctx);
  }
  if (rc == STARK_OK) rc = mg_prove_finish(P.data(), n_here);
  mark(0, "final barrier done", ranks[0]->ctx->stream);
  for (int k = 0; k < n_here; k++) {
    stark_mgpu *m = ranks[k];
    stark_ctx *ctx = m->ctx;
    mg_use(m);
    if (rc == STARK_OK) rc = P[k].download(proofs[k], nullptr);
    if (rc == STARK_OK && T[k].bcast0 &&
        cudaMemcpyAsync(&T[k].bcast_flag, mg_bcast_ptr(m, m->rank) + n, 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
      rc = stark_fail(ctx, STARK_ERR_CUDA, "D2H copy failed");
    if (rc == STARK_OK && column_roots && column_roots[k] &&
        cudaMemcpyAsync(column_roots[k], mg_colroots(m, m->rank), 32 * (size_t)n_cols, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
      rc = stark_fail(ctx, STARK_ERR_CUDA, "D2H copy failed");
  }
  for (int k = 0; k < n_here; k++) {
    stark_mgpu *m = ranks[k];
    stark_ctx *ctx = m->ctx;
    mg_use(m);
    const int e = mg_check_err(m);   // synchronises the rank's stream
    if (rc == STARK_OK) rc = e;
    if (k == 0 && !marks.empty()) {
      for (size_t i = 0; i < marks.size(); i++) {
        float ms = 0;
        cudaEventSynchronize(marks[i].second);
        cudaEventElapsedTime(&ms, marks[0].second, marks[i].second);
        fprintf(stderr, "[pipe rank %d] %-22s %8.3f ms\n", m->rank, marks[i].first, ms);
      }
      for (auto &mk : marks) cudaEventDestroy(mk.second);
      marks.clear();
    }
    if (rc == STARK_OK && fri_rounds > 0 && column_roots && column_roots[k]) memcpy(column_roots[k], proofs[k] + 1, 32);  // root of column 0
    if (rc == STARK_OK && host_cols) rc = upload_u64_check(ctx);
    if (rc == STARK_OK && T[k].bcast0 && T[k].bcast_flag)
      rc = stark_fail(ctx, STARK_ERR_ARG, "non-canonical field element (value >= p) in input");
    P[k].release();
    column_pipe_free(ctx, &T[k].cp);
    dev_free(ctx, T[k].lde);
    stark_buf_free(T[k].in);
    prio[k].leave();
  }
  return rc;
}

// ---- BASELINE config 5 on a group: ONE Fri::commit round (fri.rs:116-147) of a replicated codeword -- sharded leaf hashes
// and subtree, root exchange, alpha, sharded fold stored into every replica.  Returns views of the folded replicas.
static int mg_fold_commit_round_impl(stark_mgpu *const *ranks, int n_here, const stark_buf *const *codewords, size_t n,
                                     u32 offset, u32 omega, uint8_t *const *roots, uint64_t *alpha_raw, stark_buf **folded) {
  stark_ctx *ctx0 = ranks[0]->ctx;
  const size_t W = (size_t)ranks[0]->world;
  if (n < 8 * W || (n & (n - 1))) return stark_fail(ctx0, STARK_ERR_ARG, "codeword length must be a power of two, at least 8 per rank");
  if (n / 2 > ranks[0]->L.arena_elems) return stark_fail(ctx0, STARK_ERR_ARG, "codeword larger than the group's window (max_codeword)");
  if (offset == 0 || omega == 0) return stark_fail(ctx0, STARK_ERR_ARG, "no division by zero");   // ff.rs:182
  struct Rk {
    stark_tree *sub = nullptr;
    u8 *top = nullptr, *d_root = nullptr;
    u64 *d_alpha_raw = nullptr;
    u32 *d_alpha_m = nullptr;
    TranscriptDev *d_tr = nullptr;
    GeoTables G = {nullptr, nullptr};
  };
  std::vector<Rk> K(n_here);
  const size_t per = n / W, h = n / 2, hper = h / W;
  const u32 g0 = ff::inv(omega);
  int rc = STARK_OK;
  auto exchange = [&](stark_mgpu *m, Rk &k, int mode) {
    MgExchange X;
    memset(&X, 0, sizeof X);
    X.world = m->world, X.rank = m->rank, X.mode = mode, X.epoch = mg_epoch(m, 0);
    for (int g = 0; g < m->world; g++) X.slot_peer[g] = m->peer[g] + m->L.slots, X.flag_peer[g] = mg_flags(m, g, 0);
    X.slot_local = m->win + m->L.slots, X.flag_local = mg_flags(m, m->rank, 0);
    X.err_local = reinterpret_cast<u32 *>(m->win + m->L.err);
    X.top_nodes = k.top;
    return X;
  };
  // leaf hashes + subtree + root into every window (+ wait, top, alpha when fused)
  for (int i = 0; i < n_here && rc == STARK_OK; i++) {
    stark_mgpu *m = ranks[i];
    stark_ctx *ctx = m->ctx;
    mg_use(m);
    mg_begin_op(m);
    Rk &k = K[i];
    if (!codewords[i] || codewords[i]->n < n) rc = stark_fail(ctx, STARK_ERR_ARG, "buffer too small");
    if (rc == STARK_OK) rc = merkle_tree_alloc(ctx, per, &k.sub);
    if (rc == STARK_OK) rc = dev_alloc(ctx, (void **)&k.top, 32 * (2 * W - 1));
    if (rc == STARK_OK) rc = dev_alloc(ctx, (void **)&k.d_root, 32);
    if (rc == STARK_OK) rc = dev_alloc(ctx, (void **)&k.d_alpha_raw, 8);
    if (rc == STARK_OK) rc = dev_alloc(ctx, (void **)&k.d_alpha_m, 4);
    if (rc == STARK_OK) rc = dev_alloc(ctx, (void **)&k.d_tr, sizeof(TranscriptDev));
    if (rc == STARK_OK) rc = geo_tables(ctx, g0, 1, h, &k.G);
    if (rc != STARK_OK) break;
    TranscriptDev t;
    tr_init(t);
    if (cudaMemcpyAsync(k.d_tr, &t, sizeof t, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = stark_fail(ctx, STARK_ERR_CUDA, "H2D copy failed");
    if (rc == STARK_OK) rc = merkle_leaves_dev(ctx, codewords[i]->ptr + per * (size_t)m->rank, per, 1, 1, 0, k.sub->nodes);
    const TranscriptArgs tr = {k.d_tr, k.d_root, 1, k.d_alpha_raw, k.d_alpha_m};
    const MgExchange X = exchange(m, k, W == 1 ? MG_X_FUSED : (m->lockstep ? MG_X_SIGNAL : MG_X_FUSED));
    if (rc == STARK_OK) rc = merkle_climb_dev(ctx, k.sub->nodes, per, &tr, W > 1 ? &X : nullptr);
    m->bytes_sent += 36ull * (W - 1);
  }
  for (int i = 0; i < n_here && rc == STARK_OK; i++) {
    stark_mgpu *m = ranks[i];
    if (!m->lockstep || W == 1) continue;
    mg_use(m);
    const TranscriptArgs tr = {K[i].d_tr, K[i].d_root, 1, K[i].d_alpha_raw, K[i].d_alpha_m};
    const MgExchange X = exchange(m, K[i], MG_X_WAIT);
    rc = merkle_mg_top_dev(m->ctx, &tr, &X);
  }
  // fold of the rank's output range into every replica (k_fri_fold_bcast), then the barrier that completes the replicas
  for (int i = 0; i < n_here && rc == STARK_OK; i++) {
    stark_mgpu *m = ranks[i];
    stark_ctx *ctx = m->ctx;
    mg_use(m);
    FoldPeers P;
    memset(&P, 0, sizeof P);
    P.n = m->world, P.mc = nullptr;
    for (int g = 0; g < m->world; g++) P.out[g] = reinterpret_cast<u32 *>(m->peer[g] + m->L.arena);
    const size_t i0 = hper * (size_t)m->rank;
    size_t blocks = (hper / 4 + 255) / 256, cap = (size_t)ctx->sm_count * 8;
    LAUNCH(ctx, "fri_fold_bcast", 12ull * hper,
           k_fri_fold_bcast<<<(u32)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(
               codewords[i]->ptr, P, h, i0, i0 + hper, 0, K[i].G, ff::to_mont(g0), (const u32 *)K[i].d_alpha_m, 0u,
               ff::to_mont(ff::inv(ff::mul(2, offset)))));
    m->bytes_sent += 4ull * hper * (W - 1);
  }
  const u32 fin = MG_MAX_ROUNDS + 1;
  for (int i = 0; i < n_here && rc == STARK_OK; i++) {
    mg_use(ranks[i]);
    rc = ranks[i]->lockstep ? mg_barrier_signal(ranks[i], 1, mg_epoch(ranks[i], fin)) : mg_barrier(ranks[i], 1, mg_epoch(ranks[i], fin));
  }
  for (int i = 0; i < n_here && rc == STARK_OK; i++) {
    mg_use(ranks[i]);
    if (ranks[i]->lockstep) rc = mg_barrier_wait(ranks[i], 1, mg_epoch(ranks[i], fin));
  }
  for (int i = 0; i < n_here; i++) {
    stark_mgpu *m = ranks[i];
    stark_ctx *ctx = m->ctx;
    mg_use(m);
    Rk &k = K[i];
    if (rc == STARK_OK && roots && roots[i] && cudaMemcpyAsync(roots[i], k.d_root, 32, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
      rc = stark_fail(ctx, STARK_ERR_CUDA, "D2H copy failed");
    if (rc == STARK_OK && alpha_raw && cudaMemcpyAsync(alpha_raw + i, k.d_alpha_raw, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
      rc = stark_fail(ctx, STARK_ERR_CUDA, "D2H copy failed");
    const int e = mg_check_err(m);
    if (rc == STARK_OK) rc = e;
    if (k.sub) stark_merkle_free(k.sub);
    dev_free(ctx, k.top), dev_free(ctx, k.d_root), dev_free(ctx, k.d_alpha_raw), dev_free(ctx, k.d_alpha_m), dev_free(ctx, k.d_tr);
    if (rc == STARK_OK && folded) rc = stark_buf_wrap(ctx, m->win + m->L.arena, h, &folded[i]);
  }
  return rc;
}

// ---- BASELINE config 4 on a group: n_groups fixed groups of group_width trace columns; rank g owns
// groups {g, g + G, ...}.  Per owned group: coset LDE of its columns (no communication) and one Merkle tree whose leaf i
// is Hash::from_field_elements(row i of the group's LDE) (hash.rs:32-35), built on a side stream while the next group's
// LDE runs (the NTT loads the FMA-heavy pipe, the hashing the ALU pipe); the group roots are gathered -- ncclAllGather
// in a multi-process group, peer stores in a local one -- and every rank builds MerkleTree::new over them.
static int mg_lde_commit_impl(stark_mgpu *const *ranks, int n_here, const uint64_t *host_cols,
                              const stark_buf *const *owned_groups, uint32_t n_groups, uint32_t gw, uint32_t log_n,
                              uint32_t log_blowup, uint64_t offset, uint8_t *const *group_roots, uint8_t *const *commitments) {
  stark_ctx *ctx0 = ranks[0]->ctx;
  const u32 W = (u32)ranks[0]->world;
  if (n_groups == 0 || gw == 0 || n_groups % W) return stark_fail(ctx0, STARK_ERR_ARG, "the group count must be a multiple of the world size");
  if (n_groups & (n_groups - 1)) return stark_fail(ctx0, STARK_ERR_ARG, "Number of leaves must be power of 2");   // merkle.rs:13-16
  if (log_n + log_blowup > (u32)ff::TWO_ADICITY) return stark_fail(ctx0, STARK_ERR_ARG, "n > 2^23 not supported by this modulus");
  if (offset == 0 || offset >= ff::P) return stark_fail(ctx0, STARK_ERR_ARG, "offset must be a non-zero canonical element");
  if (n_groups > ranks[0]->L.max_cols / 2) return stark_fail(ctx0, STARK_ERR_ARG, "more groups than the group's window holds");
  const size_t n = (size_t)1 << log_n, N = n << log_blowup;
  const u32 own = n_groups / W;
  struct Rk {
    std::vector<u32 *> lde;
    std::vector<stark_tree *> trees;
    std::vector<stark_buf *> in;
    u8 *d_mine = nullptr, *d_all = nullptr, *d_ordered = nullptr;
    stark_tree *top = nullptr;
  };
  std::vector<Rk> K(n_here);
  int rc = STARK_OK;
  for (int i = 0; i < n_here && rc == STARK_OK; i++) {
    stark_mgpu *m = ranks[i];
    stark_ctx *ctx = m->ctx;
    mg_use(m);
    mg_begin_op(m);
    Rk &k = K[i];
    rc = dev_alloc(ctx, (void **)&k.d_mine, 32 * (size_t)own);
    if (rc == STARK_OK) rc = dev_alloc(ctx, (void **)&k.d_all, 32 * (size_t)n_groups);
    if (rc == STARK_OK) rc = dev_alloc(ctx, (void **)&k.d_ordered, 32 * (size_t)n_groups);
    if (rc == STARK_OK) rc = side_streams(ctx, 4);
    if (rc == STARK_OK && host_cols) rc = upload_flag_reset(ctx);
    cudaStream_t main_stream = ctx->stream, tree_stream = ctx->side[3];
    for (u32 j = 0; j < own && rc == STARK_OK; j++) {
      const u32 grp = (u32)m->rank + j * W;
      const u32 *cols_dev = nullptr;
      if (host_cols) {
        stark_buf *b = nullptr;
        rc = stark_buf_alloc(ctx, n * gw, &b);
        if (rc == STARK_OK) k.in.push_back(b), rc = upload_u64_nosync(ctx, host_cols + (size_t)grp * gw * n, n * gw, b->ptr);
        if (rc == STARK_OK) cols_dev = b->ptr;
      } else {
        const stark_buf *b = owned_groups[(size_t)i * own + j];
        if (!b || b->n < n * gw) rc = stark_fail(ctx, STARK_ERR_ARG, "buffer too small");
        else cols_dev = b->ptr;
      }
      u32 *lde = nullptr;
      if (rc == STARK_OK) rc = dev_alloc(ctx, (void **)&lde, N * gw * 4);
      if (rc == STARK_OK) k.lde.push_back(lde), rc = lde_dev(ctx, cols_dev, gw, log_n, log_blowup, (u32)offset, lde);
      if (rc != STARK_OK) break;
      // the group's tree on the tree stream, after its LDE
      cudaEventRecord(ctx->fork_ev, main_stream);
      cudaStreamWaitEvent(tree_stream, ctx->fork_ev, 0);
      ctx->stream = tree_stream, ctx->climb_counter = ctx->flag + TICKET_SIDE;
      stark_tree *t = nullptr;
      rc = merkle_build_from_dev_values(ctx, lde, N, gw, 1, N, &t);   // column-major [gw][N]: row stride 1, column stride N
      if (rc == STARK_OK) {
        k.trees.push_back(t);
        if (cudaMemcpyAsync(k.d_mine + 32 * (size_t)j, t->nodes + 32 * (2 * N - 2), 32, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)
          rc = stark_fail(ctx, STARK_ERR_CUDA, "D2D copy failed");
      }
      ctx->stream = main_stream, ctx->climb_counter = ctx->flag + TICKET_MAIN;
    }
    cudaEventRecord(ctx->side_done[3], tree_stream);
    cudaStreamWaitEvent(main_stream, ctx->side_done[3], 0);
    if (rc != STARK_OK) break;
    // gather the group roots
    if (m->mode == MG_PROC && W > 1) {
      rc = mg_all_gather(m, k.d_mine, k.d_all, 32 * (size_t)own);   // rank-major: entry (g, j) = group g + j W
    } else {
      std::vector<u32> idx(own);
      for (u32 j = 0; j < own; j++) idx[j] = (u32)m->rank + j * W;
      rc = mg_put_roots(m, k.d_mine, idx.data(), own);
    }
  }
  const u32 fin = MG_MAX_ROUNDS + 1;
  const bool peer_path = ranks[0]->mode != MG_PROC || W == 1;
  if (peer_path) {
    for (int i = 0; i < n_here && rc == STARK_OK; i++) {
      mg_use(ranks[i]);
      rc = ranks[i]->lockstep ? mg_barrier_signal(ranks[i], 1, mg_epoch(ranks[i], fin)) : mg_barrier(ranks[i], 1, mg_epoch(ranks[i], fin));
    }
    for (int i = 0; i < n_here && rc == STARK_OK; i++) {
      mg_use(ranks[i]);
      if (ranks[i]->lockstep) rc = mg_barrier_wait(ranks[i], 1, mg_epoch(ranks[i], fin));
    }
  }
  for (int i = 0; i < n_here && rc == STARK_OK; i++) {
    stark_mgpu *m = ranks[i];
    stark_ctx *ctx = m->ctx;
    mg_use(m);
    Rk &k = K[i];
    if (peer_path) {
      if (cudaMemcpyAsync(k.d_ordered, mg_colroots(m, m->rank), 32 * (size_t)n_groups, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)
        rc = stark_fail(ctx, STARK_ERR_CUDA, "D2D copy failed");
    } else {
      // rank-major (g, j) -> group order g + j W
      for (u32 g = 0; g < W && rc == STARK_OK; g++)
        if (cudaMemcpy2DAsync(k.d_ordered + 32 * (size_t)g, 32 * (size_t)W, k.d_all + 32 * (size_t)g * own, 32, 32, own,
                              cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)
          rc = stark_fail(ctx, STARK_ERR_CUDA, "D2D copy failed");
    }
    if (rc == STARK_OK) rc = stark_merkle_build_dev(ctx, k.d_ordered, n_groups, &k.top);   // MerkleTree::new over the group roots
    if (rc == STARK_OK && group_roots && group_roots[i] &&
        cudaMemcpyAsync(group_roots[i], k.d_ordered, 32 * (size_t)n_groups, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
      rc = stark_fail(ctx, STARK_ERR_CUDA, "D2H copy failed");
    if (rc == STARK_OK && commitments && commitments[i] &&
        cudaMemcpyAsync(commitments[i], k.top->nodes + 32 * (2 * (size_t)n_groups - 2), 32, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
      rc = stark_fail(ctx, STARK_ERR_CUDA, "D2H copy failed");
  }
  for (int i = 0; i < n_here; i++) {
    stark_mgpu *m = ranks[i];
    stark_ctx *ctx = m->ctx;
    mg_use(m);
    Rk &k = K[i];
    const int e = mg_check_err(m);
    if (rc == STARK_OK) rc = e;
    if (rc == STARK_OK && host_cols) rc = upload_u64_check(ctx);
    for (stark_tree *t : k.trees) stark_merkle_free(t);
    if (k.top) stark_merkle_free(k.top);
    for (u32 *p : k.lde) dev_free(ctx, p);
    for (stark_buf *b : k.in) stark_buf_free(b);
    dev_free(ctx, k.d_mine), dev_free(ctx, k.d_all), dev_free(ctx, k.d_ordered);
  }
  return rc;
}

// ----------------------------------------------------------------------------------------------- C ABI

extern "C" {

int stark_fri_num_rounds(size_t domain_length, uint32_t ef, uint32_t nq, uint32_t *rounds) {
  if (!rounds) return stark_fail(nullptr, STARK_ERR_ARG, "null argument");
  return fri_check(nullptr, domain_length, ef, rounds, nq);
}

int stark_fri_fold_dev(stark_ctx *ctx, const stark_buf *codeword, size_t n, uint64_t alpha_raw, uint64_t offset,
                       uint64_t omega, stark_buf *out) {
  if (!ctx || !codeword || !out) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  const size_t h = n / 2;
  if (codeword->n < n || out->n < h) return stark_fail(ctx, STARK_ERR_ARG, "buffer too small");
  if (h == 0) return STARK_OK;
  u32 off, om;
  reduce_params(ctx, offset, omega, &off, &om);
  if (off == 0 || (om == 0 && h > 1)) return stark_fail(ctx, STARK_ERR_ARG, "no division by zero");  // ff.rs:182
  const u32 g0 = om ? ff::inv(om) : 1u;
  GeoTables G;
  ST_TRY(geo_tables(ctx, g0, 1, h, &G));
  const u32 am = ff::to_mont(ff::reduce64(alpha_raw));   // alpha travels as a kernel argument: no allocation, no copy
  return fold_launch_range(ctx, codeword->ptr, out->ptr, h, 0, h, 0, G, ff::to_mont(g0), nullptr,
                           ff::to_mont(ff::inv(ff::mul(2, off))), am);
}

// one rank's share of fold_codeword (fri.rs:57-91): outputs [i0, i0 + count) written to out[out_off ..]
int stark_fri_fold_range_dev(stark_ctx *ctx, const stark_buf *codeword, size_t n, uint64_t alpha_raw, uint64_t offset,
                             uint64_t omega, size_t i0, size_t count, stark_buf *out, size_t out_off) {
  if (!ctx || !codeword || !out) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  const size_t h = n / 2;
  if (codeword->n < n || i0 + count > h || out->n < out_off + count) return stark_fail(ctx, STARK_ERR_ARG, "range out of bounds");
  if (count == 0) return STARK_OK;
  u32 off, om;
  reduce_params(ctx, offset, omega, &off, &om);
  if (off == 0 || (om == 0 && h > 1)) return stark_fail(ctx, STARK_ERR_ARG, "no division by zero");  // ff.rs:182
  const u32 g0 = om ? ff::inv(om) : 1u;
  GeoTables G;
  ST_TRY(geo_tables(ctx, g0, 1, h, &G));
  const u32 am = ff::to_mont(ff::reduce64(alpha_raw));
  // the kernel indexes out by the global output index
  u32 *base = out->ptr + out_off;
  return fold_launch_range(ctx, codeword->ptr, base - i0, h, i0, i0 + count, 0, G, ff::to_mont(g0), nullptr,
                           ff::to_mont(ff::inv(ff::mul(2, off))), am);
}

// fold_codeword range fused with the replication of the result (k_fri_fold_bcast).  peers[g] = rank g's replica of the
// next codeword (device addresses valid on THIS device: CUDA IPC / symmetric memory), multicast = NVSwitch multicast
// address of the same buffer or NULL; all indexed from output 0.  i0, count multiples of 4.
int stark_fri_fold_bcast_dev(stark_ctx *ctx, const stark_buf *codeword, size_t n, uint64_t alpha_raw, uint64_t offset,
                             uint64_t omega, size_t i0, size_t count, void *const *peers, int n_peers, void *multicast) {
  if (!ctx || !codeword || (!peers && !multicast)) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  const size_t h = n / 2;
  if (codeword->n < n || i0 + count > h) return stark_fail(ctx, STARK_ERR_ARG, "range out of bounds");
  if (n_peers < 0 || n_peers > 8) return stark_fail(ctx, STARK_ERR_ARG, "at most 8 peers");
  if ((h % 4) || (i0 % 4) || (count % 4)) return stark_fail(ctx, STARK_ERR_ARG, "fold range must be a multiple of 4");
  if (((uintptr_t)codeword->ptr | (uintptr_t)multicast) & 15u) return stark_fail(ctx, STARK_ERR_ARG, "device pointer must be 16-byte aligned");
  for (int g = 0; g < n_peers; g++)
    if ((uintptr_t)peers[g] & 15u) return stark_fail(ctx, STARK_ERR_ARG, "device pointer must be 16-byte aligned");
  if (count == 0) return STARK_OK;
  u32 off, om;
  reduce_params(ctx, offset, omega, &off, &om);
  if (off == 0 || (om == 0 && h > 1)) return stark_fail(ctx, STARK_ERR_ARG, "no division by zero");  // ff.rs:182
  const u32 g0 = om ? ff::inv(om) : 1u;
  GeoTables G;
  ST_TRY(geo_tables(ctx, g0, 1, h, &G));
  const u32 am = ff::to_mont(ff::reduce64(alpha_raw));
  FoldPeers P;
  memset(&P, 0, sizeof P);
  P.n = n_peers, P.mc = (u32 *)multicast;
  for (int g = 0; g < n_peers; g++) P.out[g] = (u32 *)peers[g];
  size_t blocks = (count / 4 + 255) / 256, cap = (size_t)ctx->sm_count * 8;
  LAUNCH(ctx, "fri_fold_bcast", 12ull * count,
         k_fri_fold_bcast<<<(u32)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(
             codeword->ptr, P, h, i0, i0 + count, 0, G, ff::to_mont(g0), nullptr, am, ff::to_mont(ff::inv(ff::mul(2, off)))));
  return STARK_OK;
}

// FiatShamir::challenge (fiat_shamir.rs:19-25) for a host-held transcript: first 8 bytes, little-endian, of
// Hash::from_bytes(transcript), UNREDUCED.  Host logic of the sharded prover (32 R bytes per proof).
int stark_fiat_shamir_challenge(const uint8_t *transcript, size_t len, uint64_t *challenge_raw) {
  if ((!transcript && len) || !challenge_raw) return stark_fail(nullptr, STARK_ERR_ARG, "null argument");
  u8 h[32];
  hs::from_bytes(transcript, len, h);
  u64 v = 0;
  for (int b = 0; b < 8; b++) v |= (u64)h[b] << (8 * b);
  *challenge_raw = v;
  return STARK_OK;
}
// Hash::from_u64 (hash.rs:37-39) on the host: the index seed of fri.rs:272
int stark_hash_from_u64(uint64_t value, uint8_t out[32]) {
  if (!out) return stark_fail(nullptr, STARK_ERR_ARG, "null argument");
  u8 m[8];
  for (int b = 0; b < 8; b++) m[b] = (u8)(value >> (8 * b));
  hs::from_bytes(m, 8, out);
  return STARK_OK;
}

int stark_fri_fold(stark_ctx *ctx, const uint64_t *codeword, size_t n, uint64_t alpha_raw, uint64_t offset,
                   uint64_t omega, uint64_t *out) {
  if (!ctx || (n && !codeword) || (n / 2 && !out)) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  stark_buf *in = nullptr, *o = nullptr;
  ST_TRY(stark_buf_upload(ctx, codeword, n, &in));
  int rc = stark_buf_alloc(ctx, n / 2, &o);
  if (rc == STARK_OK) rc = stark_fri_fold_dev(ctx, in, n, alpha_raw, offset, omega, o);
  if (rc == STARK_OK) rc = download_u64(ctx, o->ptr, n / 2, out);
  stark_buf_free(in), stark_buf_free(o);
  return rc;
}

int stark_fri_commit_dev(stark_ctx *ctx, const stark_buf *codeword, size_t n, uint64_t offset, uint64_t omega,
                         uint32_t ef, uint32_t nq, const uint8_t *transcript, size_t transcript_len,
                         stark_fri_state **out) {
  if (!ctx || !codeword || !out) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (codeword->n < n) return stark_fail(ctx, STARK_ERR_ARG, "buffer too small");
  u32 off, om;
  reduce_params(ctx, offset, omega, &off, &om);
  return fri_commit_dev(ctx, codeword->ptr, n, off, om, ef, nq, transcript, transcript_len, true, out);
}
int stark_fri_commit(stark_ctx *ctx, const uint64_t *codeword, size_t n, uint64_t offset, uint64_t omega, uint32_t ef,
                     uint32_t nq, const uint8_t *transcript, size_t transcript_len, stark_fri_state **out) {
  if (!ctx || !codeword || !out) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  u32 R;
  ST_TRY(fri_check(ctx, n, ef, &R, nq));
  stark_buf *in = nullptr;
  ST_TRY(stark_buf_upload(ctx, codeword, n, &in));
  int rc = stark_fri_commit_dev(ctx, in, n, offset, omega, ef, nq, transcript, transcript_len, out);
  stark_buf_free(in);
  return rc;
}
uint32_t stark_fri_rounds(const stark_fri_state *s) { return s ? s->rounds : 0; }
int stark_fri_roots(stark_fri_state *s, uint8_t *out) {
  if (!s || !out) return stark_fail(nullptr, STARK_ERR_ARG, "null argument");
  if (!s->rounds) return STARK_OK;
  CU_TRY(s->ctx, cudaMemcpyAsync(out, s->d_roots, 32 * (size_t)s->rounds, cudaMemcpyDeviceToHost, s->ctx->stream));
  CU_TRY(s->ctx, cudaStreamSynchronize(s->ctx->stream));
  return STARK_OK;
}
int stark_fri_alphas(stark_fri_state *s, uint64_t *out) {
  if (!s || !out) return stark_fail(nullptr, STARK_ERR_ARG, "null argument");
  if (s->rounds < 2) return STARK_OK;
  CU_TRY(s->ctx, cudaMemcpyAsync(out, s->d_alpha_raw, 8 * (size_t)(s->rounds - 1), cudaMemcpyDeviceToHost, s->ctx->stream));
  CU_TRY(s->ctx, cudaStreamSynchronize(s->ctx->stream));
  return STARK_OK;
}
int stark_fri_codeword_len(const stark_fri_state *s, uint32_t round, size_t *len) {
  if (!s || !len || round >= s->cw.size()) return stark_fail(nullptr, STARK_ERR_ARG, "round out of range");
  *len = s->len[round];
  return STARK_OK;
}
int stark_fri_codeword(stark_fri_state *s, uint32_t round, uint64_t *out) {
  if (!s || !out || round >= s->cw.size()) return stark_fail(nullptr, STARK_ERR_ARG, "round out of range");
  return download_u64(s->ctx, s->cw[round], s->len[round], out);
}
int stark_fri_open(stark_fri_state *s, uint32_t round, size_t index, uint8_t *out, size_t *n_hashes) {
  if (!s || round >= s->trees.size()) return stark_fail(nullptr, STARK_ERR_ARG, "round out of range");
  return stark_merkle_open(s->trees[round], index, out, n_hashes);
}
void stark_fri_free(stark_fri_state *s) { fri_state_free(s); }

int stark_fri_sample_indices(const uint8_t *seed, size_t seed_len, size_t size, size_t reduced_size, size_t number,
                             uint64_t *out) {
  // host-side mirror of fri.rs:176-213 for callers that hold their own seed (tiny: `number` 36-byte hashes)
  if ((!seed && seed_len) || (!out && number)) return stark_fail(nullptr, STARK_ERR_ARG, "null argument");
  if (number > 2 * reduced_size) return stark_fail(nullptr, STARK_ERR_ARG, "not enough entropy in indices wrt last codeword");
  if (number > reduced_size)
    return stark_fail(nullptr, STARK_ERR_ARG, "cannot sample more indices than available in last codeword; requested: %zu, available: %zu", number, reduced_size);
  std::vector<u8> msg(seed_len + 4);
  for (size_t i = 0; i < seed_len; i++) msg[i] = seed[i];
  size_t got = 0;
  for (u32 counter = 0; got < number; counter++) {
    for (int b = 0; b < 4; b++) msg[seed_len + b] = (u8)(counter >> (8 * b));
    u8 h[32];
    hs::from_bytes(msg.data(), msg.size(), h);
    u64 v = 0;
    for (int b = 24; b < 32; b++) v = (v << 8) | h[b];
    const u64 idx = v % size, ri = idx % reduced_size;
    bool seen = false;
    for (size_t j = 0; j < got; j++)
      if (out[j] % reduced_size == ri) seen = true;
    if (!seen) out[got++] = idx;
  }
  return STARK_OK;
}

int stark_fri_proof_size(size_t domain_length, uint32_t ef, uint32_t nq, size_t *bytes) {
  if (!bytes) return stark_fail(nullptr, STARK_ERR_ARG, "null argument");
  u32 R = 0;
  ST_TRY(fri_check(nullptr, domain_length, ef, &R, nq));
  ProofLayout L;
  proof_layout(domain_length, R, nq, &L);
  *bytes = L.total;
  return STARK_OK;
}

int stark_fri_prove_dev(stark_ctx *ctx, const stark_buf *codeword, size_t n, size_t domain_length, uint64_t offset,
                        uint64_t omega, uint32_t ef, uint32_t nq, const uint8_t *transcript, size_t transcript_len,
                        uint8_t *proof, size_t proof_cap, size_t *proof_len, uint64_t *top_indices) {
  if (!ctx || !codeword) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (codeword->n < n) return stark_fail(ctx, STARK_ERR_ARG, "buffer too small");
  u32 off, om;
  reduce_params(ctx, offset, omega, &off, &om);
  return fri_prove_dev(ctx, codeword->ptr, n, domain_length, off, om, ef, nq, transcript, transcript_len, proof,
                       proof_cap, proof_len, top_indices);
}
int stark_fri_prove(stark_ctx *ctx, const uint64_t *codeword, size_t n, size_t domain_length, uint64_t offset,
                    uint64_t omega, uint32_t ef, uint32_t nq, const uint8_t *transcript, size_t transcript_len,
                    uint8_t *proof, size_t proof_cap, size_t *proof_len, uint64_t *top_indices) {
  if (!ctx || (n && !codeword)) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (n != domain_length)
    return stark_fail(ctx, STARK_ERR_ARG, "initial codeword length does not match domain length");
  stark_buf *in = nullptr;
  ST_TRY(stark_buf_upload(ctx, codeword, n, &in));
  int rc = stark_fri_prove_dev(ctx, in, n, domain_length, offset, omega, ef, nq, transcript, transcript_len, proof,
                               proof_cap, proof_len, top_indices);
  stark_buf_free(in);
  return rc;
}

// BASELINE config 3: LDE of every column, one Merkle tree per column, FRI on column 0
int stark_prove_trace_dev(stark_ctx *ctx, const stark_buf *cols, uint32_t n_cols, uint32_t log_n, uint32_t log_blowup,
                          uint64_t offset, uint32_t nq, uint8_t *column_roots, uint8_t *proof, size_t proof_cap,
                          size_t *proof_len) {
  if (!ctx || !cols || n_cols == 0) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (log_n + log_blowup > (u32)ff::TWO_ADICITY)
    return stark_fail(ctx, STARK_ERR_ARG, "n > 2^23 not supported by this modulus");
  if (offset == 0 || offset >= ff::P) return stark_fail(ctx, STARK_ERR_ARG, "offset must be a non-zero canonical element");
  const size_t n = (size_t)1 << log_n, N = n << log_blowup;
  if (cols->n < n * n_cols) return stark_fail(ctx, STARK_ERR_ARG, "buffer too small");
  return prove_trace_pipeline(ctx, nullptr, cols->ptr, n_cols, log_n, log_blowup, (u32)offset, nq, column_roots, proof, proof_cap,
                              proof_len);
}

int stark_prove_trace(stark_ctx *ctx, const uint64_t *cols, uint32_t n_cols, uint32_t log_n, uint32_t log_blowup,
                      uint64_t offset, uint32_t nq, uint8_t *column_roots, uint8_t *proof, size_t proof_cap,
                      size_t *proof_len) {
  if (!ctx || !cols || n_cols == 0) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (log_n + log_blowup > (u32)ff::TWO_ADICITY) return stark_fail(ctx, STARK_ERR_ARG, "n > 2^23 not supported by this modulus");
  if (offset == 0 || offset >= ff::P) return stark_fail(ctx, STARK_ERR_ARG, "offset must be a non-zero canonical element");
  stark_buf *in = nullptr;
  ST_TRY(stark_buf_alloc(ctx, ((size_t)1 << log_n) * n_cols, &in));
  // H2D (group by group on the copy stream) + narrowing without a host round trip: the canonical check is read after
  // the pipeline's final synchronisation
  const int rc = prove_trace_pipeline(ctx, cols, in->ptr, n_cols, log_n, log_blowup, (u32)offset, nq, column_roots, proof, proof_cap,
                                      proof_len);
  stark_buf_free(in);
  return rc;
}

// the same from the reference's own trace layout (trace.rs:4-34): row-major i128 values, 16 bytes each
int stark_prove_trace_rows(stark_ctx *ctx, const void *rows_i128, uint32_t n_cols, uint32_t log_n, uint32_t log_blowup,
                           uint64_t offset, uint32_t nq, uint8_t *column_roots, uint8_t *proof, size_t proof_cap,
                           size_t *proof_len) {
  if (!ctx || !rows_i128 || n_cols == 0) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (log_n > (u32)ff::TWO_ADICITY) return stark_fail(ctx, STARK_ERR_ARG, "n > 2^23 not supported by this modulus");
  const size_t n = (size_t)1 << log_n;
  stark_buf *in = nullptr;
  ST_TRY(stark_buf_alloc(ctx, n * n_cols, &in));
  int rc = trace_to_columns_dev(ctx, rows_i128, n, n_cols, in->ptr);
  if (rc == STARK_OK)
    rc = stark_prove_trace_dev(ctx, in, n_cols, log_n, log_blowup, offset, nq, column_roots, proof, proof_cap, proof_len);
  stark_buf_free(in);
  return rc;
}

}  // extern "C"

extern "C" {

uint32_t stark_mgpu_owned_columns(const stark_mgpu *m, uint32_t n_cols, uint32_t *out) {
  return m ? mg_owned_columns(m->rank, m->world, n_cols, out) : 0;
}
// the same partition without a group handle (pure host logic: usable, and tested, without a device)
uint32_t stark_mgpu_columns_of_rank(int rank, int world, uint32_t n_cols, uint32_t *out) {
  if (world < 1 || rank < 0 || rank >= world) return 0;
  return mg_owned_columns(rank, world, n_cols, out);
}

static int mg_check_ranks(stark_mgpu *const *ranks, int n_here) {
  if (!ranks || n_here < 1 || n_here > MG_MAX_RANKS) return stark_fail(nullptr, STARK_ERR_ARG, "null argument");
  for (int k = 0; k < n_here; k++)
    if (!ranks[k]) return stark_fail(nullptr, STARK_ERR_ARG, "null argument");
  if (ranks[0]->mode == MG_LOCAL) {
    if (n_here != ranks[0]->world) return stark_fail(ranks[0]->ctx, STARK_ERR_ARG, "a local group is driven with all of its ranks in one call");
    for (int k = 0; k < n_here; k++)
      if (ranks[k]->rank != k || ranks[k]->group != ranks[0]->group) return stark_fail(ranks[0]->ctx, STARK_ERR_ARG, "ranks of a local group must be passed in rank order");
  } else if (n_here != 1) {
    return stark_fail(ranks[0]->ctx, STARK_ERR_ARG, "a multi-process group is driven one rank per call");
  }
  return STARK_OK;
}

int stark_mgpu_prove_trace(stark_mgpu *const *ranks, int n_here, const uint64_t *cols, uint32_t n_cols, uint32_t log_n,
                           uint32_t log_blowup, uint64_t offset, uint32_t nq, uint8_t *const *column_roots,
                           uint8_t *const *proofs, size_t proof_cap, size_t *proof_len) {
  MgDeviceScope restore_device;
  ST_TRY(mg_check_ranks(ranks, n_here));
  if (!cols) return stark_fail(ranks[0]->ctx, STARK_ERR_ARG, "null argument");
  return mg_prove_trace_impl(ranks, n_here, cols, nullptr, n_cols, log_n, log_blowup, offset, nq, column_roots, proofs, proof_cap, proof_len);
}
int stark_mgpu_prove_trace_dev(stark_mgpu *const *ranks, int n_here, const stark_buf *const *my_cols, uint32_t n_cols,
                               uint32_t log_n, uint32_t log_blowup, uint64_t offset, uint32_t nq,
                               uint8_t *const *column_roots, uint8_t *const *proofs, size_t proof_cap, size_t *proof_len) {
  MgDeviceScope restore_device;
  ST_TRY(mg_check_ranks(ranks, n_here));
  if (!my_cols) return stark_fail(ranks[0]->ctx, STARK_ERR_ARG, "null argument");
  return mg_prove_trace_impl(ranks, n_here, nullptr, my_cols, n_cols, log_n, log_blowup, offset, nq, column_roots, proofs, proof_cap, proof_len);
}

int stark_mgpu_fri_prove_dev(stark_mgpu *const *ranks, int n_here, const stark_buf *const *codewords, size_t n,
                             size_t domain_length, uint64_t offset, uint64_t omega, uint32_t ef, uint32_t nq,
                             const uint8_t *transcript, size_t transcript_len, uint8_t *const *proofs, size_t proof_cap,
                             size_t *proof_len, uint64_t *const *top_indices) {
  MgDeviceScope restore_device;
  ST_TRY(mg_check_ranks(ranks, n_here));
  if (!codewords || !proofs) return stark_fail(ranks[0]->ctx, STARK_ERR_ARG, "null argument");
  u32 off, om;
  reduce_params(ranks[0]->ctx, offset, omega, &off, &om);
  std::vector<MgProve> P(n_here);
  int rc = STARK_OK;
  for (int k = 0; k < n_here; k++) {
    mg_use(ranks[k]);
    if (!codewords[k] || !proofs[k] || codewords[k]->n < n) {
      mg_begin_op(ranks[k]);
      rc = stark_fail(ranks[k]->ctx, STARK_ERR_ARG, "buffer too small");
      continue;
    }
    const int e = P[k].begin(ranks[k], codewords[k]->ptr, n, domain_length, off, om, ef, nq, transcript, transcript_len, proof_cap, proof_len);
    if (rc == STARK_OK) rc = e;
  }
  if (rc == STARK_OK) rc = mg_prove_run(P.data(), n_here);
  if (rc == STARK_OK) rc = mg_prove_finish(P.data(), n_here);
  for (int k = 0; k < n_here; k++) {
    mg_use(ranks[k]);
    if (rc == STARK_OK) rc = P[k].download(proofs[k], top_indices ? top_indices[k] : nullptr);
  }
  for (int k = 0; k < n_here; k++) {
    mg_use(ranks[k]);
    const int e = mg_check_err(ranks[k]);
    if (rc == STARK_OK) rc = e;
    P[k].release();
  }
  return rc;
}

int stark_mgpu_fold_commit_round(stark_mgpu *const *ranks, int n_here, const stark_buf *const *codewords, size_t n,
                                 uint64_t offset, uint64_t omega, uint8_t *const *roots, uint64_t *alpha_raw,
                                 stark_buf **folded) {
  MgDeviceScope restore_device;
  ST_TRY(mg_check_ranks(ranks, n_here));
  if (!codewords) return stark_fail(ranks[0]->ctx, STARK_ERR_ARG, "null argument");
  u32 off, om;
  reduce_params(ranks[0]->ctx, offset, omega, &off, &om);
  return mg_fold_commit_round_impl(ranks, n_here, codewords, n, off, om, roots, alpha_raw, folded);
}

int stark_mgpu_lde_commit(stark_mgpu *const *ranks, int n_here, const uint64_t *cols, uint32_t n_groups, uint32_t group_width,
                          uint32_t log_n, uint32_t log_blowup, uint64_t offset, uint8_t *const *group_roots,
                          uint8_t *const *commitments) {
  MgDeviceScope restore_device;
  ST_TRY(mg_check_ranks(ranks, n_here));
  if (!cols) return stark_fail(ranks[0]->ctx, STARK_ERR_ARG, "null argument");
  return mg_lde_commit_impl(ranks, n_here, cols, nullptr, n_groups, group_width, log_n, log_blowup, offset, group_roots, commitments);
}
int stark_mgpu_lde_commit_dev(stark_mgpu *const *ranks, int n_here, const stark_buf *const *owned_groups, uint32_t n_groups,
                              uint32_t group_width, uint32_t log_n, uint32_t log_blowup, uint64_t offset,
                              uint8_t *const *group_roots, uint8_t *const *commitments) {
  MgDeviceScope restore_device;
  ST_TRY(mg_check_ranks(ranks, n_here));
  if (!owned_groups) return stark_fail(ranks[0]->ctx, STARK_ERR_ARG, "null argument");
  return mg_lde_commit_impl(ranks, n_here, nullptr, owned_groups, n_groups, group_width, log_n, log_blowup, offset, group_roots, commitments);
}

}  // extern "C"
