// verify.cu -- Fri::verify (reference src/fri.rs:313-505) + test_colinearity (fri.rs:507-525) +
// MerkleTree::verify (merkle.rs:82-97) as a batch verifier on the device.
//
// The reference pops the deserialised proof stream object by object and returns false at the first failed check.
// Here the host only walks the stream STRUCTURE (ProofStream::deserialize, stream.rs:66-168: tags and counts, no
// payload inspection) and the device does every payload check in parallel:
//   k_verify_transcript   the R roots are absorbed one after the other and a challenge is drawn after each
//                         (fri.rs:324-334) -- one warp, the 4-lanes-per-hash form of the sponge (transcript.cuh)
//   k_verify_last_leaves  + merkle_climb_dev: Merkle root of the last codeword (fri.rs:349-357)
//   iNTT on the last coset + k_verify_last: root comparison and degree bound (fri.rs:359-399)
//   k_sample_indices      (fri.cu) the query indices from the transcript (fri.rs:401-407)
//   k_verify_rounds       one thread per (round, query, {path a, path b, path c, colinearity}): leaf hash, climb along
//                         the authentication path, comparison with the round's root (fri.rs:410-500)
// Every check has a KEY that orders it the way the reference reaches it (stream object index, then the check's rank
// at that object); kernels atomicMin their first failing key, the host merges it with the structural failures, and
// the smallest key names the reference's `println!` reason.  All checks passing <=> Fri::verify returns true.
#include <vector>

#include "common.cuh"
#include "hash.cuh"
#include "merkle.h"
#include "merkle_dev.cuh"
#include "transcript.cuh"

using hs::State;

int fri_sample_indices_dev(stark_ctx *ctx, const TranscriptDev *T, u64 *d_challenge, u64 size, u64 reduced, u32 number,
                           u64 *d_out);   // fri.cu

namespace {

struct VObj {
  u32 tag, count;   // stream.rs:40-60 tag; number of COMPLETE items present (1 for a root / single element)
  u64 off;          // byte offset of the payload (after the tag and the u64 count)
};

// reasons = the reference's println! texts, see stark_fri_verify_reason
enum {
  V_OK = 0, V_ROOT_EXTRACT = 1, V_LAST_EXTRACT = 2, V_NO_ROOTS = 3, V_LAST_ROOT = 4, V_LAST_SMALL = 5, V_REEVAL = 6,
  V_DEGREE = 7, V_TRIPLE_EXTRACT = 8, V_TRIPLE_LEN = 9, V_COLINEAR = 10, V_PATH_A = 11, V_PATH_B = 12, V_PATH_C = 13,
  V_PATH_A_EXTRACT = 14, V_PATH_B_EXTRACT = 15, V_PATH_C_EXTRACT = 16,
  V_PANIC_SAMPLE_ENTROPY = 17, V_PANIC_SAMPLE_COUNT = 18, V_PANIC_SUB = 19
};
constexpr u64 NO_FAIL = ~0ull;
__host__ __device__ inline u64 vkey(u64 obj, u32 rank, u32 reason) { return ((obj * 8 + rank) << 8) | reason; }

__device__ __forceinline__ u64 load_u64(const u8 *p) {
  u64 v = 0;
#pragma unroll
  for (int b = 0; b < 8; b++) v |= (u64)p[b] << (8 * b);
  return v;
}
__device__ __forceinline__ void load_hash_bytes(const u8 *p, u32 *w) {   // unaligned 32 bytes -> 8 LE words
#pragma unroll
  for (int g = 0; g < 8; g++) w[g] = (u32)p[4 * g] | ((u32)p[4 * g + 1] << 8) | ((u32)p[4 * g + 2] << 16) | ((u32)p[4 * g + 3] << 24);
}
// Hash::from_field_elements(&[v]) for a RAW u64 (hash.rs:32-35): the stream carries FieldElement.value unreduced
__device__ __forceinline__ void leaf_u64(u64 v, u32 *out) {
  State st;
  hs::init(st);
#pragma unroll
  for (int i = 0; i < 8; i++) hs::absorb_byte(st, i, (u32)(v >> (8 * i)) & 0xffu);
  hs::mix_lazy<false>(st);
  hs::finalize<true>(st);
  hs::pack_words(st, out);
}

// fri.rs:324-334: roots[r] = pop(); absorb; alphas[r] = challenge().  One warp.
__global__ void __launch_bounds__(32) k_verify_transcript(const u8 *proof, const VObj *objs, u32 R, TranscriptDev *T,
                                                          u8 *roots_in, u8 *roots, u64 *alpha_raw, u32 *alpha_m) {
  for (u32 i = threadIdx.x; i < 32 * R; i += 32) roots_in[i] = proof[objs[i >> 5].off + (i & 31u)];
  __syncwarp();
  for (u32 r = 0; r < R; r++) {
    TranscriptArgs A = {T, roots + 32 * r, 1, alpha_raw + r, alpha_m + r};
    transcript_round_warp(A, roots_in + 32 * r);
    __syncwarp();
  }
}

// fri.rs:349-353 leaves of the last codeword; also its residues for the interpolation (FiniteField ops reduce, ff.rs:138)
__global__ void k_verify_last_leaves(const u8 *payload, u64 n, u8 *nodes, u32 *vals) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 v = load_u64(payload + 8 * i);
  u32 h[8];
  leaf_u64(v, h);
  store_hash(nodes + 32 * i, h);
  vals[i] = (u32)(v % ff::P);
}

// fri.rs:354-357 (root of the last codeword's tree against the last root) and 393-399 (degree of the interpolant)
__global__ void k_verify_last(const u8 *tree_root, const u8 *roots, u32 R, const u32 *coeffs, u64 n, u64 degree,
                              int check_degree, u64 *fail) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    bool same = true;
    for (int b = 0; b < 32; b++) same &= tree_root[b] == roots[32 * (R - 1) + b];
    if (!same) atomicMin((unsigned long long *)fail, (unsigned long long)vkey(R, 2, V_LAST_ROOT));
  }
  if (check_degree && i < n && i > degree && coeffs[i] != 0)
    atomicMin((unsigned long long *)fail, (unsigned long long)vkey(R, 4, V_DEGREE));
}

struct RoundsArgs {
  const u8 *proof;
  const VObj *objs;
  u64 n_obj;
  const u8 *roots;        // R x 32, aligned
  const u64 *alpha_raw;   // R raw challenges
  const u64 *top;         // nq top-level indices
  u64 domain_length;
  u32 R, nq;
  u32 offset_m, omega_m;  // Montgomery form of offset, omega (round 0)
  u64 *fail;
  u64 *poly_idx, *poly_val;   // 2 nq each (fri.rs:437-441) or null
};

// one thread per (round r = blockIdx.y, query s, job): job 0/1/2 = authentication path of a / b / c, job 3 = colinearity
__global__ void __launch_bounds__(64) k_verify_rounds(const __grid_constant__ RoundsArgs A) {
  pdl_entry();
  const u32 r = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 4 * A.nq) return;
  const u32 s = t >> 2, job = t & 3u;
  const u64 base = (u64)A.R + 1 + (u64)r * 4 * A.nq;      // first stream object of round r
  const u64 k_triple = base + s;
  if (k_triple >= A.n_obj) return;                         // structural failure, reported by the host
  const VObj tr = A.objs[k_triple];
  if (tr.tag != 2 || tr.count != 3) return;
  const u64 half = A.domain_length >> (r + 1);
  const u64 ci = A.top[s] % half, ai = ci, bi = ci + half;                     // fri.rs:412-424
  const u8 *tp = A.proof + tr.off;
  const u64 ay = load_u64(tp), by = load_u64(tp + 8), cy = load_u64(tp + 16);
  unsigned long long *fail = (unsigned long long *)A.fail;
  if (job == 3) {
    if (r == 0 && A.poly_idx) {
      A.poly_idx[2 * s] = ai, A.poly_val[2 * s] = ay, A.poly_idx[2 * s + 1] = bi, A.poly_val[2 * s + 1] = by;
    }
    // x_a = offset_r * omega_r^a, offset_r = offset^(2^r), omega_r = omega^(2^r)  (fri.rs:444-451, 497-498)
    const u32 xa_m = ff::canon(ff::mont_mul(ff::mont_pow(A.offset_m, 1ull << r), ff::mont_pow(A.omega_m, ai << r)));
    const u32 xb_m = ff::canon(ff::mont_mul(ff::mont_pow(A.offset_m, 1ull << r), ff::mont_pow(A.omega_m, bi << r)));
    const u32 x0 = ff::from_mont(xa_m), x1 = ff::from_mont(xb_m);
    const u64 x2 = A.alpha_raw[r];                                               // raw, unreduced (fiat_shamir.rs:21-24)
    // FiniteField::sub is ((p + l - r) as u128) % p (ff.rs:154-160): it underflows (debug panic) iff r > p + l.  Only the
    // differences against the raw y_a can: the x are canonical and p + alpha >= x_a.
    if ((ay > by && ay - by > (u64)ff::P) || (ay > cy && ay - cy > (u64)ff::P)) {
      atomicMin(fail, (unsigned long long)vkey(k_triple, 2, V_PANIC_SUB));
      return;
    }
    const u32 y0 = (u32)(ay % ff::P), y1 = (u32)(by % ff::P), y2 = (u32)(cy % ff::P);
    const u32 dy1 = ff::sub(y1, y0), dx1 = ff::sub(x1, x0);
    const u32 dy2 = ff::sub(y2, y0), dx2 = ff::sub((u32)(x2 % ff::P), x0);
    if (ff::mul(dy1, dx2) != ff::mul(dy2, dx1)) atomicMin(fail, (unsigned long long)vkey(k_triple, 3, V_COLINEAR));
    return;
  }
  // MerkleTree::verify (merkle.rs:82-97) of leaf Hash::from_field_elements(&[y]) at index idx against roots[r (+1)]
  const u64 k_path = base + A.nq + 3ull * s + job;
  if (k_path >= A.n_obj) return;
  const VObj po = A.objs[k_path];
  if (po.tag != 3) return;
  u64 idx = job == 0 ? ai : (job == 1 ? bi : ci);
  u32 cur[8];
  leaf_u64(job == 0 ? ay : (job == 1 ? by : cy), cur);
#pragma unroll 1
  for (u32 j = 0; j < po.count; j++) {
    u32 sib[8], nxt[8];
    load_hash_bytes(A.proof + po.off + 32ull * j, sib);
    if (idx & 1)
      hs::combine(sib, cur, nxt);
    else
      hs::combine(cur, sib, nxt);
#pragma unroll
    for (int g = 0; g < 8; g++) cur[g] = nxt[g];
    idx >>= 1;
  }
  const u32 *root = reinterpret_cast<const u32 *>(A.roots + 32 * (r + (job == 2 ? 1u : 0u)));
  bool same = true;
#pragma unroll
  for (int g = 0; g < 8; g++) same &= cur[g] == root[g];
  if (!same) atomicMin(fail, (unsigned long long)vkey(k_path, 1, V_PATH_A + job));
}

// ProofStream::deserialize (stream.rs:66-168): tags and counts only; truncated items are dropped like the reference
void parse_stream(const u8 *b, size_t n, std::vector<VObj> &objs) {
  size_t i = 0;
  auto rd = [&](size_t at) {
    u64 v = 0;
    for (int k = 0; k < 8; k++) v |= (u64)b[at + k] << (8 * k);
    return v;
  };
  while (i < n) {
    const u8 tag = b[i++];
    if (tag == 0) {
      if (i + 32 <= n) objs.push_back(VObj{0, 1, i}), i += 32;
    } else if (tag == 1) {
      if (i + 8 <= n) objs.push_back(VObj{1, 1, i}), i += 8;
    } else if (tag == 2 || tag == 3) {
      if (i + 8 <= n) {
        const u64 len = rd(i);
        i += 8;
        const size_t item = tag == 2 ? 8 : 32;
        const u64 fit = (n - i) / item, cnt = len < fit ? len : fit;
        objs.push_back(VObj{tag, (u32)(cnt > 0xffffffffull ? 0xffffffffull : cnt), i});
        i += cnt * item;
      }
    } else {
      break;
    }
  }
}

}  // namespace

extern "C" {

const char *stark_fri_verify_reason(uint32_t reason) {
  static const char *const TXT[] = {
      "",
      "Failed to extract Merkle root",
      "Failed to extract last codeword",
      "No FRI roots extracted",
      "last codeword is not well formed",
      "last codeword too small",
      "re-evaluated codeword does not match original!",
      "last codeword does not correspond to polynomial of low enough degree",
      "Failed to extract triple values",
      "Expected triple of values",
      "colinearity check failure",
      "merkle authentication path verification fails for aa",
      "merkle authentication path verification fails for bb",
      "merkle authentication path verification fails for cc",
      "Failed to extract path for aa",
      "Failed to extract path for bb",
      "Failed to extract path for cc",
  };
  return reason < sizeof TXT / sizeof TXT[0] ? TXT[reason] : "unknown";
}

int stark_fri_verify(stark_ctx *ctx, const uint8_t *proof, size_t proof_len, size_t domain_length, uint64_t offset,
                     uint64_t omega, uint32_t ef, uint32_t nq, const uint8_t *transcript, size_t transcript_len, int *ok,
                     uint32_t *reason, uint8_t *roots_out, uint64_t *top_indices, uint64_t *poly_indices,
                     uint64_t *poly_values) {
  if (!ctx || (!proof && proof_len) || !ok || (!transcript && transcript_len))
    return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  // Fri::new asserts (fri.rs:37-45) and num_rounds (fri.rs:93-103)
  const size_t N = domain_length;
  if (N == 0 || (N & (N - 1))) return stark_fail(ctx, STARK_ERR_ARG, "Domain length must be power of 2");
  if (ef == 0 || (ef & (ef - 1))) return stark_fail(ctx, STARK_ERR_ARG, "Expansion factor must be power of 2");
  if (ef < 4) return stark_fail(ctx, STARK_ERR_ARG, "Expansion factor must be at least 4");
  u32 R = 0;
  for (size_t len = N; len > ef && 4 * (size_t)nq < len; len >>= 1) R++;
  int log_n = 0;
  while (((size_t)1 << log_n) < N) log_n++;
  const u32 off = (u32)(offset % ff::P), om = (u32)(omega % ff::P);
  // the last coset is interpolated with the inverse NTT, which is built on FiniteField::prim_nth_root (ff.rs:215-223)
  if (log_n > ff::TWO_ADICITY || om != ff::pow(ff::GEN, (ff::P - 1) >> log_n))
    return stark_fail(ctx, STARK_ERR_ARG, "unsupported domain: omega must be FiniteField::prim_nth_root(domain_length)");
  *ok = 0;
  if (reason) *reason = 0;

  std::vector<VObj> objs;
  parse_stream(proof, proof_len, objs);
  u64 fail = NO_FAIL;
  auto note = [&](u64 k) { if (k < fail) fail = k; };
  auto finish = [&](void) {
    *ok = fail == NO_FAIL;
    if (reason) *reason = (u32)(fail == NO_FAIL ? 0 : fail & 0xff);
    return STARK_OK;
  };
  // ---- structure (host): fri.rs:324-343
  for (u32 r = 0; r < R; r++)
    if (r >= objs.size() || objs[r].tag != 0) note(vkey(r, 0, V_ROOT_EXTRACT));
  if (R >= objs.size() || objs[R].tag != 2) note(vkey(R, 0, V_LAST_EXTRACT));
  if (fail != NO_FAIL) return finish();
  if (R == 0) {
    note(vkey(R, 1, V_NO_ROOTS));
    return finish();
  }
  const u64 ln = objs[R].count;
  if (ln == 0) return stark_fail(ctx, STARK_ERR_ARG, "Cannot create tree from empty leaves");        // merkle.rs:12
  if (ln & (ln - 1)) return stark_fail(ctx, STARK_ERR_ARG, "Number of leaves must be power of 2");   // merkle.rs:13-16
  const u64 expect_last = (u64)N >> (R - 1);
  const u64 degree_bound = ln / ef;
  if (degree_bound == 0) note(vkey(R, 3, V_LAST_SMALL));
  // fri.rs:183-192 (asserts inside sample_indices; reached only when every earlier check passed)
  const u64 reduced = expect_last, sample_size = (u64)N >> 1;
  u64 sample_panic = NO_FAIL;
  if ((u64)nq > 2 * reduced)
    sample_panic = vkey(R, 7, V_PANIC_SAMPLE_ENTROPY);
  else if ((u64)nq > reduced)
    sample_panic = vkey(R, 7, V_PANIC_SAMPLE_COUNT);
  if (sample_panic == NO_FAIL) {
    for (u32 r = 0; r + 1 < R; r++) {
      const u64 base = (u64)R + 1 + (u64)r * 4 * nq;
      for (u32 s = 0; s < nq; s++) {
        const u64 k = base + s;
        if (k >= objs.size() || objs[k].tag != 2) note(vkey(k, 0, V_TRIPLE_EXTRACT));
        else if (objs[k].count != 3) note(vkey(k, 1, V_TRIPLE_LEN));
      }
      for (u32 s = 0; s < nq; s++)
        for (u32 j = 0; j < 3; j++) {
          const u64 k = base + nq + 3ull * s + j;
          if (k >= objs.size() || objs[k].tag != 3) note(vkey(k, 0, V_PATH_A_EXTRACT + j));
        }
    }
  } else {
    note(sample_panic);
  }

  // ---- payload checks (device)
  u8 *d_proof = nullptr, *d_roots_in = nullptr, *d_roots = nullptr, *d_nodes = nullptr;
  VObj *d_objs = nullptr;
  u64 *d_alpha_raw = nullptr, *d_fail = nullptr, *d_top = nullptr, *d_seed = nullptr, *d_poly = nullptr;
  u32 *d_alpha_m = nullptr, *d_vals = nullptr, *d_coeffs = nullptr;
  TranscriptDev *d_tr = nullptr;
  int rc = STARK_OK;
#define TRY_(e) \
  if (rc == STARK_OK) rc = (e);
#define CU_(e) \
  if (rc == STARK_OK && (e) != cudaSuccess) rc = stark_fail(ctx, STARK_ERR_CUDA, "%s failed: %s", #e, cudaGetErrorString(cudaGetLastError()));
  TRY_(dev_alloc(ctx, (void **)&d_proof, proof_len));
  TRY_(dev_alloc(ctx, (void **)&d_objs, objs.size() * sizeof(VObj)));
  TRY_(dev_alloc(ctx, (void **)&d_roots_in, 32 * (size_t)R));
  TRY_(dev_alloc(ctx, (void **)&d_roots, 32 * (size_t)R));
  TRY_(dev_alloc(ctx, (void **)&d_alpha_raw, 8 * (size_t)R));
  TRY_(dev_alloc(ctx, (void **)&d_alpha_m, 4 * (size_t)R));
  TRY_(dev_alloc(ctx, (void **)&d_tr, sizeof(TranscriptDev)));
  TRY_(dev_alloc(ctx, (void **)&d_fail, 8));
  TRY_(dev_alloc(ctx, (void **)&d_seed, 8));
  TRY_(dev_alloc(ctx, (void **)&d_top, 8 * (size_t)(nq ? nq : 1)));
  TRY_(dev_alloc(ctx, (void **)&d_poly, 32 * (size_t)(nq ? nq : 1)));
  TRY_(dev_alloc(ctx, (void **)&d_nodes, 32 * (2 * (size_t)ln)));
  TRY_(dev_alloc(ctx, (void **)&d_vals, 4 * (size_t)ln));
  TRY_(dev_alloc(ctx, (void **)&d_coeffs, 4 * (size_t)ln));
  TranscriptDev t;
  tr_init(t);
  if (transcript_len) tr_absorb(t, transcript, transcript_len);   // the caller's FiatShamir state
  CU_(cudaMemcpyAsync(d_proof, proof, proof_len, cudaMemcpyHostToDevice, ctx->stream));
  CU_(cudaMemcpyAsync(d_objs, objs.data(), objs.size() * sizeof(VObj), cudaMemcpyHostToDevice, ctx->stream));
  CU_(cudaMemcpyAsync(d_tr, &t, sizeof t, cudaMemcpyHostToDevice, ctx->stream));
  CU_(cudaMemsetAsync(d_fail, 0xff, 8, ctx->stream));
  if (rc == STARK_OK)
    LAUNCH(ctx, "verify_transcript", 0, k_verify_transcript<<<1, 32, 0, ctx->stream>>>(d_proof, d_objs, R, d_tr, d_roots_in, d_roots, d_alpha_raw, d_alpha_m));
  if (rc == STARK_OK)
    LAUNCH(ctx, "verify_last_leaves", 0, k_verify_last_leaves<<<(u32)((ln + 63) / 64), 64, 0, ctx->stream>>>(d_proof + objs[R].off, ln, d_nodes, d_vals));
  TRY_(merkle_climb_dev(ctx, d_nodes, (size_t)ln, nullptr));
  // interpolant of the last codeword on last_offset * last_omega^i (fri.rs:362-383); the reference's re-evaluation
  // check (fri.rs:385-391) holds for every exact interpolation and needs no work here
  const bool check_degree = ln == expect_last && degree_bound > 0;
  if (rc == STARK_OK && check_degree) {
    u32 last_off = off;
    for (u32 k = 0; k + 1 < R; k++) last_off = ff::mul(last_off, last_off);
    if (last_off == 0) {
      rc = stark_fail(ctx, STARK_ERR_ARG, "no inverse");   // all points coincide (interpolate.rs:33)
    } else {
      int log_ln = 0;
      while ((1ull << log_ln) < ln) log_ln++;
      ScaleSpec none = {ntt::SCALE_NONE, 1, 1};
      ScaleSpec post = {ntt::SCALE_GEO, ff::inv((u32)(ln % ff::P)), ff::inv(last_off)};
      rc = ntt_transform(ctx, d_vals, d_coeffs, log_ln, true, 1, ln, ln, ln, none, post);
    }
  }
  if (rc == STARK_OK)
    LAUNCH(ctx, "verify_last", 0, k_verify_last<<<(u32)((ln + 255) / 256), 256, 0, ctx->stream>>>(
        d_nodes + 32 * (2 * (size_t)ln - 2), d_roots, R, d_coeffs, ln, degree_bound ? degree_bound - 1 : 0, check_degree ? 1 : 0, d_fail));
  const bool run_rounds = sample_panic == NO_FAIL && nq > 0;
  if (rc == STARK_OK && run_rounds) {
    TRY_(fri_sample_indices_dev(ctx, d_tr, d_seed, sample_size, reduced, nq, d_top));
    if (rc == STARK_OK && R > 1) {
      RoundsArgs A = {d_proof, d_objs, (u64)objs.size(), d_roots, d_alpha_raw, d_top, (u64)N, R, nq,
                      ff::to_mont(off), ff::to_mont(om), d_fail, poly_indices && poly_values ? d_poly : nullptr,
                      poly_indices && poly_values ? d_poly + 2 * (size_t)nq : nullptr};
      LAUNCH_PDL(ctx, "verify_rounds", 0, k_verify_rounds, dim3((4 * nq + 63) / 64, R - 1), 64, A);
    }
  }
  u64 dev_fail = NO_FAIL;
  std::vector<u64> poly(4 * (size_t)nq);
  CU_(cudaMemcpyAsync(&dev_fail, d_fail, 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (roots_out) CU_(cudaMemcpyAsync(roots_out, d_roots, 32 * (size_t)R, cudaMemcpyDeviceToHost, ctx->stream));
  if (run_rounds && top_indices) CU_(cudaMemcpyAsync(top_indices, d_top, 8 * (size_t)nq, cudaMemcpyDeviceToHost, ctx->stream));
  if (run_rounds && R > 1 && poly_indices && poly_values)
    CU_(cudaMemcpyAsync(poly.data(), d_poly, 32 * (size_t)nq, cudaMemcpyDeviceToHost, ctx->stream));
  CU_(cudaStreamSynchronize(ctx->stream));
  dev_free(ctx, d_proof), dev_free(ctx, d_objs), dev_free(ctx, d_roots_in), dev_free(ctx, d_roots);
  dev_free(ctx, d_alpha_raw), dev_free(ctx, d_alpha_m), dev_free(ctx, d_tr), dev_free(ctx, d_fail), dev_free(ctx, d_seed);
  dev_free(ctx, d_top), dev_free(ctx, d_poly), dev_free(ctx, d_nodes), dev_free(ctx, d_vals), dev_free(ctx, d_coeffs);
#undef TRY_
#undef CU_
  if (rc != STARK_OK) return rc;
  note(dev_fail);
  if (ln != expect_last && fail > vkey(R, 4, 0xff))
    return stark_fail(ctx, STARK_ERR_ARG, "unsupported: last codeword of unexpected length whose root matches");
  if (fail != NO_FAIL) {
    const u32 why = (u32)(fail & 0xff);
    if (why == V_PANIC_SAMPLE_ENTROPY) return stark_fail(ctx, STARK_ERR_ARG, "not enough entropy in indices wrt last codeword");
    if (why == V_PANIC_SAMPLE_COUNT)
      return stark_fail(ctx, STARK_ERR_ARG, "cannot sample more indices than available in last codeword; requested: %u, available: %llu", nq, (unsigned long long)reduced);
    if (why == V_PANIC_SUB) return stark_fail(ctx, STARK_ERR_ARG, "attempt to subtract with overflow");
  }
  // polynomial_values is filled in query round 0 only (fri.rs:437-441): with num_rounds() <= 1 the reference pushes nothing
  if (fail == NO_FAIL && R > 1 && poly_indices && poly_values)
    for (size_t i = 0; i < 2 * (size_t)nq; i++) poly_indices[i] = poly[i], poly_values[i] = poly[2 * (size_t)nq + i];
  return finish();
}

}  // extern "C"
