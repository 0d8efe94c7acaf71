// merkle.cu -- leaf hashing and Merkle commitment kernels (reference src/hash.rs:7-46, src/merkle.rs:11-80,
// and the leaf rule of src/fri.rs:118-121).
//
// Tree storage: ONE device array of (2n-1) hashes, level l (n >> l nodes) at hash offset 2n - 2(n >> l):
// level 0 = leaves ... last = root.  That is MerkleTree.nodes (merkle.rs:18-29) flattened, so open()
// (merkle.rs:67-80) is a gather and nothing is ever rebuilt (the reference rebuilds every tree in the
// query phase, fri.rs:288-298).
//
// Kernels (hash.cuh): wide levels hash two nodes per thread (hs2, bound by the ALU pipe); levels with fewer than
// 2^17 parents are latency-bound -- a node hash is a ~2 k-instruction dependency chain -- so there one CTA climbs up to
// 10 levels of its 1024-node chunk through shared memory (k_merkle_climb), one hash per thread in the narrow steps.
// The kernel that produces the root can also absorb it into the Fiat-Shamir transcript and draw alpha (transcript.cuh),
// so a FRI round has no separate transcript launch.
#include "common.cuh"
#include "hash.cuh"
#include "merkle.h"
#include "merkle_dev.cuh"

using hs::State;

// The hs2 kernels have no shared memory and no barrier, so small CTAs cost nothing and spread a mid-sized level
// (a few hundred thousand hashes) evenly over the 148 SMs: 64 threads = 128 hashes per CTA.
constexpr int HASH_NT = 64;

// leaf i = Hash::from_field_elements(&[vals[i]])  (fri.rs:118-121, hash.rs:32-35); two leaves per thread (hs2)
// blockIdx.y = tree of a batch of equally sized trees (column c of an LDE matrix -> tree c): strides in elements / bytes
__global__ void __launch_bounds__(HASH_NT) k_leaf_hash1(const u32 *__restrict__ vals, size_t n, u8 *__restrict__ out,
                                                        size_t val_stride, size_t out_stride) {
  pdl_entry();
  const size_t i = 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n) return;
  vals += (size_t)blockIdx.y * val_stride, out += (size_t)blockIdx.y * out_stride;
  const bool two = i + 1 < n;
  u32 va, vb = 0;
  if (two) {
    const uint2 v = *reinterpret_cast<const uint2 *>(vals + i);
    va = v.x, vb = v.y;
  } else {
    va = vals[i];
  }
  u32 wa[8], wb[8];
  hs2::leaf2(va, vb, wa, wb, blockDim.y);
  store_hash(out + 32 * i, wa);
  if (two) store_hash(out + 32 * i + 32, wb);
}

// leaves i, i+1 = Hash::from_field_elements(&[vals[i*row_stride + c*col_stride] for c < width])  (hash.rs:32-35),
// two rows per thread (hs2).  One 32-byte chunk = 4 values (LE u64 each,
// high word zero).  This is the leaf rule of BASELINE config 4 (8 trace columns per Merkle leaf).
__global__ void __launch_bounds__(HASH_NT) k_leaf_hashw2(const u32 *__restrict__ vals, size_t n, u32 width, size_t row_stride,
                                                         size_t col_stride, u8 *__restrict__ out) {
  pdl_entry();
  const size_t i = 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n) return;
  const bool two = i + 1 < n;
  const u32 *ba = vals + i * row_stride, *bb = vals + (two ? i + 1 : i) * row_stride;
  hs2::State2 st;
  hs2::init(st, blockDim.y);
  bool pending = false;
  for (u32 c0 = 0; c0 < width; c0 += 4) {
    const int nv = (width - c0) < 4 ? (int)(width - c0) : 4;
    u32 va[4], vb[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      va[k] = k < nv ? ba[(size_t)(c0 + k) * col_stride] : 0u;
      vb[k] = k < nv ? bb[(size_t)(c0 + k) * col_stride] : 0u;
    }
    if (pending) hs2::settle(st);
#pragma unroll
    for (int b = 0; b < 32; b++)
      if (b < 8 * nv) hs2::absorb_pair(st, b, (b & 4) ? 0u : hs2::pair_bytes(va[b >> 3], vb[b >> 3], b & 3));
    hs2::mix_lazy<false>(st);
    pending = true;
  }
  if (pending)
    hs2::finalize<true>(st);
  else
    hs2::finalize<false>(st);
  u32 wa[8], wb[8];
  hs2::pack_words(st, wa, wb);
  store_hash(out + 32 * i, wa);
  if (two) store_hash(out + 32 * i + 32, wb);
}

// generic Hash::from_bytes of n messages of msg_len bytes (hash.rs:7-30); state stays in registers
__global__ void __launch_bounds__(128) k_hash_bytes(const u8 *__restrict__ msgs, size_t n, size_t msg_len,
                                                    u8 *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u8 *m = msgs + i * msg_len;
  State st;
  hs::init(st);
  bool pending = false;
  for (size_t off = 0; off < msg_len; off += 32) {
    const int len = (msg_len - off) < 32 ? (int)(msg_len - off) : 32;
    if (pending) hs::settle(st);
#pragma unroll
    for (int b = 0; b < 32; b++)
      if (b < len) hs::absorb_byte(st, b, m[off + b]);
    hs::mix_lazy<false>(st);
    pending = true;
  }
  if (pending)
    hs::finalize<true>(st);
  else
    hs::finalize<false>(st);
  u32 w[8];
  hs::pack_words(st, w);
  u8 *o = out + 32 * i;
#pragma unroll
  for (int g = 0; g < 8; g++) reinterpret_cast<u32 *>(o)[g] = w[g];
}

// one tree level: parent i = Hash::combine(child 2i, child 2i+1)  (merkle.rs:21-27); two parents per thread
__global__ void __launch_bounds__(HASH_NT) k_merkle_level(const u8 *__restrict__ in, u8 *__restrict__ out, size_t n_out,
                                                          size_t tree_stride) {
  pdl_entry();
  const size_t i = 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n_out) return;
  in += (size_t)blockIdx.y * tree_stride, out += (size_t)blockIdx.y * tree_stride;
  const bool two = i + 1 < n_out;
  u32 la[8], ra[8], lb[8], rb[8], wa[8], wb[8];
  load_hash(in + 64 * i, la);
  load_hash(in + 64 * i + 32, ra);
  load_hash(in + 64 * (two ? i + 1 : i), lb);
  load_hash(in + 64 * (two ? i + 1 : i) + 32, rb);
  hs2::combine2(la, ra, lb, rb, wa, wb, blockDim.y);
  store_hash(out + 32 * i, wa);
  if (two) store_hash(out + 32 * i + 32, wb);
}

// CTA b climbs `levels` levels from the `cnt` nodes [b cnt, (b+1) cnt) of level `level_in` (cta_climb).  When the
// climb ends at the root and tr.T is set, thread 0 runs the transcript round (fri.rs:129-138) on it.
//   <256>: many CTAs, chunks of 512 nodes, 9 levels per launch (two hs2 steps, then seven 4-lanes-per-hash steps)
//   <512>: the single top CTA, up to 1024 nodes
//   X.world > 1: the tree is one rank's SUBTREE of a sharded tree -- the CTA that ends up with the subtree root exchanges
//   it with the peers and climbs the replicated top levels before the transcript round (mg_exchange_top)
template <int NT>
__global__ void __launch_bounds__(NT) k_merkle_climb(u8 *nodes, size_t n, u32 level_in, u32 cnt, u32 levels,
                                                     TranscriptArgs tr, u32 *counter, const __grid_constant__ MgExchange X,
                                                     size_t tree_stride) {
  __shared__ __align__(16) u8 sm[2 * NT * 32];
  __shared__ u32 ticket;
  pdl_entry();
  nodes += (size_t)blockIdx.y * tree_stride;   // batch of equally sized trees: one ticket per tree
  if (counter != nullptr) counter += blockIdx.y;
  const u32 t = threadIdx.x;
  const size_t first = (size_t)blockIdx.x * cnt;
  cta_climb<NT>(nodes, n, level_in, first, cnt, levels, nodes + 32 * (level_off(n, level_in) + first), sm, blockDim.y);
  u32 level = level_in + levels;
  if (counter != nullptr && gridDim.x > 1) {
    // Fused top: the LAST CTA to finish its chunk climbs the gridDim.x chunk roots to the tree root in this same launch
    // (no second launch, and the hash code is already in this SM's instruction cache).
    __threadfence();
    __syncthreads();
    if (t == 0) ticket = atomicAdd(counter, 1u);
    __syncthreads();
    if (ticket != gridDim.x - 1) return;
    if (t == 0) *counter = 0u;   // ready for the next launch on this stream
    __threadfence();
    const u32 m = gridDim.x;     // <= NT nodes: 32 m bytes fit in sm
    const uint4 *src = reinterpret_cast<const uint4 *>(nodes + 32 * level_off(n, level));
    for (u32 i = t; i < 2 * m; i += NT) reinterpret_cast<uint4 *>(sm)[i] = __ldcg(src + i);   // written by other SMs
    __syncthreads();
    u32 top_levels = 0;
    for (u32 c = m; c > 1; c >>= 1) top_levels++;
    cta_climb<NT>(nodes, n, level, 0, m, top_levels, sm, sm, blockDim.y);
    level += top_levels;
  }
  if ((n >> level) != 1) return;
  if (X.world > 1 && !mg_exchange_top<NT>(X, sm, blockDim.y)) return;
  if (tr.T != nullptr && t < 32) transcript_round_warp(tr, sm);
}
// lock-step groups (virtual ranks on one device): the wait / top / transcript half as its own launch
__global__ void __launch_bounds__(128) k_mg_top(TranscriptArgs tr, const __grid_constant__ MgExchange X) {
  __shared__ __align__(16) u8 sm[2 * 128 * 32];
  pdl_entry();
  mg_exchange_top<128>(X, sm, blockDim.y);
  if (tr.T != nullptr && threadIdx.x < 32) transcript_round_warp(tr, sm);
}
// the root of a one-leaf tree is the leaf itself (merkle.rs:11-38 with n = 1)
__global__ void k_transcript_only(const u8 *root_hash, TranscriptArgs tr) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  u32 root[8];
  load_hash(root_hash, root);
  transcript_round(tr, root);
}

// gather authentication paths (merkle.rs:67-80): out[(q*depth + l)*32 ..] = nodes[level l][(idx[q] >> l) ^ 1]
__global__ void k_merkle_open(const u8 *__restrict__ nodes, size_t n, u32 depth, const u64 *__restrict__ idx, u32 n_idx,
                              u8 *__restrict__ out) {
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 16-byte half hash
  const u32 item = t >> 1, half = t & 1;
  if (item >= n_idx * depth) return;
  const u32 q = item / depth, l = item % depth;
  const size_t sib = (size_t)(idx[q] >> l) ^ 1;
  const uint4 *src = reinterpret_cast<const uint4 *>(nodes + 32 * ((2 * n - 2 * (n >> l)) + sib));
  reinterpret_cast<uint4 *>(out + 32 * (size_t)item)[half] = src[half];
}

// ------------------------------------------------------------------------------------------ device API

int merkle_check_n(stark_ctx *ctx, size_t n) {
  if (n == 0) return stark_fail(ctx, STARK_ERR_ARG, "Cannot create tree from empty leaves");     // merkle.rs:12
  if (n & (n - 1)) return stark_fail(ctx, STARK_ERR_ARG, "Number of leaves must be power of 2");  // merkle.rs:13-16
  return STARK_OK;
}

int merkle_leaves_dev(stark_ctx *ctx, const u32 *vals, size_t n, u32 width, size_t row_stride, size_t col_stride,
                      u8 *out) {
  if (n == 0) return STARK_OK;
  if (width == 1)
    LAUNCH_PDL(ctx, "leaf_hash", 36ull * n, k_leaf_hash1, (u32)((n + 2 * HASH_NT - 1) / (2 * HASH_NT)), HASH_NT, vals, n, out,
               (size_t)0, (size_t)0);
  else
    LAUNCH_PDL(ctx, "leaf_hash_w", (4ull * width + 32) * n, k_leaf_hashw2, (u32)((n + 2 * HASH_NT - 1) / (2 * HASH_NT)), HASH_NT,
               vals, n, width, row_stride, col_stride, out);
  return STARK_OK;
}

// nodes[0 .. n) already holds the leaves; fill the upper levels.  tr (optional): transcript round on the root.
int merkle_climb_dev(stark_ctx *ctx, u8 *nodes, size_t n, const TranscriptArgs *tr, const MgExchange *mx) {
  return merkle_climb_batch_dev(ctx, nodes, n, 1, 0, tr, mx);
}
// `batch` equally sized trees, tree b at nodes + b * tree_stride bytes: every level of all of them in one launch, and ONE
// climb launch whose grid covers all trees -- the latency-bound top levels of the trees overlap instead of queueing
int merkle_climb_batch_dev(stark_ctx *ctx, u8 *nodes, size_t n, u32 batch, size_t tree_stride, const TranscriptArgs *tr,
                           const MgExchange *mx) {
  const TranscriptArgs none = {nullptr, nullptr, 0, nullptr, nullptr};
  MgExchange X;
  memset(&X, 0, sizeof X);
  if (mx) X = *mx;
  if (X.world > 1 && n < 2) return stark_fail(ctx, STARK_ERR_ARG, "a sharded tree needs at least two leaves per rank");
  if (batch == 0) return STARK_OK;
  if (batch > (u32)CLIMB_TICKETS || (batch > 1 && (tr || mx)))
    return stark_fail(ctx, STARK_ERR_ARG, "unsupported tree batch");
  u32 level = 0;
  size_t m = n;
  // throughput-bound levels: one launch each.  A level of 2^17 parents is already half latency (12 us against 6 us for the
  // same step inside the climb kernel, whose launch is paid anyway), so the climb starts from 2^18 nodes.
  // STARK_CLIMB_LOG (read once at context creation, clamped to 11..18).  A batch of trees has enough hashes per level to
  // stay throughput-bound much further up: its levels run as full-width launches down to 2^13 nodes per tree, and the
  // climb kernel -- whose CTAs are latency-bound and occupancy-limited (2 per SM) -- only takes the top 13 levels
  const int batch_log = ctx->climb_log < 13 ? ctx->climb_log : 13;
  const size_t climb_from = (size_t)1 << (batch >= 4 ? batch_log : ctx->climb_log);
  while (m > climb_from) {
    const size_t half = m >> 1;
    LAUNCH_PDL(ctx, "merkle_level", 96ull * half * batch, k_merkle_level, dim3((u32)((half + 2 * HASH_NT - 1) / (2 * HASH_NT)), batch),
               HASH_NT, (const u8 *)(nodes + 32 * (2 * n - 2 * m)), nodes + 32 * (2 * n - 2 * half), half, tree_stride);
    m = half;
    level++;
  }
  // latency-bound levels (m <= 2^17 nodes left): chunks of 512 nodes climb 9 levels each and the last CTA to finish
  // climbs the <= 256 chunk roots to the root; small trees are a single CTA
  if (m > 1024) {
    // at most 256 chunks (their roots are climbed by the last CTA): 1024-node chunks (10 levels) above 2^17 nodes
    const u32 chunk = m > ((size_t)1 << 17) ? 1024u : 512u, chunk_levels = chunk == 1024u ? 10u : 9u;
    const size_t ctas = m / chunk;
    LAUNCH_PDL(ctx, "merkle_climb", 96ull * (m - 1) * batch, k_merkle_climb<256>, dim3((u32)ctas, batch), 256, nodes, n, level, chunk,
               chunk_levels, tr ? *tr : none, ctx->climb_counter, X, tree_stride);
  } else if (m > 1) {
    u32 levels = 0;
    for (size_t c = m; c > 1; c >>= 1) levels++;
    LAUNCH_PDL(ctx, "merkle_top", 96ull * (m - 1) * batch, k_merkle_climb<512>, dim3(1u, batch), 512, nodes, n, level, (u32)m, levels,
               tr ? *tr : none, (u32 *)nullptr, X, tree_stride);
  }
  if (n == 1 && tr) LAUNCH(ctx, "transcript", 0, k_transcript_only<<<1, 32, 0, ctx->stream>>>(nodes, *tr));
  return STARK_OK;
}
// the second half of a sharded tree's root step for lock-step groups (MG_X_WAIT)
int merkle_mg_top_dev(stark_ctx *ctx, const TranscriptArgs *tr, const MgExchange *mx) {
  const TranscriptArgs none = {nullptr, nullptr, 0, nullptr, nullptr};
  LAUNCH_PDL(ctx, "mg_top", 0, k_mg_top, 1u, 128, tr ? *tr : none, *mx);
  return STARK_OK;
}

int merkle_tree_alloc(stark_ctx *ctx, size_t n, stark_tree **out) {
  ST_TRY(merkle_check_n(ctx, n));
  stark_tree *t = new stark_tree();
  t->ctx = ctx, t->n = n, t->levels = 1;
  for (size_t m = n; m > 1; m >>= 1) t->levels++;
  t->nodes = nullptr;
  int rc = dev_alloc(ctx, (void **)&t->nodes, (2 * n - 1) * 32);
  if (rc != STARK_OK) {
    delete t;
    return rc;
  }
  *out = t;
  return STARK_OK;
}

int merkle_open_dev(stark_ctx *ctx, const u8 *nodes, size_t n, const u64 *idx_dev, u32 n_idx, u8 *out_dev) {
  u32 depth = 0;
  for (size_t m = n; m > 1; m >>= 1) depth++;
  if (depth == 0 || n_idx == 0) return STARK_OK;
  const u32 threads = n_idx * depth * 2;
  LAUNCH(ctx, "merkle_open", 0, k_merkle_open<<<(threads + 127) / 128, 128, 0, ctx->stream>>>(nodes, n, depth, idx_dev, n_idx, out_dev));
  return STARK_OK;
}

// ----------------------------------------------------------------------------------------------- C ABI

extern "C" {

int stark_hash_bytes(stark_ctx *ctx, const uint8_t *msgs, size_t n_msgs, size_t msg_len, uint8_t *out) {
  if (!ctx || (!msgs && n_msgs && msg_len) || !out) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (n_msgs == 0) return STARK_OK;
  u8 *d_in = nullptr, *d_out = nullptr;
  Scratch sc(ctx);
  ST_TRY(sc.get(&d_in, n_msgs * msg_len));
  ST_TRY(sc.get(&d_out, n_msgs * 32));
  if (msg_len) CU_TRY(ctx, cudaMemcpyAsync(d_in, msgs, n_msgs * msg_len, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(ctx, "hash_bytes", (msg_len + 32) * n_msgs,
         k_hash_bytes<<<(u32)((n_msgs + 127) / 128), 128, 0, ctx->stream>>>(d_in, n_msgs, msg_len, d_out));
  CU_TRY(ctx, cudaMemcpyAsync(out, d_out, n_msgs * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return STARK_OK;
}

int stark_hash_leaves(stark_ctx *ctx, const uint64_t *vals, size_t n_leaves, uint32_t width, uint8_t *out) {
  if (!ctx || !out || (!vals && n_leaves)) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (n_leaves == 0) return STARK_OK;
  stark_buf *b = nullptr;
  ST_TRY(stark_buf_upload(ctx, vals, n_leaves * (size_t)width, &b));
  u8 *d_out = nullptr;
  int rc = dev_alloc(ctx, (void **)&d_out, n_leaves * 32);
  if (rc == STARK_OK) rc = merkle_leaves_dev(ctx, (const u32 *)stark_buf_ptr(b), n_leaves, width, width, 1, d_out);
  if (rc == STARK_OK && cudaMemcpyAsync(out, d_out, n_leaves * 32, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
    rc = stark_fail(ctx, STARK_ERR_CUDA, "D2H copy failed");
  dev_free(ctx, d_out);
  stark_buf_free(b);
  if (rc == STARK_OK) CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return rc;
}

int stark_merkle_build(stark_ctx *ctx, const uint8_t *leaves, size_t n, stark_tree **out) {
  if (!ctx || !out) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  stark_tree *t = nullptr;
  ST_TRY(merkle_tree_alloc(ctx, n, &t));
  int rc = STARK_OK;
  if (cudaMemcpyAsync(t->nodes, leaves, n * 32, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
    rc = stark_fail(ctx, STARK_ERR_CUDA, "H2D copy failed");
  if (rc == STARK_OK) rc = merkle_climb_dev(ctx, t->nodes, n);
  if (rc != STARK_OK) {
    stark_merkle_free(t);
    return rc;
  }
  *out = t;
  return STARK_OK;
}

}  // extern "C"
int merkle_build_from_dev_values(stark_ctx *ctx, const u32 *vals, size_t n, u32 width, size_t row_stride,
                                 size_t col_stride, stark_tree **out, const TranscriptArgs *tr) {
  stark_tree *t = nullptr;
  ST_TRY(merkle_tree_alloc(ctx, n, &t));
  int rc = merkle_leaves_dev(ctx, vals, n, width, row_stride, col_stride, t->nodes);
  if (rc == STARK_OK) rc = merkle_climb_dev(ctx, t->nodes, n, tr);
  if (rc != STARK_OK) {
    stark_merkle_free(t);
    return rc;
  }
  *out = t;
  return STARK_OK;
}

// `batch` trees over the columns of a column-major matrix (tree b over vals[b * val_stride ..], one value per leaf, the
// rule of fri.rs:118-121), tree b at nodes + b * tree_stride bytes ((2n - 1) * 32 bytes each, rounded up by the caller)
int merkle_build_batch_dev(stark_ctx *ctx, const u32 *vals, size_t n, u32 batch, size_t val_stride, u8 *nodes,
                           size_t tree_stride) {
  if (batch == 0) return STARK_OK;
  ST_TRY(merkle_check_n(ctx, n));
  LAUNCH_PDL(ctx, "leaf_hash", 36ull * n * batch, k_leaf_hash1, dim3((u32)((n + 2 * HASH_NT - 1) / (2 * HASH_NT)), batch), HASH_NT,
             vals, n, nodes, val_stride, tree_stride);
  return merkle_climb_batch_dev(ctx, nodes, n, batch, tree_stride, nullptr, nullptr);
}

extern "C" {
int stark_merkle_build_from_values(stark_ctx *ctx, const uint64_t *vals, size_t n_leaves, uint32_t width,
                                   stark_tree **out) {
  if (!ctx || !out || width == 0) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  ST_TRY(merkle_check_n(ctx, n_leaves));
  stark_buf *b = nullptr;
  ST_TRY(stark_buf_upload(ctx, vals, n_leaves * (size_t)width, &b));
  int rc = merkle_build_from_dev_values(ctx, (const u32 *)stark_buf_ptr(b), n_leaves, width, width, 1, out);
  stark_buf_free(b);
  return rc;
}

int stark_merkle_build_from_buf(stark_ctx *ctx, const stark_buf *vals, size_t n_leaves, uint32_t width,
                                stark_tree **out) {
  if (!ctx || !out || !vals || width == 0) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  ST_TRY(merkle_check_n(ctx, n_leaves));
  if (stark_buf_len(vals) < n_leaves * (size_t)width) return stark_fail(ctx, STARK_ERR_ARG, "buffer too small");
  // column-major [width][n_leaves]
  return merkle_build_from_dev_values(ctx, (const u32 *)stark_buf_ptr(vals), n_leaves, width, 1, n_leaves, out);
}

// MerkleTree::new (merkle.rs:11-38) over leaves that are already on the device (e.g. the gathered subtree roots of
// the sharded prover)
int stark_merkle_build_dev(stark_ctx *ctx, const void *leaves_dev, size_t n, stark_tree **out) {
  if (!ctx || !out || !leaves_dev) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  stark_tree *t = nullptr;
  ST_TRY(merkle_tree_alloc(ctx, n, &t));
  int rc = STARK_OK;
  if (cudaMemcpyAsync(t->nodes, leaves_dev, n * 32, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)
    rc = stark_fail(ctx, STARK_ERR_CUDA, "D2D copy failed");
  if (rc == STARK_OK) rc = merkle_climb_dev(ctx, t->nodes, n);
  if (rc != STARK_OK) {
    stark_merkle_free(t);
    return rc;
  }
  *out = t;
  return STARK_OK;
}

// device address of the flattened node array: level l (n >> l hashes) starts at hash offset 2n - 2(n >> l)
void *stark_merkle_nodes_ptr(const stark_tree *t) { return t ? (void *)t->nodes : nullptr; }

// MerkleTree::open (merkle.rs:67-80) for n_idx leaves at once: out[(q*depth + l)*32 ..] = sibling at level l
int stark_merkle_open_batch(stark_tree *t, const uint64_t *idx, size_t n_idx, uint8_t *out) {
  if (!t || (n_idx && (!idx || !out))) return stark_fail(nullptr, STARK_ERR_ARG, "null argument");
  stark_ctx *ctx = t->ctx;
  for (size_t i = 0; i < n_idx; i++)
    if (idx[i] >= t->n) return stark_fail(ctx, STARK_ERR_ARG, "Index out of bounds");  // merkle.rs:68
  const u32 depth = t->levels - 1;
  if (depth == 0 || n_idx == 0) return STARK_OK;
  u64 *d_idx = nullptr;
  u8 *d_out = nullptr;
  Scratch sc(ctx);
  ST_TRY(sc.get(&d_idx, 8 * n_idx));
  ST_TRY(sc.get(&d_out, 32 * (size_t)depth * n_idx));
  CU_TRY(ctx, cudaMemcpyAsync(d_idx, idx, 8 * n_idx, cudaMemcpyHostToDevice, ctx->stream));
  ST_TRY(merkle_open_dev(ctx, t->nodes, t->n, d_idx, (u32)n_idx, d_out));
  CU_TRY(ctx, cudaMemcpyAsync(out, d_out, 32 * (size_t)depth * n_idx, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return STARK_OK;
}

int stark_merkle_commit(stark_ctx *ctx, const uint8_t *leaves, size_t n, uint8_t root[32]) {
  stark_tree *t = nullptr;
  ST_TRY(stark_merkle_build(ctx, leaves, n, &t));
  int rc = stark_merkle_root(t, root);
  stark_merkle_free(t);
  return rc;
}

int stark_merkle_root(stark_tree *t, uint8_t root[32]) {
  if (!t || !root) return stark_fail(nullptr, STARK_ERR_ARG, "null argument");
  stark_ctx *ctx = t->ctx;
  CU_TRY(ctx, cudaMemcpyAsync(root, t->nodes + 32 * (2 * t->n - 2), 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return STARK_OK;
}

size_t stark_merkle_num_leaves(const stark_tree *t) { return t ? t->n : 0; }
uint32_t stark_merkle_num_levels(const stark_tree *t) { return t ? t->levels : 0; }

int stark_merkle_level(stark_tree *t, uint32_t level, uint8_t *out) {
  if (!t || !out) return stark_fail(nullptr, STARK_ERR_ARG, "null argument");
  stark_ctx *ctx = t->ctx;
  if (level >= t->levels) return stark_fail(ctx, STARK_ERR_ARG, "level out of range");
  const size_t m = t->n >> level;
  CU_TRY(ctx, cudaMemcpyAsync(out, t->nodes + 32 * (2 * t->n - 2 * m), 32 * m, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return STARK_OK;
}

int stark_merkle_open(stark_tree *t, size_t index, uint8_t *out, size_t *n_hashes) {
  if (!t || !out) return stark_fail(nullptr, STARK_ERR_ARG, "null argument");
  stark_ctx *ctx = t->ctx;
  if (index >= t->n) return stark_fail(ctx, STARK_ERR_ARG, "Index out of bounds");  // merkle.rs:68
  const u32 depth = t->levels - 1;
  if (n_hashes) *n_hashes = depth;
  if (depth == 0) return STARK_OK;
  u64 *d_idx = nullptr;
  u8 *d_out = nullptr;
  Scratch sc(ctx);
  ST_TRY(sc.get(&d_idx, 8));
  ST_TRY(sc.get(&d_out, 32 * depth));
  u64 idx = index;
  CU_TRY(ctx, cudaMemcpyAsync(d_idx, &idx, 8, cudaMemcpyHostToDevice, ctx->stream));
  ST_TRY(merkle_open_dev(ctx, t->nodes, t->n, d_idx, 1, d_out));
  CU_TRY(ctx, cudaMemcpyAsync(out, d_out, 32 * depth, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return STARK_OK;
}

void stark_merkle_free(stark_tree *t) {
  if (!t) return;
  if (t->nodes) dev_free(t->ctx, t->nodes);
  delete t;
}

}  // extern "C"
