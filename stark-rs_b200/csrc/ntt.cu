// ntt.cu -- NTT / iNTT / coset-NTT kernels for sm_100a and their launcher.
//
// Replaces the structured-domain uses of Polynomial::eval_domain (reference src/univariate/eval.rs:16-21)
// and Polynomial::interpolate_domain (interpolate.rs:6-44).  See ntt_core.cuh for the round structure.
//
//   N <= 4           one thread per transform (direct DFT)
//   8 <= N <= 4096   one pass, one transform per CTA
//   2^13..2^23       two passes (four-step):  N = N1*N2, x[n1*N2 + n2]
//        pass 1: N1-point transforms down the columns (stride N2), times w_N^(n2*k1), stored TRANSPOSED
//                T[n2*N1 + k1]  (column tile in, contiguous rows out)
//        pass 2: N2-point transforms over n2 at stride N1, in place, X[k2*N1 + k1] -- natural order.
//   HBM traffic: 8N bytes per pass (4 read + 4 written), nothing else (twiddle tables are <= 48 KB, L1/L2).
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "ntt_pass.cuh"

using namespace ntt;

// ------------------------------------------------------------------------------------------- tables

__global__ void k_root_tables(u32 *lo, u32 *hi, u32 w_m /* w23 * R */) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 4096) lo[i] = ff::mont_pow(w_m, i);
  if (i < 2048) hi[i] = ff::mont_pow(w_m, (u64)i << 12);
}
// tw[(1 << logL) + e] = w_L^(+-e) for logL = 0..12
__global__ void k_sub_tables(u32 *tw, RootTables T, int inverse) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 || i >= 8192) {
    if (i == 0) tw[0] = ff::R1;
    return;
  }
  int logL = 31 - __clz(i);
  u32 e = i - (1u << logL);
  u32 idx = e << (23 - logL);
  if (inverse) idx = ((1u << 23) - idx) & ((1u << 23) - 1u);
  tw[i] = root_pow(T, idx);
}
// Shoup form of a Montgomery-form table: out[i] = (w, floor(w 2^32 / p)), w = in[i] / R
__global__ void k_shoup_table(const u32 *in, wpair *out, u32 n) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u32 w = ff::from_mont(in[i]);
  out[i] = wpair{w, ff::shoup_of(w)};
}
// out[i] = w23^(+-(e(i))) in Shoup form: otw table e = i << 7 (i < 2^16); row table i = (logN - 13) * 2048 + row, e = row << (23 - logN)
__global__ void k_shoup_roots(wpair *out, RootTables T, int inverse, int row_table, u32 n) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 idx = row_table ? (i & 2047u) << (10 - (i >> 11)) : i << 7;
  if (inverse) idx = ((1u << 23) - idx) & ((1u << 23) - 1u);
  const u32 w = ff::from_mont(root_pow(T, idx));
  out[i] = wpair{w, ff::shoup_of(w)};
}
// lo[i] = c * g^i (i < 4096), hi[j] = g^(4096 j) (j < hi_len); Montgomery form
__global__ void k_geo_tables(u32 *lo, u32 *hi, u32 g_m, u32 c_m, u32 hi_len) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 4096) lo[i] = ff::canon(ff::mont_mul(ff::mont_pow(g_m, i), c_m));
  if (i < hi_len) hi[i] = ff::mont_pow(g_m, (u64)i << 12);
}

int ntt_init(stark_ctx *ctx) {
  CU_TRY(ctx, cudaMalloc(&ctx->root_lo, 4096 * 4));
  CU_TRY(ctx, cudaMalloc(&ctx->root_hi, 2048 * 4));
  CU_TRY(ctx, cudaMalloc(&ctx->tw_sub[0], 8192 * 4));
  CU_TRY(ctx, cudaMalloc(&ctx->tw_sub[1], 8192 * 4));
  CU_TRY(ctx, cudaMalloc(&ctx->tw_sh[0], 8192 * sizeof(wpair)));
  CU_TRY(ctx, cudaMalloc(&ctx->tw_sh[1], 8192 * sizeof(wpair)));
  const u32 w23 = ff::pow(ff::GEN, (ff::P - 1) >> 23);  // ff.rs:215-223
  k_root_tables<<<16, 256, 0, ctx->stream>>>(ctx->root_lo, ctx->root_hi, ff::to_mont(w23));
  KERNEL_CHECK(ctx);
  RootTables T = {ctx->root_lo, ctx->root_hi};
  for (int d = 0; d < 2; d++) {
    k_sub_tables<<<32, 256, 0, ctx->stream>>>(ctx->tw_sub[d], T, d);
    KERNEL_CHECK(ctx);
    k_shoup_table<<<32, 256, 0, ctx->stream>>>(ctx->tw_sub[d], ctx->tw_sh[d], 8192);
    KERNEL_CHECK(ctx);
    {
      std::vector<wpair> h(ntt2::INNER_TWIDDLE_PAIRS);
      ntt2::fill_inner_twiddles(h.data(), d);
      CU_TRY(ctx, cudaMalloc(&ctx->tw_in_sh[d], h.size() * sizeof(wpair)));
      CU_TRY(ctx, cudaMemcpy(ctx->tw_in_sh[d], h.data(), h.size() * sizeof(wpair), cudaMemcpyHostToDevice));
    }
    CU_TRY(ctx, cudaMalloc(&ctx->otw_sh[d], (1u << 16) * sizeof(wpair)));
    CU_TRY(ctx, cudaMalloc(&ctx->row_sh[d], 11 * 2048 * sizeof(wpair)));
    k_shoup_roots<<<(1u << 16) / 256, 256, 0, ctx->stream>>>(ctx->otw_sh[d], T, d, 0, 1u << 16);
    KERNEL_CHECK(ctx);
    k_shoup_roots<<<11 * 8, 256, 0, ctx->stream>>>(ctx->row_sh[d], T, d, 1, 11 * 2048);
    KERNEL_CHECK(ctx);
    {
      // FIRST-pass row table of the fused 2^12 kernel: w_4096^(+-row), row < 128
      u32 w12 = ff::pow(ff::GEN, (ff::P - 1) >> 12);
      if (d) w12 = ff::inv(w12);
      std::vector<wpair> h(128);
      u32 v = 1;
      for (int r = 0; r < 128; r++) h[r] = wpair{v, ff::shoup_of(v)}, v = ff::mul(v, w12);
      CU_TRY(ctx, cudaMalloc(&ctx->row12_sh[d], h.size() * sizeof(wpair)));
      CU_TRY(ctx, cudaMemcpy(ctx->row12_sh[d], h.data(), h.size() * sizeof(wpair), cudaMemcpyHostToDevice));
    }
    u32 w8 = ff::pow(ff::GEN, (ff::P - 1) >> 3);
    if (d) w8 = ff::inv(w8);
    for (int k = 0; k < 4; k++) ctx->w8_sh[d][k] = wpair{ff::pow(w8, k), ff::shoup_of(ff::pow(w8, k))};
  }
  for (int i = 0; i < 8; i++) ctx->geo[i] = GeoCacheEntry{0, 0, nullptr, nullptr, 0, 0};
  ctx->geo_stamp = 0;
  ctx->ntt_small_off = getenv("STARK_NTT_SMALL_OFF") && atoi(getenv("STARK_NTT_SMALL_OFF")) != 0;
  const char *e = getenv("STARK_NTT_STREAMS");
  ctx->ntt_streams = e ? atoi(e) : 2;
  if (ctx->ntt_streams < 1 || ctx->ntt_streams > 4) ctx->ntt_streams = 2;
  e = getenv("STARK_NTT_GROUP_MB");
  ctx->ntt_group_mb = e ? atoi(e) : 16;
  if (ctx->ntt_group_mb < 1) ctx->ntt_group_mb = 16;
  ctx->n_side = 0;
  e = getenv("STARK_NTT_BIG");
  ctx->ntt_big = e ? atoi(e) : 0;
  e = getenv("STARK_NTT_L2_PERSIST");
  ctx->ntt_l2_persist = e ? atoi(e) : 0;
  ctx->l2_persist_ready = 0;
  return STARK_OK;
}
// L2 persistence for the grouped path: the destination of a pass is read by the next pass of the same group
static void l2_window(stark_ctx *ctx, cudaStream_t st, const void *base, size_t bytes) {
  if (!ctx->l2_persist_ready) {
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
    size_t want = (size_t)ctx->ntt_l2_persist << 20;   // MB of set-aside requested through the knob
    if (want > (size_t)max_persist) want = (size_t)max_persist;
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
    ctx->l2_persist_ready = max_window > 0 ? max_window : -1;
    fprintf(stderr, "[stark] L2 persistence: set-aside %zu MB (max %d MB), window max %d MB\n", want >> 20, max_persist >> 20, max_window >> 20);
  }
  if (ctx->l2_persist_ready <= 0) return;
  cudaStreamAttrValue v;
  memset(&v, 0, sizeof v);
  v.accessPolicyWindow.base_ptr = const_cast<void *>(base);
  v.accessPolicyWindow.num_bytes = bytes < (size_t)ctx->l2_persist_ready ? bytes : (size_t)ctx->l2_persist_ready;
  v.accessPolicyWindow.hitRatio = 1.0f;
  v.accessPolicyWindow.hitProp = bytes ? cudaAccessPropertyPersisting : cudaAccessPropertyNormal;
  v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v);
}
// lazily created side streams + events
int side_streams(stark_ctx *ctx, int n) {
  if (!ctx->fork_ev) CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
  while (ctx->n_side < n) {
    CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->side[ctx->n_side], cudaStreamNonBlocking));
    CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->side_done[ctx->n_side], cudaEventDisableTiming));
    ctx->n_side++;
  }
  return STARK_OK;
}
void ntt_destroy(stark_ctx *ctx) {
  if (ctx->prio_stream) cudaStreamDestroy(ctx->prio_stream);
  if (ctx->prio_ev) cudaEventDestroy(ctx->prio_ev);
  for (int i = 0; i < ctx->n_side; i++) cudaStreamDestroy(ctx->side[i]), cudaEventDestroy(ctx->side_done[i]);
  if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
  cudaFree(ctx->root_lo), cudaFree(ctx->root_hi), cudaFree(ctx->tw_sub[0]), cudaFree(ctx->tw_sub[1]);
  cudaFree(ctx->tw_sh[0]), cudaFree(ctx->tw_sh[1]);
  cudaFree(ctx->tw_in_sh[0]), cudaFree(ctx->tw_in_sh[1]);
  cudaFree(ctx->otw_sh[0]), cudaFree(ctx->otw_sh[1]), cudaFree(ctx->row_sh[0]), cudaFree(ctx->row_sh[1]);
  cudaFree(ctx->row12_sh[0]), cudaFree(ctx->row12_sh[1]);
  for (int i = 0; i < 8; i++) {
    cudaFree(ctx->geo[i].lo);
    if (ctx->geo[i].ready) cudaEventDestroy(ctx->geo[i].ready);
  }
}

// small LRU of geometric tables keyed by (g, c); max_index = largest exponent that will be looked up
int geo_tables(stark_ctx *ctx, u32 g, u32 c, u64 max_index, GeoTables *out) {
  const u32 hi_len = (u32)(max_index >> 12) + 1;
  int victim = 0;
  for (int i = 0; i < 8; i++) {
    GeoCacheEntry &e = ctx->geo[i];
    if (e.lo && e.g == g && e.c == c && e.hi_len >= hi_len) {
      e.stamp = ++ctx->geo_stamp;
      out->lo = e.lo, out->hi = e.hi;
      // the tables may have been filled on another stream of this context (column pipeline, fri.cu)
      if (e.ready) CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, e.ready, 0));
      return STARK_OK;
    }
    if (ctx->geo[i].stamp < ctx->geo[victim].stamp) victim = i;
  }
  GeoCacheEntry &e = ctx->geo[victim];
  if (e.lo) {
    // the old tables may still be in use by queued kernels (on any stream of the context): order the free after them
    CU_TRY(ctx, cudaDeviceSynchronize());
    cudaFree(e.lo);
    e.lo = nullptr;
  }
  const u32 alloc_hi = hi_len < 2048 ? 2048 : hi_len;
  CU_TRY(ctx, cudaMalloc(&e.lo, (size_t)(4096 + alloc_hi) * 4));
  e.hi = e.lo + 4096;
  e.g = g, e.c = c, e.hi_len = alloc_hi, e.stamp = ++ctx->geo_stamp;
  LAUNCH(ctx, "geo_tables", 0, k_geo_tables<<<(alloc_hi + 4095 + 255) / 256, 256, 0, ctx->stream>>>(
                                   e.lo, e.hi, ff::to_mont(g), ff::to_mont(c), alloc_hi));
  if (!e.ready) CU_TRY(ctx, cudaEventCreateWithFlags(&e.ready, cudaEventDisableTiming));
  CU_TRY(ctx, cudaEventRecord(e.ready, ctx->stream));
  out->lo = e.lo, out->hi = e.hi;
  return STARK_OK;
}

// ------------------------------------------------------------------------------------------ kernels

// N <= 4: one thread per transform, direct DFT  X[k] = sum_j x[j] w^(jk)
struct TinyArgs {
  const u32 *in;
  u32 *out;
  int log_n;
  u32 batch;
  u64 in_batch, out_batch, n_valid;
  u32 w_m;  // w_N^(+-1), Montgomery
  int pre_mode, post_mode;
  GeoTables pre_geo, post_geo;
  u32 post_const;
};
__global__ void k_ntt_tiny(TinyArgs A) {
  u32 b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= A.batch) return;
  const int n = 1 << A.log_n;
  u32 x[4], wp[4];
  wp[0] = ff::R1;
  for (int j = 1; j < n; j++) wp[j] = ff::canon(ff::mont_mul(wp[j - 1], A.w_m));
  for (int j = 0; j < n; j++) {
    u32 v = (u64)j < A.n_valid ? A.in[b * A.in_batch + j] : 0u;
    if (A.pre_mode == SCALE_GEO) v = ff::canon(ff::mont_mul(v, geo_pow(A.pre_geo, j)));
    x[j] = v;
  }
  for (int k = 0; k < n; k++) {
    u32 acc = 0;
    for (int j = 0; j < n; j++) acc = ff::add(acc, ff::canon(ff::mont_mul(x[j], wp[(j * k) & (n - 1)])));
    if (A.post_mode == SCALE_CONST)
      acc = ff::canon(ff::mont_mul(acc, A.post_const));
    else if (A.post_mode == SCALE_GEO)
      acc = ff::canon(ff::mont_mul(acc, geo_pow(A.post_geo, k)));
    A.out[b * A.out_batch + k] = acc;
  }
}

template <int LOGR>
__device__ __forceinline__ void first_round(const PassArgs &A, u32 tid, u32 nt, u32 b, bool only, u32 *smem, u32 *regs) {
  round_load_compute<LOGR, true>(tid, nt, b, A, 0, smem, regs);
  if (only)
    round_store<LOGR, true>(tid, nt, b, A, 0, smem, regs);
  else
    round_store<LOGR, false>(tid, nt, b, A, 0, smem, regs);
}

// 8 <= N <= 4096: one CTA of N/8 threads per transform (ntt_core.cuh)
__global__ void __launch_bounds__(512) k_ntt_single(const __grid_constant__ PassArgs A) {
  __shared__ u32 smem[4096];
  const u32 tid = threadIdx.x, nt = blockDim.x, b = blockIdx.x;
  int logr[4];
  const int nr = plan_rounds(A.logL, logr);
  u32 regs[8];
  switch (logr[0]) {
    case 1: first_round<1>(A, tid, nt, b, nr == 1, smem, regs); break;
    case 2: first_round<2>(A, tid, nt, b, nr == 1, smem, regs); break;
    default: first_round<3>(A, tid, nt, b, nr == 1, smem, regs); break;
  }
  int logS = logr[0];
  for (int r = 1; r < nr; r++) {
    const bool last = r == nr - 1;
    __syncthreads();
    round_load_compute<3, false>(tid, nt, b, A, logS, smem, regs);
    __syncthreads();
    if (last)
      round_store<3, true>(tid, nt, b, A, logS, smem, regs);
    else
      round_store<3, false>(tid, nt, b, A, logS, smem, regs);
    logS += 3;
  }
}

// N >= 2^13: one multi-pass Stockham pass (ntt_pass.cuh).  TL = 12: 4096-element tile, 128 threads, 16 KB of static shared
// memory, 8 CTAs per SM.  TL = 14 (the two-pass plans, STARK_NTT_BIG): 16384-element tile, 512 threads, 64 KB of dynamic
// shared memory, 2 CTAs per SM, radix 2^10 / 2^11 in four rounds.
template <int LOGR, int KIND, int MODE, int TL = ntt2::TILE_LOG>
__global__ void __launch_bounds__(1 << (TL - 5), TL == ntt2::TILE_LOG ? 8 : 2) k_ntt2_pass(const __grid_constant__ ntt2::PassParams A) {
  using namespace ntt2;
  typedef Plan<LOGR> PL;
  extern __shared__ __align__(16) unsigned char dyn_tile[];
  __shared__ __align__(16) q4 static_tile[TL == TILE_LOG ? (1 << (TILE_LOG - 2)) : 1];
  __shared__ wpair otw[KIND == MIDDLE ? (1 << LOGR) : 1];
  q4 *tile = TL == TILE_LOG ? static_tile : reinterpret_cast<q4 *>(dyn_tile);
  pdl_entry();
  const u32 tid = threadIdx.x;
  const TileCtx T = tile_ctx<LOGR, TL>(A, blockIdx.x);
  if (KIND == MIDDLE) fill_outer_table<LOGR, TL>(tid, A, T, otw);
  u32 regs[32];
  round_compute<LOGR, KIND, 0, MODE, TL>(tid, A, T, tile, otw, regs);
  round_store<LOGR, KIND, 0, TL>(tid, A, T, tile, regs);
  __syncthreads();
  if constexpr (PL::NR >= 3) {
    round_compute<LOGR, KIND, 1, MODE, TL>(tid, A, T, tile, otw, regs);
    __syncthreads();
    round_store<LOGR, KIND, 1, TL>(tid, A, T, tile, regs);
    __syncthreads();
  }
  if constexpr (PL::NR == 4) {
    round_compute<LOGR, KIND, 2, MODE, TL>(tid, A, T, tile, otw, regs);
    __syncthreads();
    round_store<LOGR, KIND, 2, TL>(tid, A, T, tile, regs);
    __syncthreads();
  }
  round_compute<LOGR, KIND, PL::NR - 1, MODE, TL>(tid, A, T, tile, otw, regs);
  round_store<LOGR, KIND, PL::NR - 1, TL>(tid, A, T, tile, regs);
}
// Batches of 2^12-point transforms: ONE 128-thread CTA per transform on the 4096-element-tile machinery of ntt_pass.cuh
// instead of the 512-thread, 8-elements-per-thread k_ntt_single (1024 x 2^12 ran at 20 % of the HBM peak: one wave and a
// half of fat CTAs, little instruction-level parallelism).  The transform is the two-pass plan {2^7, 2^5} with both passes
// in one launch: the radix-2^7 FIRST pass (four-step twiddles w_4096^(u k)) leaves its transposed result Y[128 u + k] in a
// second 16 KB shared buffer, the radix-2^5 LAST pass reads it from there and stores to HBM with the fused post-scale.
// 32 KB of shared memory, 7 CTAs per SM, 32 elements per thread, 128-bit HBM accesses on both ends.
template <int FM, int LM>
__global__ void __launch_bounds__(ntt2::NT, 7) k_ntt_small12(const __grid_constant__ ntt2::PassParams A,
                                                            const __grid_constant__ ntt2::PassParams B) {
  using namespace ntt2;
  __shared__ __align__(16) q4 tile[1 << (TILE_LOG - 2)];
  __shared__ __align__(16) u32 ybuf[1 << TILE_LOG];
  pdl_entry();
  const u32 tid = threadIdx.x, b = blockIdx.x;
  const wpair *otw = nullptr;
  u32 regs[32];
  TileCtx T1;
  T1.in = A.in + (u64)b * A.in_batch, T1.out = ybuf, T1.col0 = 0, T1.q0 = 0, T1.p = 0;
  round_compute<7, FIRST, 0, FM>(tid, A, T1, tile, otw, regs);
  round_store<7, FIRST, 0>(tid, A, T1, tile, regs);
  __syncthreads();
  round_compute<7, FIRST, 1, FM>(tid, A, T1, tile, otw, regs);
  __syncthreads();
  round_store<7, FIRST, 1>(tid, A, T1, tile, regs);
  __syncthreads();
  round_compute<7, FIRST, 2, FM>(tid, A, T1, tile, otw, regs);
  round_store<7, FIRST, 2>(tid, A, T1, tile, regs);   // -> ybuf
  __syncthreads();
  TileCtx T2;
  T2.in = ybuf, T2.out = B.out + (u64)b * B.out_batch, T2.col0 = 0, T2.q0 = 0, T2.p = 0;
  round_compute<5, LAST, 0, LM>(tid, B, T2, tile, otw, regs);
  round_store<5, LAST, 0>(tid, B, T2, tile, regs);
  __syncthreads();
  round_compute<5, LAST, 1, LM>(tid, B, T2, tile, otw, regs);
  round_store<5, LAST, 1>(tid, B, T2, tile, regs);
}

// launch of a 16384-element-tile pass: 64 KB of dynamic shared memory (opt-in above 48 KB, once per instance)
template <int LOGR, int KIND, int MODE>
static cudaError_t launch_big_pass(cudaStream_t st, u32 grid, const ntt2::PassParams &B) {
  static bool ready = false;
  auto kernel = k_ntt2_pass<LOGR, KIND, MODE, 14>;
  if (!ready) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    if (e != cudaSuccess) return e;
    ready = true;
  }
  return launch_pdl(kernel, dim3(grid), dim3(512), 65536, st, B);
}

// ----------------------------------------------------------------------------------------- launcher

static int resolve_scale(stark_ctx *ctx, const ScaleSpec &s, u64 max_index, int *mode, u32 *c_m, GeoTables *geo) {
  *mode = s.mode;
  *c_m = ff::R1;
  if (s.mode == SCALE_CONST) {
    if (s.c == 1) *mode = SCALE_NONE;
    *c_m = ff::to_mont(s.c);
  } else if (s.mode == SCALE_GEO) {
    if (s.g == 1) {
      *mode = s.c == 1 ? SCALE_NONE : SCALE_CONST;
      *c_m = ff::to_mont(s.c);
    } else {
      ST_TRY(geo_tables(ctx, s.g, s.c, max_index, geo));
    }
  }
  return STARK_OK;
}

int ntt_transform(stark_ctx *ctx, const u32 *in, u32 *out, int log_n, bool inverse, u32 batch, u64 in_batch,
                  u64 out_batch, u64 n_valid, ScaleSpec pre, ScaleSpec post) {
  if (log_n < 0 || log_n > ff::TWO_ADICITY)
    return stark_fail(ctx, STARK_ERR_ARG, "n > 2^23 not supported by this modulus");  // ff.rs:218
  if (batch == 0) return STARK_OK;
  const u64 N = 1ull << log_n;
  if (n_valid > N) n_valid = N;
  const int d = inverse ? 1 : 0;
  int pre_mode, post_mode;
  u32 pre_c, post_c;
  GeoTables pre_geo = {nullptr, nullptr}, post_geo = {nullptr, nullptr};
  ST_TRY(resolve_scale(ctx, pre, N, &pre_mode, &pre_c, &pre_geo));
  ST_TRY(resolve_scale(ctx, post, N, &post_mode, &post_c, &post_geo));
  if (pre_mode == SCALE_CONST) {
    // a constant commutes with the (linear) transform: fold it into the post scale
    if (post_mode == SCALE_NONE) {
      post_mode = SCALE_CONST, post_c = pre_c;
    } else if (post_mode == SCALE_CONST) {
      post_c = ff::canon(ff::mont_mul(post_c, pre_c));
    } else {
      return stark_fail(ctx, STARK_ERR_ARG, "unsupported scale combination");
    }
    pre_mode = SCALE_NONE;
  }
  RootTables roots = {ctx->root_lo, ctx->root_hi};

  if (log_n <= 2) {
    u32 w = ff::pow(ff::GEN, (ff::P - 1) >> log_n);
    if (inverse) w = ff::inv(w);
    TinyArgs A = {in, out, log_n, batch, in_batch, out_batch, n_valid, ff::to_mont(w), pre_mode, post_mode,
                  pre_geo, post_geo, post_c};
    // in == out is fine: each thread reads its whole transform before writing
    LAUNCH(ctx, "ntt_tiny", 8ull * N * batch, k_ntt_tiny<<<(batch + 127) / 128, 128, 0, ctx->stream>>>(A));
    return STARK_OK;
  }

  if (log_n == 12 && batch >= 8 && !ctx->ntt_small_off) {
    // batches of 2^12: the fused two-pass kernel (k_ntt_small12)
    ntt2::PassParams A, B;
    memset(&A, 0, sizeof A);
    A.logN = 12, A.log_tiles = 0, A.roots = roots, A.inverse = d, A.zero = 0;
    for (int k = 0; k < 4; k++) A.w8[k] = ctx->w8_sh[d][k];
    B = A;
    A.in = in, A.in_batch = in_batch, A.n_valid = n_valid, A.logS = 0;
    A.tw_in = ctx->tw_in_sh[d] + ntt2::inner_twiddle_offset(7), A.row_tab = ctx->row12_sh[d];
    A.pre_mode = pre_mode, A.pre_geo = pre_geo;
    A.pre_g1 = A.pre_gj = wpair{1, ff::shoup_of(1)};
    if (pre_mode == SCALE_GEO) {
      const u32 gj = ff::pow(pre.g, N >> 1);   // the FIRST pass's first round is the radix-2 one (7 = 1 + 3 + 3)
      A.pre_g1 = wpair{pre.g, ff::shoup_of(pre.g)}, A.pre_gj = wpair{gj, ff::shoup_of(gj)};
    }
    B.out = out, B.out_batch = out_batch, B.n_valid = N, B.logS = 7;
    B.tw_in = ctx->tw_in_sh[d] + ntt2::inner_twiddle_offset(5);
    const u32 post_plain = ff::from_mont(post_c);
    B.post_mode = post_mode, B.post_const = wpair{post_plain, ff::shoup_of(post_plain)}, B.post_geo = post_geo;
    B.post_g1 = B.post_gk = wpair{1, ff::shoup_of(1)};
    if (post_mode == SCALE_GEO) {
      const u32 gk = ff::pow(post.g, N >> 3);
      B.post_g1 = wpair{post.g, ff::shoup_of(post.g)}, B.post_gk = wpair{gk, ff::shoup_of(gk)};
    }
    const int fm = (n_valid < N ? 1 : 0) | (pre_mode == SCALE_GEO ? 2 : 0);
    const u64 bytes = 4ull * batch * (n_valid + N);
#define SMALL12(F_, L_) LAUNCH_PDL(ctx, "ntt_small12", bytes, (k_ntt_small12<F_, L_>), batch, ntt2::NT, A, B)
#define SMALL12_L(F_)                       \
  if (post_mode == SCALE_NONE) {            \
    SMALL12(F_, 0);                         \
  } else if (post_mode == SCALE_CONST) {    \
    SMALL12(F_, 1);                         \
  } else {                                  \
    SMALL12(F_, 2);                         \
  }
    if (fm == 0) {
      SMALL12_L(0)
    } else if (fm == 1) {
      SMALL12_L(1)
    } else {
      SMALL12_L(3)   // pre-scale, with or without zero padding
    }
#undef SMALL12
#undef SMALL12_L
    return STARK_OK;
  }

  if (log_n <= 12) {
    // one pass, one transform per CTA
    PassArgs A;
    memset(&A, 0, sizeof A);
    for (int k = 0; k < 4; k++) A.w8[k] = ctx->w8_sh[d][k];
    A.pre_mode = pre_mode, A.pre_geo = pre_geo;
    A.post_mode = post_mode, A.post_const = post_c, A.post_geo = post_geo;
    A.in = in, A.out = out, A.logL = log_n;
    A.in_batch = in_batch, A.out_batch = out_batch, A.n_valid = n_valid;
    A.tw = ctx->tw_sh[d] + (1u << log_n);
    LAUNCH(ctx, "ntt_single", 4ull * batch * (n_valid + N), k_ntt_single<<<batch, (u32)(N >> 3), 0, ctx->stream>>>(A));
    return STARK_OK;
  }

  // N >= 2^13: 2 or 3 Stockham passes with radices 2^5 .. 2^8 (ntt_pass.cuh)
  int plan[3];
  int n_pass = ntt2::pass_plan(log_n, plan);
  // experiment (STARK_NTT_BIG=1): TWO passes of radix 2^10 / 2^11 on 16384-element tiles for 2^20 .. 2^22
  const bool big = ctx->ntt_big && log_n >= 20 && log_n <= 22;
  if (big) plan[0] = (log_n + 1) / 2, plan[1] = log_n / 2, plan[2] = 0, n_pass = 2;
  const int tile_log = big ? 14 : ntt2::TILE_LOG;
  // FIRST and MIDDLE passes are out of place, the LAST pass may run in place:
  //   2 passes: in -> out -> out           (in == out: in -> tmp -> out)
  //   3 passes: in -> tmp -> out -> out
  // A batch that fits in L2 goes through each pass in one launch.  A batch much larger than L2 is cut into column
  // GROUPS of ~16 MB whose passes run back to back, so a group's intermediates are still in the 126 MB L2 when the
  // next pass reads them (HBM traffic ~8N per transform instead of 8N per pass); one group alone is less than a wave
  // of CTAs and its launch tails would cost more than the traffic saved (measured: 468 vs 406 us for 16 x 2^22), so
  // ctx->ntt_streams groups are in flight on side streams, forked from / joined to the context's stream by events.
  const u32 batch_total = batch;
  const u32 *in_all = in;
  u32 *out_all = out;
  const u64 in_batch_all = in_batch;
  const bool need_tmp = n_pass == 3 || in == out;
  u32 group = batch;
  int n_streams = 1;
  if (ctx->ntt_streams > 1 && (u64)batch * N * 4 >= (48ull << 20)) {
    const u64 per_group = ((u64)ctx->ntt_group_mb << 20) / (N * 4);
    group = (u32)(per_group ? per_group : 1);
    n_streams = ctx->ntt_streams;
    if ((batch + group - 1) / group < 2u * n_streams) group = batch, n_streams = 1;   // too few groups to pipeline
  }
  u32 *tmp_all = nullptr;
  if (need_tmp) ST_TRY(dev_alloc(ctx, (void **)&tmp_all, (size_t)std::min(batch_total, group) * N * 4 * n_streams));
  cudaStream_t main_stream = ctx->stream;
  // Every exit of this function -- the launch macros return on a failed launch -- must leave the context on its own
  // stream, join the side streams (a later call on this context must not race with work still queued there), drop
  // the L2 window and release the scratch: a scope guard does all of it.
  struct PassGuard {
    stark_ctx *c;
    cudaStream_t main;
    int n_streams;
    u32 *tmp;
    ~PassGuard() {
      c->stream = main;
      if (n_streams > 1) {
        if (c->ntt_l2_persist)
          for (int i = 0; i < n_streams; i++) l2_window(c, c->side[i], nullptr, 0);
        for (int i = 0; i < n_streams; i++) {
          cudaEventRecord(c->side_done[i], c->side[i]);
          cudaStreamWaitEvent(main, c->side_done[i], 0);
        }
      }
      dev_free(c, tmp);
    }
  } guard{ctx, main_stream, 1, tmp_all};
  if (n_streams > 1) {
    ST_TRY(side_streams(ctx, n_streams));
    guard.n_streams = n_streams;
    cudaEventRecord(ctx->fork_ev, main_stream);
    for (int i = 0; i < n_streams; i++) cudaStreamWaitEvent(ctx->side[i], ctx->fork_ev, 0);
  }
  ntt2::PassParams B;
  memset(&B, 0, sizeof B);
  B.logN = log_n, B.log_tiles = log_n - tile_log;
  B.roots = roots, B.inverse = d, B.zero = 0;
  for (int k = 0; k < 4; k++) B.w8[k] = ctx->w8_sh[d][k];
  const u32 post_plain = ff::from_mont(post_c);
  const wpair post_sh = {post_plain, ff::shoup_of(post_plain)};
  // geometric pre-scale walks: g and g^(N / radix of the FIRST pass's first round)
  wpair pre_g1 = {1, ff::shoup_of(1)}, pre_gj = pre_g1;
  if (pre_mode == SCALE_GEO && log_n >= 13) {
    const int lr0 = plan[0] % 3 ? plan[0] % 3 : 3;
    const u32 gj = ff::pow(pre.g, N >> lr0);
    pre_g1 = wpair{pre.g, ff::shoup_of(pre.g)}, pre_gj = wpair{gj, ff::shoup_of(gj)};
  }
  // geometric post-scale walks (ntt_pass.cuh): g and g^(N/8), the distance between the outputs of the last radix-8 round
  wpair post_g1 = {1, ff::shoup_of(1)}, post_gk = post_g1;
  if (post_mode == SCALE_GEO) {
    const u32 gk = ff::pow(post.g, N >> 3);
    post_g1 = wpair{post.g, ff::shoup_of(post.g)}, post_gk = wpair{gk, ff::shoup_of(gk)};
  }
  int rc = STARK_OK;
  u32 gi = 0;
  for (u32 b0 = 0; b0 < batch_total && rc == STARK_OK; b0 += group, gi++) {
  batch = std::min(group, batch_total - b0);
  u32 *tmp = tmp_all;
  if (n_streams > 1) {
    ctx->stream = ctx->side[gi % n_streams];   // the LAUNCH macros launch on ctx->stream
    if (tmp_all) tmp = tmp_all + (size_t)(gi % n_streams) * group * N;
  }
  in = in_all + (u64)b0 * in_batch_all, out = out_all + (u64)b0 * out_batch, in_batch = in_batch_all;
  const u32 grid = batch << B.log_tiles;
  const u32 *src = in;
  u64 src_batch = in_batch;
  int logS = 0;
  for (int i = 0; i < n_pass && rc == STARK_OK; i++) {
    const int r = plan[i];
    const int kind = i == 0 ? ntt2::FIRST : (i == n_pass - 1 ? ntt2::LAST : ntt2::MIDDLE);
    u32 *dst;
    u64 dst_batch;
    if (kind == ntt2::LAST || (kind == ntt2::MIDDLE) || (n_pass == 2 && !need_tmp))
      dst = out, dst_batch = out_batch;
    else
      dst = tmp, dst_batch = N;
    B.in = src, B.out = dst, B.in_batch = src_batch, B.out_batch = dst_batch;
    B.n_valid = kind == ntt2::FIRST ? n_valid : N;
    B.logS = logS;
    B.tw_in = ctx->tw_in_sh[d] + ntt2::inner_twiddle_offset(r);
    B.otw_tab = ctx->otw_sh[d], B.otw_shift = 16 - (log_n - logS);
    B.row_tab = ctx->row_sh[d] + (log_n - 13) * 2048;
    B.pre_mode = kind == ntt2::FIRST ? pre_mode : (int)SCALE_NONE, B.pre_geo = pre_geo;
    B.post_mode = kind == ntt2::LAST ? post_mode : (int)SCALE_NONE, B.post_const = post_sh, B.post_geo = post_geo, B.post_g1 = post_g1, B.post_gk = post_gk;
    B.pre_g1 = pre_g1, B.pre_gj = pre_gj;
    if (n_streams > 1 && ctx->ntt_l2_persist) l2_window(ctx, ctx->stream, dst, (size_t)batch * N * 4);
    const u64 bytes = kind == ntt2::FIRST ? 4ull * batch * (n_valid + N) : 8ull * batch * N;
    const char *tag = kind == ntt2::FIRST ? "ntt_pass1" : (kind == ntt2::LAST ? "ntt_pass_last" : "ntt_pass_mid");
    // compile-time specialisation of the per-element options (see round_compute)
    const int mode = kind == ntt2::FIRST ? ((B.n_valid < N ? 1 : 0) | (B.pre_mode == SCALE_GEO ? 2 : 0))
                                         : (kind == ntt2::LAST ? B.post_mode : 0);
    if (big) {
      cudaError_t le = cudaErrorInvalidValue;
      if (ctx->prof_on) prof_begin(ctx, tag, bytes);
      const int kmode = (kind == ntt2::FIRST && mode == 2) ? 3 : mode;   // FIRST: pre-scale with or without padding
#define BIG_CASE(R_, K_, M_) if (r == R_ && kind == K_ && kmode == M_) le = launch_big_pass<R_, K_, M_>(ctx->stream, grid, B); else
      BIG_CASE(10, ntt2::FIRST, 0) BIG_CASE(10, ntt2::FIRST, 1) BIG_CASE(10, ntt2::FIRST, 3) BIG_CASE(11, ntt2::FIRST, 0)
      BIG_CASE(11, ntt2::FIRST, 1) BIG_CASE(11, ntt2::FIRST, 3) BIG_CASE(10, ntt2::LAST, 0) BIG_CASE(10, ntt2::LAST, 1)
      BIG_CASE(10, ntt2::LAST, 2) BIG_CASE(11, ntt2::LAST, 0) BIG_CASE(11, ntt2::LAST, 1) BIG_CASE(11, ntt2::LAST, 2)
      le = cudaErrorInvalidValue;
#undef BIG_CASE
      if (ctx->prof_on) prof_end(ctx);
      ctx->launches++;
      if (le != cudaSuccess) rc = stark_fail(ctx, STARK_ERR_CUDA, "NTT pass launch failed: %s", cudaGetErrorString(le));
      src = dst, src_batch = dst_batch;
      logS += r;
      continue;
    }
#define NTT2_LAUNCH(R_, K_, M_) LAUNCH_PDL(ctx, tag, bytes, (k_ntt2_pass<R_, K_, M_>), grid, ntt2::NT, B)
#define NTT2_CASE(R_, K_)                                                                                        \
  if (r == R_ && kind == K_) {                                                                                   \
    if (K_ == ntt2::MIDDLE || mode == 0) {                                                                       \
      NTT2_LAUNCH(R_, K_, 0);                                                                                    \
    } else if (mode == 1) {                                                                                      \
      NTT2_LAUNCH(R_, K_, 1);   /* FIRST: zero padding only; LAST: constant post-scale */                        \
    } else if (K_ == ntt2::FIRST) {                                                                              \
      NTT2_LAUNCH(R_, K_, 3);   /* FIRST: pre-scale (with or without zero padding) */                            \
    } else {                                                                                                     \
      NTT2_LAUNCH(R_, K_, 2);   /* LAST: geometric post-scale */                                                 \
    }                                                                                                            \
  } else
    NTT2_CASE(6, ntt2::FIRST) NTT2_CASE(7, ntt2::FIRST) NTT2_CASE(8, ntt2::FIRST)
    NTT2_CASE(6, ntt2::MIDDLE) NTT2_CASE(7, ntt2::MIDDLE) NTT2_CASE(8, ntt2::MIDDLE)
    NTT2_CASE(5, ntt2::LAST) NTT2_CASE(6, ntt2::LAST) NTT2_CASE(7, ntt2::LAST) NTT2_CASE(8, ntt2::LAST)
    rc = stark_fail(ctx, STARK_ERR_ARG, "no NTT pass kernel for radix 2^%d", r);
#undef NTT2_LAUNCH
#undef NTT2_CASE
    src = dst, src_batch = dst_batch;
    logS += r;
  }
  }
  return rc;
}
