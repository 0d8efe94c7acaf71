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
#include "common.cuh"

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
  const u32 w23 = ff::pow(ff::GEN, (ff::P - 1) >> 23);  // ff.rs:215-223
  k_root_tables<<<16, 256, 0, ctx->stream>>>(ctx->root_lo, ctx->root_hi, ff::to_mont(w23));
  KERNEL_CHECK(ctx);
  RootTables T = {ctx->root_lo, ctx->root_hi};
  for (int d = 0; d < 2; d++) {
    k_sub_tables<<<32, 256, 0, ctx->stream>>>(ctx->tw_sub[d], T, d);
    KERNEL_CHECK(ctx);
    u32 w8 = ff::pow(ff::GEN, (ff::P - 1) >> 3);
    if (d) w8 = ff::inv(w8);
    ctx->w8[d][0] = ff::R1;
    for (int k = 1; k < 4; k++) ctx->w8[d][k] = ff::to_mont(ff::pow(w8, k));
  }
  for (int i = 0; i < 8; i++) ctx->geo[i] = GeoCacheEntry{0, 0, nullptr, nullptr, 0, 0};
  ctx->geo_stamp = 0;
  return STARK_OK;
}
void ntt_destroy(stark_ctx *ctx) {
  cudaFree(ctx->root_lo), cudaFree(ctx->root_hi), cudaFree(ctx->tw_sub[0]), cudaFree(ctx->tw_sub[1]);
  for (int i = 0; i < 8; i++) cudaFree(ctx->geo[i].lo);
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
      return STARK_OK;
    }
    if (ctx->geo[i].stamp < ctx->geo[victim].stamp) victim = i;
  }
  GeoCacheEntry &e = ctx->geo[victim];
  if (e.lo) {
    // the old tables may still be in use by queued kernels: order the free after them
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(e.lo);
    e.lo = nullptr;
  }
  const u32 alloc_hi = hi_len < 2048 ? 2048 : hi_len;
  CU_TRY(ctx, cudaMalloc(&e.lo, (size_t)(4096 + alloc_hi) * 4));
  e.hi = e.lo + 4096;
  e.g = g, e.c = c, e.hi_len = alloc_hi, e.stamp = ++ctx->geo_stamp;
  LAUNCH(ctx, "geo_tables", 0, k_geo_tables<<<(alloc_hi + 4095 + 255) / 256, 256, 0, ctx->stream>>>(
                                   e.lo, e.hi, ff::to_mont(g), ff::to_mont(c), alloc_hi));
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

template <int LOGR, int V, bool ROWOUT>
__device__ __forceinline__ void first_round(const PassArgs &A, u32 tid, u32 nt, u32 tile, bool only,
                                            typename Slot<V>::type *smem, u32 *regs) {
  round_load_compute<LOGR, V, true>(tid, nt, tile, A, 0, smem, only && ROWOUT, regs);
  if (only)
    round_store<LOGR, V, true, ROWOUT>(tid, nt, tile, A, 0, smem, regs);
  else
    round_store<LOGR, V, false, ROWOUT>(tid, nt, tile, A, 0, smem, regs);
}

template <int V, bool ROWOUT>
__global__ void __launch_bounds__(1024, 1) k_ntt_pass(const __grid_constant__ PassArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef typename Slot<V>::type slot_t;
  slot_t *smem = reinterpret_cast<slot_t *>(smem_raw);
  const u32 tid = threadIdx.x, nt = blockDim.x, tile = blockIdx.x;
  int logr[4];
  const int nr = plan_rounds(A.logL, logr);
  u32 regs[32];
  switch (logr[0]) {
    case 1: first_round<1, V, ROWOUT>(A, tid, nt, tile, nr == 1, smem, regs); break;
    case 2: first_round<2, V, ROWOUT>(A, tid, nt, tile, nr == 1, smem, regs); break;
    default: first_round<3, V, ROWOUT>(A, tid, nt, tile, nr == 1, smem, regs); break;
  }
  int logS = logr[0];
  for (int r = 1; r < nr; r++) {
    const bool last = r == nr - 1;
    __syncthreads();
    round_load_compute<3, V, false>(tid, nt, tile, A, logS, smem, last && ROWOUT, regs);
    __syncthreads();
    if (last)
      round_store<3, V, true, ROWOUT>(tid, nt, tile, A, logS, smem, regs);
    else
      round_store<3, V, false, ROWOUT>(tid, nt, tile, A, logS, smem, regs);
    logS += 3;
  }
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

template <int V, bool ROWOUT>
static int launch_pass(stark_ctx *ctx, const PassArgs &A, u32 tiles, const char *tag, u64 bytes) {
  const u32 nt = (1u << (A.logL - 3)) << A.logC4;
  const size_t smem = ((size_t)V * 4) << (A.logL + A.logC4);
  static bool configured = false;  // per instantiation
  if (!configured) {
    CU_TRY(ctx, cudaFuncSetAttribute(k_ntt_pass<V, ROWOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = true;
  }
  LAUNCH(ctx, tag, bytes, k_ntt_pass<V, ROWOUT><<<tiles, nt, smem, ctx->stream>>>(A));
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

  PassArgs A;
  memset(&A, 0, sizeof A);
  A.roots = roots;
  A.shiftN = 23 - log_n;
  A.inverse = d;
  for (int k = 0; k < 4; k++) A.w8[k] = ctx->w8[d][k];
  A.pre_mode = pre_mode, A.pre_geo = pre_geo;
  A.post_mode = post_mode, A.post_const = post_c, A.post_geo = post_geo;

  if (log_n <= 12) {
    // one pass, one transform per CTA, scalar columns (V = 1)
    A.in = in, A.out = out;
    A.logL = log_n, A.logC4 = 0;
    A.in_batch = in_batch, A.in_stride = 1, A.n_valid = n_valid;
    A.out_batch = out_batch, A.out_stride = 1;
    A.tiles_per_batch = 1;
    A.tw = ctx->tw_sub[d] + (1u << log_n);
    return launch_pass<1, false>(ctx, A, batch, "ntt_single", 4ull * batch * (n_valid + N));
  }

  // two passes.  pass 1 cannot run in place (it transposes), so in == out goes through a scratch copy.
  const int log_n1 = log_n / 2, log_n2 = log_n - log_n1;
  const u64 N1 = 1ull << log_n1, N2 = 1ull << log_n2;
  u32 *tmp = nullptr;
  const u32 *src = in;
  if (in == out) {
    ST_TRY(dev_alloc(ctx, (void **)&tmp, (size_t)batch * N * 4));
    CU_TRY(ctx, cudaMemcpy2DAsync(tmp, N * 4, in, in_batch * 4, N * 4, batch, cudaMemcpyDeviceToDevice, ctx->stream));
    src = tmp;
    in_batch = N;
  }
  {  // pass 1: L = N1, columns n2 (stride N2), row-mode (transposed) store + four-step twiddle
    PassArgs B = A;
    B.in = src, B.out = out;
    B.logL = log_n1;
    int logC = log_n2 < (15 - log_n1) ? log_n2 : (15 - log_n1);  // tile <= 32K elements
    B.logC4 = logC - 2;
    B.in_batch = in_batch, B.in_stride = N2, B.n_valid = n_valid;
    B.out_batch = out_batch, B.out_stride = 0;
    B.tiles_per_batch = (int)(N2 >> logC);
    B.tw = ctx->tw_sub[d] + (1u << log_n1);
    B.post_mode = SCALE_NONE;
    ST_TRY((launch_pass<4, true>(ctx, B, batch * (u32)B.tiles_per_batch, "ntt_pass1", 4ull * batch * (n_valid + N))));
  }
  {  // pass 2: L = N2, columns k1 (stride N1), in place on out
    PassArgs B = A;
    B.in = out, B.out = out;
    B.logL = log_n2;
    int logC = log_n1 < (15 - log_n2) ? log_n1 : (15 - log_n2);
    B.logC4 = logC - 2;
    B.in_batch = out_batch, B.in_stride = N1, B.n_valid = N;
    B.out_batch = out_batch, B.out_stride = N1;
    B.tiles_per_batch = (int)(N1 >> logC);
    B.tw = ctx->tw_sub[d] + (1u << log_n2);
    B.pre_mode = SCALE_NONE;
    ST_TRY((launch_pass<4, false>(ctx, B, batch * (u32)B.tiles_per_batch, "ntt_pass2", 8ull * batch * N)));
  }
  dev_free(ctx, tmp);
  return STARK_OK;
}
