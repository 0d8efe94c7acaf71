// poly.cu -- univariate Polynomial operations and the trace low-degree extension on top of the NTT.
//
// Reference: src/univariate/{mul,eval,interpolate,mod}.rs.  The reference multiplies by schoolbook
// (mul.rs:6-29), evaluates by Horner per point (eval.rs:6-21) and interpolates by O(n^3) Lagrange
// (interpolate.rs:6-44); the trace LDE is the composition interpolate_domain + eval_domain on a coset
// (pattern fri.rs:575-578).  All have unique mathematical results, so the NTT pipelines below
// are bit-exact as long as outputs are canonical and in natural order; the vector-LENGTH rules of the
// reference are reproduced explicitly.
#include "common.cuh"
#include "merkle.h"

using ntt::GeoTables;

// ------------------------------------------------------------------------------------------- kernels

// a[i] = a[i] * b[i] / R  (Montgomery product of two canonical values; the R is repaid by the iNTT post-scale)
__global__ void k_pointwise(u32 *__restrict__ a, const u32 *__restrict__ b, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    a[i] = ff::canon(ff::mont_mul(a[i], b[i]));
}

// out[i] = c[i] * g^i   (Polynomial::scale, mod.rs:99-113)
__global__ void k_scale_geo(const u32 *__restrict__ c, u32 *__restrict__ out, size_t n, GeoTables G) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = ff::canon(ff::mont_mul(c[i], ntt::geo_pow(G, i)));
}

// Polynomial::eval_domain on an arbitrary domain (eval.rs:6-21): one thread per point, Horner from the top
// coefficient; coefficients staged through shared memory in tiles of 1024 (read once per CTA).
__global__ void __launch_bounds__(256) k_eval_domain(const u32 *__restrict__ coeffs, size_t nc,
                                                     const u32 *__restrict__ dom, size_t m, u32 *__restrict__ out) {
  __shared__ u32 tile[1024];
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const u32 x_m = i < m ? ff::to_mont(dom[i]) : 0u;
  u32 acc = 0;  // lazily in [0, 2p)
  for (size_t hi = nc; hi > 0;) {
    const size_t lo = hi >= 1024 ? hi - 1024 : 0, len = hi - lo;
    __syncthreads();
    for (size_t k = threadIdx.x; k < len; k += blockDim.x) tile[k] = coeffs[lo + k];
    __syncthreads();
    for (size_t k = len; k-- > 0;) acc = ff::red2p(ff::mont_mul(acc, x_m) + tile[k]);
    hi = lo;
  }
  if (i < m) out[i] = ff::canon(acc);
}

// barycentric weights for Lagrange interpolation on an arbitrary domain:
//   c[i] = y[i] / prod_{j != i} (x[i] - x[j])         (the `denom` products of interpolate.rs:33-39)
// sets *flag if two points coincide ("no inverse", mod.rs:613-625)
__global__ void __launch_bounds__(256) k_lagrange_weights(const u32 *__restrict__ x, const u32 *__restrict__ y, size_t n,
                                                          u32 *__restrict__ c, u32 *flag) {
  __shared__ u32 tile[1024];
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const u32 xi = i < n ? x[i] : 0u;
  u32 prod = ff::R1;  // Montgomery form
  bool dup = false;
  for (size_t lo = 0; lo < n; lo += 1024) {
    const size_t len = n - lo < 1024 ? n - lo : 1024;
    __syncthreads();
    for (size_t k = threadIdx.x; k < len; k += blockDim.x) tile[k] = x[lo + k];
    __syncthreads();
    if (i < n)
      for (size_t k = 0; k < len; k++) {
        if (lo + k == i) continue;
        const u32 d = ff::sub(xi, tile[k]);
        if (d == 0) dup = true;
        prod = ff::canon(ff::mont_mul(prod, ff::to_mont(d)));
      }
  }
  if (i >= n) return;
  if (dup) {
    atomicOr(flag, 2u);
    c[i] = 0;
    return;
  }
  const u32 inv = ff::mont_pow(prod, (u64)ff::P - 2);
  c[i] = ff::canon(ff::mont_mul(inv, y[i]));  // canonical y/denominator
}

// master polynomial M(X) = prod_j (X - x[j]) (Polynomial::zerofier, mod.rs:77-96): one CTA, coefficients in
// shared memory, one sequential step per root, all coefficients updated in parallel.  n <= ZF_MAX.
#define ZF_MAX 8192
__global__ void __launch_bounds__(1024) k_zerofier(const u32 *__restrict__ x, u32 n, u32 *__restrict__ out) {
  extern __shared__ u32 zf[];  // 2 * (n + 1)
  u32 *cur = zf, *nxt = zf + (n + 1);
  for (u32 k = threadIdx.x; k <= n; k += blockDim.x) cur[k] = k == 0 ? 1u : 0u;
  __syncthreads();
  for (u32 j = 0; j < n; j++) {
    const u32 xj_m = ff::to_mont(x[j]);
    // degree j -> j + 1 :  new[k] = cur[k-1] - x_j * cur[k]
    for (u32 k = threadIdx.x; k <= j + 1; k += blockDim.x) {
      const u32 a = k > 0 ? cur[k - 1] : 0u;
      const u32 b = k <= j ? ff::canon(ff::mont_mul(cur[k], xj_m)) : 0u;
      nxt[k] = ff::sub(a, b);
    }
    __syncthreads();
    u32 *t = cur;
    cur = nxt, nxt = t;
  }
  for (u32 k = threadIdx.x; k <= n; k += blockDim.x) out[k] = cur[k];
}

// f = sum_i c[i] * M(X) / (X - x[i]).  Thread i runs the synthetic division q_{k-1} = m_k + x_i q_k downwards;
// at every step the CTA reduces c[i]*q_k over its threads and adds the partial sum to out[k-1].
// Partial sums of different CTAs go to separate rows of `partial` (no atomics; summed by k_reduce_rows).
__global__ void __launch_bounds__(256) k_lagrange_accumulate(const u32 *__restrict__ x, const u32 *__restrict__ c,
                                                             const u32 *__restrict__ mpoly, u32 n,
                                                             u32 *__restrict__ partial) {
  __shared__ u32 red[8];
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  const u32 xi_m = live ? ff::to_mont(x[i]) : 0u, ci_m = live ? ff::to_mont(c[i]) : 0u;
  u32 q = 0;  // canonical
  u32 *row = partial + (size_t)blockIdx.x * n;
  for (u32 k = n; k >= 1; k--) {
    // q_{k-1} = m_k + x_i * q_k   (q_n = 0, so q_{n-1} = m_n = 1)
    q = ff::add(mpoly[k], ff::canon(ff::mont_mul(q, xi_m)));
    u32 term = live ? ff::canon(ff::mont_mul(q, ci_m)) : 0u;
    // CTA sum of `term` (values < p: pairwise canonical adds)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) term = ff::add(term, __shfl_xor_sync(0xffffffffu, term, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = term;
    __syncthreads();
    if (threadIdx.x == 0) {
      u32 s = 0;
      for (u32 w = 0; w < (blockDim.x >> 5); w++) s = ff::add(s, red[w]);
      row[k - 1] = s;
    }
    __syncthreads();
  }
}
__global__ void k_reduce_rows(const u32 *__restrict__ partial, u32 rows, u32 n, u32 *__restrict__ out) {
  const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  u32 s = 0;
  for (u32 r = 0; r < rows; r++) s = ff::add(s, partial[(size_t)r * n + k]);
  out[k] = s;
}

// ---- Polynomial::div helpers (div.rs:6-53 by Newton inversion of the reversed divisor)
// dst[i] = src[hi - i] for i < cnt (i <= hi), zero beyond: the first cnt coefficients of the reversed polynomial
__global__ void k_reverse_copy(const u32 *__restrict__ src, size_t hi, u32 *__restrict__ dst, size_t cnt) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += stride) dst[i] = i <= hi ? src[hi - i] : 0u;
}
// t = 2 - t  (mod x^n)
__global__ void k_two_minus(u32 *__restrict__ t, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    t[i] = i == 0 ? ff::sub(2u, t[0]) : ff::neg(t[i]);
}
// r[i] = a[i] - qb[i] for i < n_low (a beyond na and qb beyond nqb count as zero), r[i] = 0 for n_low <= i < len;
// flag |= 4 when a coefficient at or above n_low of a - qb is non-zero (cannot happen for a correct quotient)
__global__ void k_remainder(const u32 *__restrict__ a, size_t na, const u32 *__restrict__ qb, size_t nqb, size_t n_low,
                            u32 *__restrict__ r, size_t len, u32 *flag) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
    const u32 x = i < na ? a[i] : 0u, y = i < nqb ? qb[i] : 0u;
    const u32 d = ff::sub(x, y);
    if (i < n_low) {
      r[i] = d;
    } else {
      r[i] = 0u;
      if (d) atomicOr(flag, 4u);
    }
  }
}

// ------------------------------------------------------------------------------------- device-level

static u32 grid_for(stark_ctx *ctx, size_t n, u32 block) {
  size_t b = (n + block - 1) / block, cap = (size_t)ctx->sm_count * 16;
  return (u32)(b < cap ? (b ? b : 1) : cap);
}

// values on w_n^i -> values on offset * w_{bn}^i, batched over columns (column-major)
int lde_dev(stark_ctx *ctx, const u32 *cols, u32 n_cols, u32 log_n, u32 log_blowup, u32 offset, u32 *out) {
  if (log_n + log_blowup > (u32)ff::TWO_ADICITY)
    return stark_fail(ctx, STARK_ERR_ARG, "n > 2^23 not supported by this modulus");  // ff.rs:218
  if (n_cols == 0) return STARK_OK;
  const u64 n = 1ull << log_n, N = n << log_blowup;
  u32 *coef = nullptr;
  Scratch sc(ctx);
  ST_TRY(sc.get(&coef, (size_t)n_cols * n * 4));
  // interpolate on the trace domain, and scale coefficient j by n^-1 * offset^j on the way out
  ScaleSpec none = {ntt::SCALE_NONE, 1, 1};
  ScaleSpec post = {ntt::SCALE_GEO, ff::inv((u32)(n % ff::P)), offset};
  int rc = ntt_transform(ctx, cols, coef, (int)log_n, true, n_cols, n, n, n, none, post);
  // evaluate on the big domain; the zero padding is never read
  if (rc == STARK_OK) rc = ntt_transform(ctx, coef, out, (int)(log_n + log_blowup), false, n_cols, n, N, n, none, none);
  return rc;
}

static bool all_zero(const uint64_t *v, size_t n) {
  for (size_t i = 0; i < n; i++)
    if (v[i] != 0) return false;
  return true;
}
static int log2_exact(size_t n) {
  int l = 0;
  while (((size_t)1 << l) < n) l++;
  return l;
}

// out[0 .. keep) = the low `keep` coefficients of a * b (device, canonical); na, nb >= 1
static int poly_mul_dev(stark_ctx *ctx, const u32 *a, size_t na, const u32 *b, size_t nb, u32 *out, size_t keep) {
  const size_t m = na + nb - 1;
  const int lg = log2_exact(m) < 3 ? 3 : log2_exact(m);
  if (lg > ff::TWO_ADICITY) return stark_fail(ctx, STARK_ERR_ARG, "n > 2^23 not supported by this modulus");
  const size_t M = (size_t)1 << lg;
  u32 *fa = nullptr, *fb = nullptr;
  Scratch sc(ctx);
  ST_TRY(sc.get(&fa, M * 4));
  ST_TRY(sc.get(&fb, M * 4));
  ScaleSpec none = {ntt::SCALE_NONE, 1, 1};
  int rc = ntt_transform(ctx, a, fa, lg, false, 1, M, M, na, none, none);
  if (rc == STARK_OK) rc = ntt_transform(ctx, b, fb, lg, false, 1, M, M, nb, none, none);
  if (rc == STARK_OK) {
    k_pointwise<<<grid_for(ctx, M, 256), 256, 0, ctx->stream>>>(fa, fb, M);
    ctx->launches++;
    ScaleSpec post = {ntt::SCALE_CONST, ff::mul(ff::inv((u32)(M % ff::P)), ff::R1), 1};
    rc = ntt_transform(ctx, fa, fb, lg, true, 1, M, M, M, none, post);
  }
  if (rc == STARK_OK) {
    const size_t n = keep < m ? keep : m;
    CU_TRY(ctx, cudaMemcpyAsync(out, fb, n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    if (keep > m) CU_TRY(ctx, cudaMemsetAsync(out + m, 0, (keep - m) * 4, ctx->stream));
  }
  return rc;
}

// Polynomial::div (div.rs:6-53) for deg a = da >= deg b = db >= 0 on device data: q gets k = da - db + 1 coefficients,
// r gets r_len coefficients (zeros above db - 1).  Quotient by Newton inversion of the reversed divisor:
//   rev(q) = rev(a) * rev(b)^-1 mod x^k,   g <- g (2 - rev(b) g) doubling the precision each step.
static int poly_div_dev(stark_ctx *ctx, const u32 *a, size_t na, size_t da, const u32 *b, size_t db, u32 *q, u32 *r,
                        size_t r_len) {
  const size_t k = da - db + 1;
  u32 *ra = nullptr, *rb = nullptr, *g = nullptr, *t = nullptr, *qb = nullptr;
  Scratch sc(ctx);
  ST_TRY(sc.get(&ra, k * 4));
  ST_TRY(sc.get(&rb, k * 4));
  ST_TRY(sc.get(&g, k * 4));
  ST_TRY(sc.get(&t, k * 4));
  ST_TRY(sc.get(&qb, (k + db) * 4));
  k_reverse_copy<<<grid_for(ctx, k, 256), 256, 0, ctx->stream>>>(a, da, ra, k);
  k_reverse_copy<<<grid_for(ctx, k, 256), 256, 0, ctx->stream>>>(b, db, rb, k);
  ctx->launches += 2;
  // g0 = lead(b)^-1 (div.rs:25 divides by the leading coefficient; it is non-zero by definition of the degree)
  u32 lead = 0;
  CU_TRY(ctx, cudaMemcpyAsync(&lead, b + db, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const u32 g0 = ff::inv(lead);
  CU_TRY(ctx, cudaMemcpyAsync(g, &g0, 4, cudaMemcpyHostToDevice, ctx->stream));
  int rc = STARK_OK;
  for (size_t prec = 1; prec < k && rc == STARK_OK; prec *= 2) {
    const size_t np = 2 * prec < k ? 2 * prec : k;
    rc = poly_mul_dev(ctx, rb, np, g, prec, t, np);                         // t = rev(b) g mod x^np
    if (rc != STARK_OK) break;
    k_two_minus<<<grid_for(ctx, np, 256), 256, 0, ctx->stream>>>(t, np);    // t = 2 - t
    ctx->launches++;
    rc = poly_mul_dev(ctx, g, prec, t, np, g, np);                          // g = g t mod x^np
  }
  if (rc == STARK_OK) rc = poly_mul_dev(ctx, ra, k, g, k, t, k);            // rev(q) = rev(a) g mod x^k
  if (rc == STARK_OK) {
    k_reverse_copy<<<grid_for(ctx, k, 256), 256, 0, ctx->stream>>>(t, k - 1, q, k);
    ctx->launches++;
    rc = poly_mul_dev(ctx, q, k, b, db + 1, qb, k + db);                    // q b, da + 1 coefficients
  }
  if (rc == STARK_OK) {
    CU_TRY(ctx, cudaMemsetAsync(ctx->flag, 0, 4, ctx->stream));
    k_remainder<<<grid_for(ctx, r_len, 256), 256, 0, ctx->stream>>>(a, na, qb, k + db, db, r, r_len, ctx->flag);
    ctx->launches++;
  }
  return rc;
}

// ----------------------------------------------------------------------------------------------- C ABI

extern "C" {

// Polynomial::div (div.rs:6-53): (quotient, remainder) with the reference's vector lengths:
//   deg b = -1 -> STARK_ERR_ARG "No division by zero";  deg a < deg b -> q = [], r = a (na coefficients);
//   else q has deg a - deg b + 1 coefficients and r has max(na, deg a + 1 + (nb - 1 - deg b)) (the length the
//   reference's repeated Polynomial::sub leaves behind), zeros above deg b - 1.
int stark_poly_div(stark_ctx *ctx, const uint64_t *a, size_t na, const uint64_t *b, size_t nb, uint64_t *q,
                   size_t *q_len, uint64_t *r, size_t *r_len) {
  if (!ctx || !q_len || !r_len || (na && !a) || (nb && !b)) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  long da = -1, db = -1;
  for (size_t i = 0; i < na; i++)
    if (a[i]) da = (long)i;
  for (size_t i = 0; i < nb; i++)
    if (b[i]) db = (long)i;
  if (db < 0) return stark_fail(ctx, STARK_ERR_ARG, "No division by zero");   // div.rs:7-9
  if (da < db) {                                                               // div.rs:10-18
    *q_len = 0, *r_len = na;
    if (na && !r) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
    for (size_t i = 0; i < na; i++) {
      if (a[i] >= ff::P) return stark_fail(ctx, STARK_ERR_ARG, "non-canonical field element (value >= p) in input");
      r[i] = a[i];
    }
    return STARK_OK;
  }
  const size_t k = (size_t)(da - db) + 1;
  const size_t tz = nb - 1 - (size_t)db;
  const size_t rl = na > (size_t)da + 1 + tz ? na : (size_t)da + 1 + tz;
  if (!q || !r) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  u32 *d_a = nullptr, *d_b = nullptr, *d_q = nullptr, *d_r = nullptr;
  Scratch sc(ctx);
  ST_TRY(sc.get(&d_a, na * 4));
  ST_TRY(sc.get(&d_b, nb * 4));
  ST_TRY(sc.get(&d_q, k * 4));
  ST_TRY(sc.get(&d_r, rl * 4));
  int rc = upload_u64(ctx, a, na, d_a);
  if (rc == STARK_OK) rc = upload_u64(ctx, b, nb, d_b);
  if (rc == STARK_OK) rc = poly_div_dev(ctx, d_a, na, (size_t)da, d_b, (size_t)db, d_q, d_r, rl);
  if (rc == STARK_OK) rc = download_u64(ctx, d_q, k, q);
  if (rc == STARK_OK) rc = download_u64(ctx, d_r, rl, r);
  if (rc == STARK_OK) {
    CU_TRY(ctx, cudaMemcpyAsync(ctx->h_flag, ctx->flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (*ctx->h_flag & 4u) rc = stark_fail(ctx, STARK_ERR_CUDA, "internal error: quotient check failed");
  }
  if (rc == STARK_OK) *q_len = k, *r_len = rl;
  return rc;
}

int stark_poly_mul(stark_ctx *ctx, const uint64_t *a, size_t na, const uint64_t *b, size_t nb, uint64_t *out,
                   size_t *out_len) {
  if (!ctx || !out_len || (na && !a) || (nb && !b)) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  // mul.rs:7-12: a zero polynomial (empty or all-zero coefficients) on either side gives []
  if (na == 0 || nb == 0 || all_zero(a, na) || all_zero(b, nb)) {
    *out_len = 0;
    return STARK_OK;
  }
  const size_t m = na + nb - 1;  // mul.rs:14: length by vector length
  if (!out) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  const int lg = log2_exact(m);
  if (lg > ff::TWO_ADICITY) return stark_fail(ctx, STARK_ERR_ARG, "n > 2^23 not supported by this modulus");
  const size_t M = (size_t)1 << lg;
  u32 *da = nullptr, *db = nullptr, *fa = nullptr, *fb = nullptr;
  Scratch sc(ctx);
  ST_TRY(sc.get(&da, na * 4));
  ST_TRY(sc.get(&db, nb * 4));
  ST_TRY(sc.get(&fa, M * 4));
  ST_TRY(sc.get(&fb, M * 4));
  // uploads without a host round trip; the canonical-input flag is inspected after the download's sync
  int rc = upload_flag_reset(ctx);
  if (rc == STARK_OK) rc = upload_u64_nosync(ctx, a, na, da);
  if (rc == STARK_OK) rc = upload_u64_nosync(ctx, b, nb, db);
  ScaleSpec none = {ntt::SCALE_NONE, 1, 1};
  if (rc == STARK_OK) rc = ntt_transform(ctx, da, fa, lg, false, 1, M, M, na, none, none);
  if (rc == STARK_OK) rc = ntt_transform(ctx, db, fb, lg, false, 1, M, M, nb, none, none);
  if (rc == STARK_OK) {
    k_pointwise<<<grid_for(ctx, M, 256), 256, 0, ctx->stream>>>(fa, fb, M);
    ctx->launches++;
    // the pointwise Montgomery product left a factor R^-1: scale by M^-1 * R
    ScaleSpec post = {ntt::SCALE_CONST, ff::mul(ff::inv((u32)(M % ff::P)), ff::R1), 1};
    rc = ntt_transform(ctx, fa, fb, lg, true, 1, M, M, M, none, post);
  }
  if (rc == STARK_OK) rc = download_u64(ctx, fb, m, out);
  if (rc == STARK_OK) rc = upload_u64_check(ctx);
  if (rc == STARK_OK) *out_len = m;
  return rc;
}

int stark_poly_eval_coset(stark_ctx *ctx, const uint64_t *coeffs, size_t nc, uint64_t offset, uint32_t log_n,
                          uint64_t *out) {
  if (!ctx || !out || (nc && !coeffs)) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (log_n > (u32)ff::TWO_ADICITY) return stark_fail(ctx, STARK_ERR_ARG, "n > 2^23 not supported by this modulus");
  const size_t N = (size_t)1 << log_n;
  if (nc > N) return stark_fail(ctx, STARK_ERR_ARG, "more coefficients than domain points");
  if (offset >= ff::P) return stark_fail(ctx, STARK_ERR_ARG, "non-canonical field element (value >= p) in input");
  u32 *dc = nullptr, *dv = nullptr;
  Scratch sc(ctx);
  ST_TRY(sc.get(&dc, (nc ? nc : 1) * 4));
  ST_TRY(sc.get(&dv, N * 4));
  int rc = upload_flag_reset(ctx);
  if (rc == STARK_OK) rc = upload_u64_nosync(ctx, coeffs, nc, dc);
  ScaleSpec none = {ntt::SCALE_NONE, 1, 1};
  ScaleSpec pre = {ntt::SCALE_GEO, 1, (u32)offset};
  if (rc == STARK_OK) rc = ntt_transform(ctx, dc, dv, (int)log_n, false, 1, N, N, nc, pre, none);
  if (rc == STARK_OK) rc = download_u64(ctx, dv, N, out);
  if (rc == STARK_OK) rc = upload_u64_check(ctx);
  return rc;
}

int stark_poly_interpolate_coset(stark_ctx *ctx, const uint64_t *vals, uint64_t offset, uint32_t log_n,
                                 uint64_t *coeffs, size_t *out_len) {
  if (!ctx || !vals || !coeffs || !out_len) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (log_n > (u32)ff::TWO_ADICITY) return stark_fail(ctx, STARK_ERR_ARG, "n > 2^23 not supported by this modulus");
  if (offset == 0) return stark_fail(ctx, STARK_ERR_ARG, "no inverse");  // all points coincide (interpolate.rs:33)
  if (offset >= ff::P) return stark_fail(ctx, STARK_ERR_ARG, "non-canonical field element (value >= p) in input");
  const size_t N = (size_t)1 << log_n;
  u32 *dv = nullptr, *dc = nullptr;
  Scratch sc(ctx);
  ST_TRY(sc.get(&dv, N * 4));
  ST_TRY(sc.get(&dc, N * 4));
  int rc = upload_flag_reset(ctx);
  if (rc == STARK_OK) rc = upload_u64_nosync(ctx, vals, N, dv);
  ScaleSpec none = {ntt::SCALE_NONE, 1, 1};
  ScaleSpec post = {ntt::SCALE_GEO, ff::inv((u32)(N % ff::P)), ff::inv((u32)offset)};
  if (rc == STARK_OK) rc = ntt_transform(ctx, dv, dc, (int)log_n, true, 1, N, N, N, none, post);
  if (rc == STARK_OK) rc = download_u64(ctx, dc, N, coeffs);
  if (rc == STARK_OK) rc = upload_u64_check(ctx);
  // shape rule (add.rs:7-12, mul.rs:7-12): all-zero values -> [] for N >= 2, [0] for N == 1
  if (rc == STARK_OK) *out_len = all_zero(vals, N) ? (N == 1 ? 1 : 0) : N;
  return rc;
}

int stark_poly_eval_domain(stark_ctx *ctx, const uint64_t *coeffs, size_t nc, const uint64_t *domain, size_t m,
                           uint64_t *out) {
  if (!ctx || (nc && !coeffs) || (m && (!domain || !out))) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (m == 0) return STARK_OK;
  u32 *dc = nullptr, *dd = nullptr, *dout = nullptr;
  Scratch sc(ctx);
  ST_TRY(sc.get(&dc, (nc ? nc : 1) * 4));
  ST_TRY(sc.get(&dd, m * 4));
  ST_TRY(sc.get(&dout, m * 4));
  int rc = upload_u64(ctx, coeffs, nc, dc);
  if (rc == STARK_OK) rc = upload_u64(ctx, domain, m, dd);
  if (rc == STARK_OK) {
    k_eval_domain<<<(u32)((m + 255) / 256), 256, 0, ctx->stream>>>(dc, nc, dd, m, dout);
    ctx->launches++;
    rc = download_u64(ctx, dout, m, out);
  }
  return rc;
}

static int zerofier_dev(stark_ctx *ctx, const u32 *dx, size_t n, u32 *dm) {
  if (n > ZF_MAX) return stark_fail(ctx, STARK_ERR_ARG, "arbitrary-domain size above %d not supported", ZF_MAX);
  const size_t smem = 2 * (n + 1) * 4;
  static bool configured = false;
  if (!configured) {
    CU_TRY(ctx, cudaFuncSetAttribute(k_zerofier, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (ZF_MAX + 1) * 4));
    configured = true;
  }
  k_zerofier<<<1, 1024, smem, ctx->stream>>>(dx, (u32)n, dm);
  KERNEL_CHECK(ctx);
  return STARK_OK;
}

int stark_poly_zerofier_domain(stark_ctx *ctx, const uint64_t *domain, size_t n, uint64_t *out) {
  if (!ctx || !out || (n && !domain)) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (n == 0) return stark_fail(ctx, STARK_ERR_ARG, "empty domain");  // mod.rs:78 indexes domain[0]
  u32 *dx = nullptr, *dm = nullptr;
  Scratch sc(ctx);
  ST_TRY(sc.get(&dx, n * 4));
  ST_TRY(sc.get(&dm, (n + 1) * 4));
  int rc = upload_u64(ctx, domain, n, dx);
  if (rc == STARK_OK) rc = zerofier_dev(ctx, dx, n, dm);
  if (rc == STARK_OK) rc = download_u64(ctx, dm, n + 1, out);
  return rc;
}

int stark_poly_interpolate_domain(stark_ctx *ctx, const uint64_t *domain, const uint64_t *vals, size_t n,
                                  uint64_t *coeffs, size_t *out_len) {
  if (!ctx || !domain || !vals || !coeffs || !out_len) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (n == 0) return stark_fail(ctx, STARK_ERR_ARG, "assertion failed: domain.len() > 0");  // interpolate.rs:11
  u32 *dx = nullptr, *dy = nullptr, *dc = nullptr, *dm = nullptr, *part = nullptr, *dout = nullptr;
  const u32 rows = (u32)((n + 255) / 256);
  Scratch sc(ctx);
  ST_TRY(sc.get(&dx, n * 4));
  ST_TRY(sc.get(&dy, n * 4));
  ST_TRY(sc.get(&dc, n * 4));
  ST_TRY(sc.get(&dm, (n + 1) * 4));
  ST_TRY(sc.get(&part, (size_t)rows * n * 4));
  ST_TRY(sc.get(&dout, n * 4));
  int rc = upload_u64(ctx, domain, n, dx);
  if (rc == STARK_OK) rc = upload_u64(ctx, vals, n, dy);
  if (rc == STARK_OK) {
    cudaMemsetAsync(ctx->flag, 0, 4, ctx->stream);
    k_lagrange_weights<<<rows, 256, 0, ctx->stream>>>(dx, dy, n, dc, ctx->flag);
    ctx->launches++;
    CU_TRY(ctx, cudaMemcpyAsync(ctx->h_flag, ctx->flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (*ctx->h_flag) rc = stark_fail(ctx, STARK_ERR_ARG, "no inverse");  // duplicate points, ff.rs:171 via interpolate.rs:33
  }
  if (rc == STARK_OK) rc = zerofier_dev(ctx, dx, n, dm);
  if (rc == STARK_OK) {
    k_lagrange_accumulate<<<rows, 256, 0, ctx->stream>>>(dx, dc, dm, (u32)n, part);
    k_reduce_rows<<<(u32)((n + 255) / 256), 256, 0, ctx->stream>>>(part, rows, (u32)n, dout);
    ctx->launches += 2;
    rc = download_u64(ctx, dout, n, coeffs);
  }
  if (rc == STARK_OK) *out_len = all_zero(vals, n) ? (n == 1 ? 1 : 0) : n;
  return rc;
}

int stark_poly_scale(stark_ctx *ctx, const uint64_t *coeffs, size_t n, uint64_t factor, uint64_t *out) {
  if (!ctx || (n && (!coeffs || !out))) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (n == 0) return STARK_OK;
  if (factor >= ff::P) return stark_fail(ctx, STARK_ERR_ARG, "non-canonical field element (value >= p) in input");
  u32 *dc = nullptr, *dout = nullptr;
  Scratch sc(ctx);
  ST_TRY(sc.get(&dc, n * 4));
  ST_TRY(sc.get(&dout, n * 4));
  int rc = upload_u64(ctx, coeffs, n, dc);
  GeoTables G;
  if (rc == STARK_OK) rc = geo_tables(ctx, (u32)factor, 1, n, &G);
  if (rc == STARK_OK) {
    k_scale_geo<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(dc, dout, n, G);
    ctx->launches++;
    rc = download_u64(ctx, dout, n, out);
  }
  return rc;
}

int stark_poly_zerofier_coset(stark_ctx *ctx, uint64_t offset, uint32_t log_n, uint64_t *out) {
  if (!ctx || !out) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (log_n > (u32)ff::TWO_ADICITY) return stark_fail(ctx, STARK_ERR_ARG, "n > 2^23 not supported by this modulus");
  if (offset >= ff::P) return stark_fail(ctx, STARK_ERR_ARG, "non-canonical field element (value >= p) in input");
  // prod_i (X - offset * w^i) = X^N - offset^N : a closed form, nothing to launch
  const size_t N = (size_t)1 << log_n;
  for (size_t i = 0; i <= N; i++) out[i] = 0;
  out[0] = ff::neg(ff::pow((u32)offset, N));
  out[N] = 1;
  return STARK_OK;
}

int stark_lde_dev(stark_ctx *ctx, const stark_buf *cols, uint32_t n_cols, uint32_t log_n, uint32_t log_blowup,
                  uint64_t offset, stark_buf *out) {
  if (!ctx || !cols || !out) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (log_n + log_blowup > (u32)ff::TWO_ADICITY)
    return stark_fail(ctx, STARK_ERR_ARG, "n > 2^23 not supported by this modulus");
  if (offset >= ff::P || offset == 0) return stark_fail(ctx, STARK_ERR_ARG, "offset must be a non-zero canonical element");
  const size_t n = (size_t)1 << log_n, N = n << log_blowup;
  if (cols->n < n * n_cols || out->n < N * n_cols) return stark_fail(ctx, STARK_ERR_ARG, "buffer too small");
  return lde_dev(ctx, cols->ptr, n_cols, log_n, log_blowup, (u32)offset, out->ptr);
}

int stark_lde(stark_ctx *ctx, const uint64_t *cols, uint32_t n_cols, uint32_t log_n, uint32_t log_blowup,
              uint64_t offset, uint64_t *out) {
  if (!ctx || !cols || !out) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (log_n + log_blowup > (u32)ff::TWO_ADICITY)
    return stark_fail(ctx, STARK_ERR_ARG, "n > 2^23 not supported by this modulus");
  const size_t n = (size_t)1 << log_n, N = n << log_blowup;
  stark_buf *in = nullptr, *o = nullptr;
  ST_TRY(stark_buf_upload(ctx, cols, n * n_cols, &in));
  int rc = stark_buf_alloc(ctx, N * n_cols, &o);
  if (rc == STARK_OK) rc = stark_lde_dev(ctx, in, n_cols, log_n, log_blowup, offset, o);
  if (rc == STARK_OK) rc = download_u64(ctx, o->ptr, N * n_cols, out);
  stark_buf_free(in), stark_buf_free(o);
  return rc;
}

int stark_ntt_dev(stark_ctx *ctx, const stark_buf *in, stark_buf *out, uint32_t log_n, uint32_t batch, int inverse) {
  if (!ctx || !in || !out) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (log_n > (u32)ff::TWO_ADICITY) return stark_fail(ctx, STARK_ERR_ARG, "n > 2^23 not supported by this modulus");
  const size_t N = (size_t)1 << log_n;
  if (in->n < N * batch || out->n < N * batch) return stark_fail(ctx, STARK_ERR_ARG, "buffer too small");
  ScaleSpec none = {ntt::SCALE_NONE, 1, 1};
  ScaleSpec post = {inverse ? ntt::SCALE_CONST : ntt::SCALE_NONE, inverse ? ff::inv((u32)(N % ff::P)) : 1u, 1};
  return ntt_transform(ctx, in->ptr, out->ptr, (int)log_n, inverse != 0, batch, N, N, N, none, post);
}

}  // extern "C"
