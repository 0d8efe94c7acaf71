/*
 * fast_cpu.c -- algorithm-matched CPU baseline.  TEST / BENCH INFRASTRUCTURE ONLY.
 *
 * NOT a restatement of the reference: the reference has no NTT (it evaluates by Horner and
 * interpolates by O(n^3) Lagrange, univariate/eval.rs:6-21, interpolate.rs:6-44).  This file
 * exists so that (1) speed-ups can also be quoted against a CPU that uses the same O(n log n)
 * algorithm as the GPU, and (2) parity tests at sizes where the reference algorithm cannot
 * finish (LDE at 2^20) have a CPU checker.  It is itself validated against stark_oracle.c at
 * small sizes (tests/test_oracle_fast.py) before it is trusted as a checker.
 *
 * Same field as ff.rs:191-223: p = 998244353, g = 3, w_n = 3^((p-1)/n).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t u8;

#define P 998244353ull

static inline u64 mulm(u64 a, u64 b) { return a * b % P; } /* a,b < 2^32 */
static u64 powm(u64 b, u64 e) {
  u64 r = 1;
  b %= P;
  while (e) {
    if (e & 1) r = mulm(r, b);
    b = mulm(b, b);
    e >>= 1;
  }
  return r;
}

/* in-place iterative radix-2 DIT, natural order in and out; root = primitive n-th root */
static void ntt_inplace(u64 *a, size_t n, u64 root) {
  for (size_t i = 1, j = 0; i < n; i++) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      u64 t = a[i];
      a[i] = a[j];
      a[j] = t;
    }
  }
  u64 *tw = (u64 *)malloc((n / 2 + 1) * sizeof(u64));
  for (size_t len = 2; len <= n; len <<= 1) {
    u64 wl = powm(root, n / len);
    size_t h = len / 2;
    tw[0] = 1;
    for (size_t k = 1; k < h; k++) tw[k] = mulm(tw[k - 1], wl);
    for (size_t i = 0; i < n; i += len)
      for (size_t k = 0; k < h; k++) {
        u64 u = a[i + k], v = mulm(a[i + k + h], tw[k]);
        a[i + k] = u + v >= P ? u + v - P : u + v;
        a[i + k + h] = u >= v ? u - v : u + P - v;
      }
  }
  free(tw);
}

/* out[i] = sum_j coeffs[j] * (offset * w_N^i)^j ,  N = 2^log_n, nc <= N, natural order */
int fast_eval_coset(const u64 *coeffs, size_t nc, u64 offset, u32 log_n, u64 *out) {
  size_t n = (size_t)1 << log_n;
  if (nc > n || log_n > 23) return 1;
  u64 s = 1;
  for (size_t j = 0; j < n; j++) {
    out[j] = j < nc ? mulm(coeffs[j] % P, s) : 0;
    s = mulm(s, offset % P);
  }
  ntt_inplace(out, n, powm(3, (P - 1) / n));
  return 0;
}
/* coefficients (exactly N of them) of the interpolant through (offset * w_N^i, vals[i]) */
int fast_interpolate_coset(const u64 *vals, u64 offset, u32 log_n, u64 *coeffs) {
  size_t n = (size_t)1 << log_n;
  if (log_n > 23) return 1;
  for (size_t i = 0; i < n; i++) coeffs[i] = vals[i] % P;
  u64 w = powm(3, (P - 1) / n);
  ntt_inplace(coeffs, n, powm(w, P - 2));
  u64 ninv = powm(n % P, P - 2), oinv = powm(offset % P, P - 2), s = ninv;
  for (size_t j = 0; j < n; j++) {
    coeffs[j] = mulm(coeffs[j], s);
    s = mulm(s, oinv);
  }
  return 0;
}
/* column LDE: values on w_n^i  ->  values on offset * w_{bn}^i */
int fast_lde(const u64 *col, u32 log_n, u32 log_blowup, u64 offset, u64 *out) {
  size_t n = (size_t)1 << log_n;
  u64 *c = (u64 *)malloc(n * sizeof(u64));
  if (!c) return 1;
  int rc = fast_interpolate_coset(col, 1, log_n, c);
  if (!rc) rc = fast_eval_coset(c, n, offset, log_n + log_blowup, out);
  free(c);
  return rc;
}
/* product of two coefficient vectors, length na+nb-1 (no zero-polynomial shape rule here) */
int fast_poly_mul(const u64 *a, size_t na, const u64 *b, size_t nb, u64 *out) {
  if (!na || !nb) return 1;
  size_t m = na + nb - 1;
  u32 lg = 0;
  while (((size_t)1 << lg) < m) lg++;
  size_t n = (size_t)1 << lg;
  u64 *fa = (u64 *)calloc(n, sizeof(u64)), *fb = (u64 *)calloc(n, sizeof(u64));
  for (size_t i = 0; i < na; i++) fa[i] = a[i] % P;
  for (size_t i = 0; i < nb; i++) fb[i] = b[i] % P;
  u64 w = powm(3, (P - 1) / n);
  ntt_inplace(fa, n, w);
  ntt_inplace(fb, n, w);
  for (size_t i = 0; i < n; i++) fa[i] = mulm(fa[i], fb[i]);
  ntt_inplace(fa, n, powm(w, P - 2));
  u64 ninv = powm(n % P, P - 2);
  for (size_t i = 0; i < m; i++) out[i] = mulm(fa[i], ninv);
  free(fa);
  free(fb);
  return 0;
}
/* closed form of fri.rs:57-91:
 * out[i] = 1/2 (c[i] + c[i+h]) + (alpha mod p) / (2 offset) * w^{-i} * (c[i] - c[i+h]) */
int fast_fri_fold(const u64 *cw, size_t n, u64 alpha, u64 offset, u64 omega, u64 *out) {
  size_t h = n / 2;
  u64 inv2 = (P + 1) / 2;
  u64 k = mulm(mulm(alpha % P, powm(offset % P, P - 2)), inv2);
  u64 winv = powm(omega % P, P - 2), s = k;
  for (size_t i = 0; i < h; i++) {
    u64 a = cw[i], b = cw[i + h];
    u64 sum = (a + b) % P, dif = (a + P - b) % P;
    out[i] = (mulm(sum, inv2) + mulm(s, dif)) % P;
    s = mulm(s, winv);
  }
  return 0;
}
