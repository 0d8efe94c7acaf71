/*
 * bench_cpu.c -- CPU baseline driver for bench.py (cpu_baseline leg and --impl reference).  TEST/BENCH
 * INFRASTRUCTURE ONLY.  Runs the BASELINE config-3 pipeline (trace column -> coset LDE -> Merkle commit ->
 * Fri::prove -> ProofStream::serialize) with the oracle's restatement of the reference, one independent
 * pipeline per thread (the reference itself is single-threaded; running one instance per core is
 * the most the host can do for it).
 *   mode 0 ("port"):    LDE by the reference's own algorithms, interpolate_domain O(n^3) + eval_domain O(n*m)
 *                       (interpolate.rs:6-44, eval.rs:16-21), then fri.rs:250-311.
 *   mode 1 ("matched"): LDE by fast_cpu.c's O(n log n) NTT (NOT in the reference), hashing/Merkle/FRI as mode 0.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef uint64_t u64;
typedef uint8_t u8;
#define P 998244353ull

int oracle_lde(u64 p, const u64 *col, size_t n, size_t blowup, u64 offset, u64 *out);
int fast_lde(const u64 *col, uint32_t log_n, uint32_t log_blowup, u64 offset, u64 *out);
int oracle_ff_prim_nth_root(u64 p, u64 n, u64 *out);
int oracle_fri_prove(u64 p, const u64 *cw, size_t n, size_t domain_length, u64 omega, u64 offset, size_t ef, size_t nq,
                     u8 *out, size_t cap, size_t *out_len, u64 *top_out, u64 *alphas_out, u64 *seed_out);
int oracle_hash_from_bytes(const u8 *b, size_t n, u8 *out);

typedef struct {
  int mode, rc;
  uint32_t log_n, log_b, nq;
  u64 seed;
  u8 digest[32];
} Job;

static u64 splitmix(u64 *s) {
  u64 z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}


static void *run_timed(void *arg) {
  Job *j = (Job *)arg;
  const size_t n = (size_t)1 << j->log_n, N = n << j->log_b;
  u64 *col = (u64 *)malloc(n * 8), *lde = (u64 *)malloc(N * 8);
  u64 s = j->seed;
  for (size_t i = 0; i < n; i++) col[i] = splitmix(&s) % P;
  j->rc = j->mode == 0 ? oracle_lde(P, col, n, (size_t)1 << j->log_b, 3, lde) : fast_lde(col, j->log_n, j->log_b, 3, lde);
  u64 w = 0;
  if (!j->rc) j->rc = oracle_ff_prim_nth_root(P, N, &w);
  /* generous upper bound on the proof size: 33 R + 9 + 8 N + nq * R * (33 + 3 * (9 + 32 * 24)) */
  size_t cap = 33 * 32 + 9 + 8 * N + (size_t)j->nq * 32 * (33 + 3 * (9 + 32 * 24)), len = 0;
  u8 *proof = (u8 *)malloc(cap);
  u64 alphas[64], seedc = 0;
  u64 *top = (u64 *)malloc(8 * (j->nq + 1));
  if (!j->rc) j->rc = oracle_fri_prove(P, lde, N, N, w, 3, (size_t)1 << j->log_b, j->nq, proof, cap, &len, top, alphas, &seedc);
  if (!j->rc && len <= cap) oracle_hash_from_bytes(proof, len, j->digest);
  free(col), free(lde), free(top), free(proof);
  return NULL;
}

int oracle_bench_pipeline(int mode, uint32_t log_n, uint32_t log_b, uint32_t nq, u64 seed, int threads, double *seconds,
                          u8 *digest_out) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  pthread_t th[256];
  Job jobs[256];
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < threads; t++) {
    jobs[t].mode = mode, jobs[t].log_n = log_n, jobs[t].log_b = log_b, jobs[t].nq = nq, jobs[t].seed = seed + (u64)t, jobs[t].rc = 0;
    memset(jobs[t].digest, 0, 32);
    pthread_create(&th[t], NULL, run_timed, &jobs[t]);
  }
  int rc = 0;
  for (int t = 0; t < threads; t++) {
    pthread_join(th[t], NULL);
    rc |= jobs[t].rc;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  *seconds = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
  if (digest_out) memcpy(digest_out, jobs[0].digest, 32);
  return rc;
}
