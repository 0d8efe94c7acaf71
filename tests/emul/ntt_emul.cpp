// CPU emulation of k_ntt_single (index-math check only; TEST INFRASTRUCTURE, never shipped).
// Runs the host+device inline round functions of ntt_core.cuh thread by thread and compares with a
// reference O(n log n) NTT done with u64 %.  Build: g++ -O2 -std=c++17 -I stark-rs_b200/csrc tests/emul/ntt_emul.cpp
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "ntt_core.cuh"
using namespace ntt;
using ff::u32; using ff::u64;

static std::vector<u32> g_lo(4096), g_hi(2048);
static std::vector<wpair> g_tw[2];
static wpair g_w8[2][4];
static void init() {
  u32 w23 = ff::to_mont(ff::pow(3, (ff::P - 1) >> 23));
  for (u32 i = 0; i < 4096; i++) g_lo[i] = ff::mont_pow(w23, i);
  for (u32 i = 0; i < 2048; i++) g_hi[i] = ff::mont_pow(w23, (u64)i << 12);
  RootTables T = {g_lo.data(), g_hi.data()};
  for (int d = 0; d < 2; d++) {
    g_tw[d].assign(8192, wpair{1, ff::shoup_of(1)});
    for (u32 i = 1; i < 8192; i++) {
      int logL = 31 - __builtin_clz(i);
      u32 e = i - (1u << logL), idx = e << (23 - logL);
      if (d) idx = ((1u << 23) - idx) & ((1u << 23) - 1u);
      const u32 w = ff::from_mont(root_pow(T, idx));
      g_tw[d][i] = wpair{w, ff::shoup_of(w)};
    }
    u32 w8 = ff::pow(3, (ff::P - 1) >> 3);
    if (d) w8 = ff::inv(w8);
    for (int k = 0; k < 4; k++) g_w8[d][k] = wpair{ff::pow(w8, k), ff::shoup_of(ff::pow(w8, k))};
  }
}

// mirrors k_ntt_single in ntt.cu
static void run_single(const PassArgs &A, u32 batch) {
  const u32 nt = 1u << (A.logL - 3);
  std::vector<u32> smem((size_t)1 << A.logL), regs((size_t)nt * 8);
  int logr[4];
  const int nr = plan_rounds(A.logL, logr);
  for (u32 b = 0; b < batch; b++) {
    int logS = 0;
    for (int r = 0; r < nr; r++) {
      const bool first = r == 0, last = r == nr - 1;
      for (u32 tid = 0; tid < nt; tid++) {
        u32 *rg = &regs[(size_t)tid * 8];
        if (first) {
          if (logr[0] == 1) round_load_compute<1, true>(tid, nt, b, A, 0, smem.data(), rg);
          else if (logr[0] == 2) round_load_compute<2, true>(tid, nt, b, A, 0, smem.data(), rg);
          else round_load_compute<3, true>(tid, nt, b, A, 0, smem.data(), rg);
        } else round_load_compute<3, false>(tid, nt, b, A, logS, smem.data(), rg);
      }
      for (u32 tid = 0; tid < nt; tid++) {
        u32 *rg = &regs[(size_t)tid * 8];
        const int lr = first ? logr[0] : 3;
#define ST(LR) (last ? round_store<LR, true>(tid, nt, b, A, logS, smem.data(), rg) : round_store<LR, false>(tid, nt, b, A, logS, smem.data(), rg))
        if (lr == 1) ST(1); else if (lr == 2) ST(2); else ST(3);
      }
      logS += first ? logr[0] : 3;
    }
  }
}

// reference: in-place iterative NTT (bit reversal + DIT) with u64 %
static void ref_ntt(std::vector<u64> &a, u64 root) {
  size_t n = a.size();
  for (size_t i = 1, j = 0; i < n; i++) { size_t bit = n >> 1; for (; j & bit; bit >>= 1) j ^= bit; j ^= bit; if (i < j) std::swap(a[i], a[j]); }
  for (size_t len = 2; len <= n; len <<= 1) {
    u64 wl = ff::pow((u32)root, n / len);
    for (size_t i = 0; i < n; i += len) { u64 w = 1;
      for (size_t k = 0; k < len / 2; k++) { u64 u = a[i + k], v = a[i + k + len / 2] * w % ff::P;
        a[i + k] = (u + v) % ff::P; a[i + k + len / 2] = (u + ff::P - v) % ff::P; w = w * wl % ff::P; } }
  }
}

// field.cuh products (host paths; the device paths are the same arithmetic as PTX): ranges and residues of the Montgomery
// and Shoup forms on edge values and 2 * 10^6 pseudo-random pairs
static int check_products() {
  int bad = 0;
  u64 s = 99;
  const u32 edge[] = {0u, 1u, 2u, ff::P - 1, ff::P, ff::P + 1, ff::P2 - 1, ff::P2, 2 * ff::P2 - 1, 0x7fffffffu, 0x80000000u, 0xffffffffu};
  for (int it = 0; it < 2000000 + 144; it++) {
    u32 x, w;
    if (it < 144) {
      x = edge[it / 12], w = edge[it % 12] % ff::P;
    } else {
      s = s * 6364136223846793005ull + 1442695040888963407ull;
      x = (u32)(s >> 32);
      s = s * 6364136223846793005ull + 1442695040888963407ull;
      w = (u32)((s >> 33) % ff::P);
    }
    const u64 want = (u64)(x % ff::P) * w % ff::P;
    const u32 sh = ff::shoup_mul(x, w, ff::shoup_of(w));
    const u32 mm = ff::mont_mul(x, ff::to_mont(w));
    if (sh >= ff::P2 || sh % ff::P != want) bad++;
    if (mm >= ff::P2 || mm % ff::P != want) bad++;
    if (ff::red2p(x % (2 * ff::P2)) >= ff::P2 || ff::canon(x % ff::P2) >= ff::P) bad++;
    if (ff::add_alu(x, w, 0u) != x + w) bad++;
  }
  if (bad) printf("field product check: %d violations\n", bad);
  return bad;
}

int main(int argc, char **argv) {
  init();
  int max_log = argc > 1 ? atoi(argv[1]) : 12;
  if (max_log > 12) max_log = 12;
  int fails = check_products();
  for (int log_n = 3; log_n <= max_log; log_n++) for (int d = 0; d < 2; d++) for (int mode = 0; mode < 2; mode++) {
    const u64 N = 1ull << log_n; const u32 batch = 3;
    std::vector<u32> in(batch * N), out(batch * N);
    u64 s = 12345 + log_n;
    for (auto &x : in) { s = s * 6364136223846793005ull + 1442695040888963407ull; x = (u32)((s >> 33) % ff::P); }
    const u64 n_valid = (log_n % 2) ? N : (N / 4 + 3);
    const u32 c = ff::inv((u32)N);
    PassArgs A; memset(&A, 0, sizeof A);
    for (int k = 0; k < 4; k++) A.w8[k] = g_w8[d][k];
    A.in = in.data(), A.out = out.data(), A.logL = log_n, A.in_batch = N, A.out_batch = N, A.n_valid = n_valid;
    A.tw = g_tw[d].data() + (1u << log_n);
    A.post_mode = mode ? SCALE_CONST : SCALE_NONE, A.post_const = ff::to_mont(c);
    run_single(A, batch);
    u64 root = ff::pow(3, (ff::P - 1) >> log_n); if (d) root = ff::inv((u32)root);
    for (u32 b = 0; b < batch; b++) {
      std::vector<u64> r(N);
      for (u64 i = 0; i < N; i++) r[i] = i < n_valid ? in[b * N + i] : 0;
      ref_ntt(r, root);
      for (u64 i = 0; i < N; i++) {
        const u64 want = mode ? r[i] * c % ff::P : r[i];
        if (want != out[b * N + i]) { if (fails < 5) printf("MISMATCH log_n=%d d=%d mode=%d b=%u i=%llu got %u want %llu\n", log_n, d, mode, b, (unsigned long long)i, out[b * N + i], (unsigned long long)want); fails++; break; }
      }
    }
  }
  printf(fails ? "FAIL %d\n" : "OK %d\n", fails ? fails : max_log);
  return fails != 0;
}
