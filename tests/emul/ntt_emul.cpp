// CPU emulation of k_ntt_pass (index-math check only; TEST INFRASTRUCTURE, never shipped).
// Runs the host+device inline round functions of ntt_core.cuh thread by thread and compares with a
// direct O(n^2) DFT.  Build: g++ -O2 -std=c++17 -I stark-rs_b200/csrc tests/emul/ntt_emul.cpp
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "ntt_core.cuh"
using namespace ntt;
using ff::u32; using ff::u64;

static std::vector<u32> g_lo(4096), g_hi(2048), g_tw[2];
static u32 g_w8[2][4];
static void init() {
  u32 w23 = ff::to_mont(ff::pow(3, (ff::P - 1) >> 23));
  for (u32 i = 0; i < 4096; i++) g_lo[i] = ff::mont_pow(w23, i);
  for (u32 i = 0; i < 2048; i++) g_hi[i] = ff::mont_pow(w23, (u64)i << 12);
  RootTables T = {g_lo.data(), g_hi.data()};
  for (int d = 0; d < 2; d++) {
    g_tw[d].assign(8192, ff::R1);
    for (u32 i = 1; i < 8192; i++) {
      int logL = 31 - __builtin_clz(i);
      u32 e = i - (1u << logL), idx = e << (23 - logL);
      if (d) idx = ((1u << 23) - idx) & ((1u << 23) - 1u);
      g_tw[d][i] = root_pow(T, idx);
    }
    u32 w8 = ff::pow(3, (ff::P - 1) >> 3);
    if (d) w8 = ff::inv(w8);
    g_w8[d][0] = ff::R1;
    for (int k = 1; k < 4; k++) g_w8[d][k] = ff::to_mont(ff::pow(w8, k));
  }
}

template <int V, bool ROWOUT>
static void run_pass(const PassArgs &A, u32 tiles) {
  typedef typename Slot<V>::type slot_t;
  const u32 nt = (1u << (A.logL - 3)) << A.logC4;
  std::vector<slot_t> smem((size_t)1 << (A.logL + A.logC4));
  std::vector<u32> regs((size_t)nt * 32);
  int logr[4];
  const int nr = plan_rounds(A.logL, logr);
  for (u32 tile = 0; tile < tiles; tile++) {
    int logS = 0;
    for (int r = 0; r < nr; r++) {
      const bool first = r == 0, last = r == nr - 1;
      for (u32 tid = 0; tid < nt; tid++) {
        u32 *rg = &regs[(size_t)tid * 32];
        if (first) {
          if (logr[0] == 1) round_load_compute<1, V, true>(tid, nt, tile, A, 0, smem.data(), last && ROWOUT, rg);
          else if (logr[0] == 2) round_load_compute<2, V, true>(tid, nt, tile, A, 0, smem.data(), last && ROWOUT, rg);
          else round_load_compute<3, V, true>(tid, nt, tile, A, 0, smem.data(), last && ROWOUT, rg);
        } else round_load_compute<3, V, false>(tid, nt, tile, A, logS, smem.data(), last && ROWOUT, rg);
      }
      for (u32 tid = 0; tid < nt; tid++) {
        u32 *rg = &regs[(size_t)tid * 32];
        const int lr = first ? logr[0] : 3;
#define ST(LR) (last ? round_store<LR, V, true, ROWOUT>(tid, nt, tile, A, logS, smem.data(), rg) \
                     : round_store<LR, V, false, ROWOUT>(tid, nt, tile, A, logS, smem.data(), rg))
        if (lr == 1) ST(1); else if (lr == 2) ST(2); else ST(3);
      }
      logS += first ? logr[0] : 3;
    }
  }
}

// mirrors ntt_transform() in ntt.cu for log_n >= 3 (no scaling)
static void transform(const u32 *in, u32 *out, int log_n, int d, u32 batch, u64 n_valid) {
  const u64 N = 1ull << log_n;
  PassArgs A; memset(&A, 0, sizeof A);
  A.roots = {g_lo.data(), g_hi.data()};
  A.shiftN = 23 - log_n; A.inverse = d;
  for (int k = 0; k < 4; k++) A.w8[k] = g_w8[d][k];
  if (log_n <= 12) {
    A.in = in, A.out = out, A.logL = log_n, A.logC4 = 0;
    A.in_batch = N, A.in_stride = 1, A.n_valid = n_valid, A.out_batch = N, A.out_stride = 1, A.tiles_per_batch = 1;
    A.tw = g_tw[d].data() + (1u << log_n);
    run_pass<1, false>(A, batch);
    return;
  }
  const int l1 = log_n / 2, l2 = log_n - l1;
  const u64 N1 = 1ull << l1, N2 = 1ull << l2;
  { PassArgs B = A; B.in = in, B.out = out, B.logL = l1;
    int logC = l2 < (15 - l1) ? l2 : (15 - l1); B.logC4 = logC - 2;
    B.in_batch = N, B.in_stride = N2, B.n_valid = n_valid, B.out_batch = N; B.tiles_per_batch = (int)(N2 >> logC);
    B.tw = g_tw[d].data() + (1u << l1);
    run_pass<4, true>(B, batch * B.tiles_per_batch); }
  { PassArgs B = A; B.in = out, B.out = out, B.logL = l2;
    int logC = l1 < (15 - l2) ? l1 : (15 - l2); B.logC4 = logC - 2;
    B.in_batch = N, B.in_stride = N1, B.n_valid = N, B.out_batch = N, B.out_stride = N1; B.tiles_per_batch = (int)(N1 >> logC);
    B.tw = g_tw[d].data() + (1u << l2);
    run_pass<4, false>(B, batch * B.tiles_per_batch); }
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

int main(int argc, char **argv) {
  init();
  int max_log = argc > 1 ? atoi(argv[1]) : 16;
  int fails = 0;
  for (int log_n = 3; log_n <= max_log; log_n++) for (int d = 0; d < 2; d++) {
    const u64 N = 1ull << log_n; const u32 batch = log_n <= 12 ? 3 : 2;
    std::vector<u32> in(batch * N), out(batch * N);
    u64 s = 12345 + log_n;
    for (auto &x : in) { s = s * 6364136223846793005ull + 1442695040888963407ull; x = (u32)((s >> 33) % ff::P); }
    u64 n_valid = (log_n % 2) ? N : (N / 4 + 3);
    transform(in.data(), out.data(), log_n, d, batch, n_valid);
    u64 root = ff::pow(3, (ff::P - 1) >> log_n); if (d) root = ff::inv((u32)root);
    for (u32 b = 0; b < batch; b++) {
      std::vector<u64> r(N);
      for (u64 i = 0; i < N; i++) r[i] = i < n_valid ? in[b * N + i] : 0;
      ref_ntt(r, root);
      for (u64 i = 0; i < N; i++) if (r[i] != out[b * N + i]) { if (fails < 5) printf("MISMATCH log_n=%d d=%d b=%u i=%llu got %u want %llu\n", log_n, d, b, (unsigned long long)i, out[b * N + i], (unsigned long long)r[i]); fails++; break; }
    }
  }
  printf(fails ? "FAIL %d\n" : "OK %d\n", fails ? fails : max_log);
  return fails != 0;
}
