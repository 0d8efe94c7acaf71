// CPU emulation of k_ntt2_pass (index-math and range check only; TEST INFRASTRUCTURE, never shipped).
// Runs the host+device inline round functions of ntt_pass.cuh thread by thread, phase by phase (a phase boundary
// is a __syncthreads in the kernel), mirrors the pass sequencing of ntt_transform() in ntt.cu, and compares with an
// O(n log n) reference NTT done with u64 %.  Also counts shared-memory bank conflicts of every 128-bit access pattern.
// Build: g++ -O2 -std=c++17 -I stark-rs_b200/csrc tests/emul/ntt2_emul.cpp
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "ntt_pass.cuh"
using namespace ntt2;

static std::vector<u32> g_lo(4096), g_hi(2048);
static std::vector<wpair> g_tw[2];
static wpair g_w8[2][4];
static std::vector<wpair> g_otw[2], g_row[2], g_twin[2];
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
      const u32 w = ff::from_mont(ntt::root_pow(T, idx));
      g_tw[d][i] = wpair{w, ff::shoup_of(w)};
    }
    g_twin[d].resize(INNER_TWIDDLE_PAIRS);
    fill_inner_twiddles(g_twin[d].data(), d);
    // mirrors k_shoup_roots in ntt.cu
    g_otw[d].resize(1u << 16), g_row[d].resize(11 * 2048);
    for (int tab = 0; tab < 2; tab++)
      for (u32 i = 0; i < (tab ? 11u * 2048 : 1u << 16); i++) {
        u32 idx = tab ? (i & 2047u) << (10 - (i >> 11)) : i << 7;
        if (d) idx = ((1u << 23) - idx) & ((1u << 23) - 1u);
        const u32 w = ff::from_mont(ntt::root_pow(T, idx));
        (tab ? g_row[d] : g_otw[d])[i] = wpair{w, ff::shoup_of(w)};
      }
    u32 w8 = ff::pow(3, (ff::P - 1) >> 3);
    if (d) w8 = ff::inv(w8);
    for (int k = 0; k < 4; k++) g_w8[d][k] = wpair{ff::pow(w8, k), ff::shoup_of(ff::pow(w8, k))};
  }
}

static long g_range_viol = 0;

template <int LOGR, int KIND, int MODE, int TL = TILE_LOG>
static void run_pass(const PassParams &A, u32 grid) {
  typedef Plan<LOGR> PL;
  constexpr u32 NTH = 1u << (TL - 5);
  std::vector<q4> tile(1 << (TL - 2));
  std::vector<wpair> otw(1 << LOGR);
  std::vector<u32> regs((size_t)NTH * 32);
  for (u32 blk = 0; blk < grid; blk++) {
    const TileCtx T = tile_ctx<LOGR, TL>(A, blk);
    if (KIND == MIDDLE) for (u32 tid = 0; tid < NTH; tid++) fill_outer_table<LOGR, TL>(tid, A, T, otw.data());
    for (u32 tid = 0; tid < NTH; tid++) round_compute<LOGR, KIND, 0, MODE, TL>(tid, A, T, tile.data(), otw.data(), &regs[tid * 32]);
    for (u32 tid = 0; tid < NTH; tid++) round_store<LOGR, KIND, 0, TL>(tid, A, T, tile.data(), &regs[tid * 32]);
    if constexpr (PL::NR >= 3) {
      for (u32 tid = 0; tid < NTH; tid++) round_compute<LOGR, KIND, 1, MODE, TL>(tid, A, T, tile.data(), otw.data(), &regs[tid * 32]);
      for (u32 tid = 0; tid < NTH; tid++) round_store<LOGR, KIND, 1, TL>(tid, A, T, tile.data(), &regs[tid * 32]);
    }
    if constexpr (PL::NR == 4) {
      for (u32 tid = 0; tid < NTH; tid++) round_compute<LOGR, KIND, 2, MODE, TL>(tid, A, T, tile.data(), otw.data(), &regs[tid * 32]);
      for (u32 tid = 0; tid < NTH; tid++) round_store<LOGR, KIND, 2, TL>(tid, A, T, tile.data(), &regs[tid * 32]);
    }
    for (auto &v : tile) if (v.x >= ff::P2 || v.y >= ff::P2 || v.z >= ff::P2 || v.w >= ff::P2) g_range_viol++;
    for (u32 tid = 0; tid < NTH; tid++) round_compute<LOGR, KIND, PL::NR - 1, MODE, TL>(tid, A, T, tile.data(), otw.data(), &regs[tid * 32]);
    for (u32 tid = 0; tid < NTH; tid++) round_store<LOGR, KIND, PL::NR - 1, TL>(tid, A, T, tile.data(), &regs[tid * 32]);
  }
}
// mirrors k_ntt_small12 in ntt.cu: radix-2^7 FIRST pass into a second shared buffer, radix-2^5 LAST pass out of it
template <int FM, int LM>
static void run_small12(const PassParams &A, const PassParams &B, u32 batch) {
  constexpr u32 NTH = 128;
  std::vector<q4> tile(1024);
  std::vector<u32> ybuf(4096), regs((size_t)NTH * 32);
  const wpair *otw = nullptr;
  for (u32 b = 0; b < batch; b++) {
    TileCtx T1;
    T1.in = A.in + (u64)b * A.in_batch, T1.out = ybuf.data(), T1.col0 = 0, T1.q0 = 0, T1.p = 0;
    for (u32 t = 0; t < NTH; t++) round_compute<7, FIRST, 0, FM>(t, A, T1, tile.data(), otw, &regs[t * 32]);
    for (u32 t = 0; t < NTH; t++) round_store<7, FIRST, 0>(t, A, T1, tile.data(), &regs[t * 32]);
    for (u32 t = 0; t < NTH; t++) round_compute<7, FIRST, 1, FM>(t, A, T1, tile.data(), otw, &regs[t * 32]);
    for (u32 t = 0; t < NTH; t++) round_store<7, FIRST, 1>(t, A, T1, tile.data(), &regs[t * 32]);
    for (u32 t = 0; t < NTH; t++) round_compute<7, FIRST, 2, FM>(t, A, T1, tile.data(), otw, &regs[t * 32]);
    for (u32 t = 0; t < NTH; t++) round_store<7, FIRST, 2>(t, A, T1, tile.data(), &regs[t * 32]);
    TileCtx T2;
    T2.in = ybuf.data(), T2.out = B.out + (u64)b * B.out_batch, T2.col0 = 0, T2.q0 = 0, T2.p = 0;
    for (u32 t = 0; t < NTH; t++) round_compute<5, LAST, 0, LM>(t, B, T2, tile.data(), otw, &regs[t * 32]);
    for (u32 t = 0; t < NTH; t++) round_store<5, LAST, 0>(t, B, T2, tile.data(), &regs[t * 32]);
    for (u32 t = 0; t < NTH; t++) round_compute<5, LAST, 1, LM>(t, B, T2, tile.data(), otw, &regs[t * 32]);
    for (u32 t = 0; t < NTH; t++) round_store<5, LAST, 1>(t, B, T2, tile.data(), &regs[t * 32]);
  }
}
// mirrors the `log_n == 12 && batch >= 8` branch of ntt_transform() in ntt.cu
static void transform12(const u32 *in, u32 *out, int d, u32 batch, u64 n_valid, int post_mode, u32 post_c, GeoTables post_geo,
                        int pre_mode, GeoTables pre_geo) {
  const u64 N = 4096;
  static std::vector<wpair> row12[2];
  if (row12[d].empty()) {
    u32 w12 = ff::pow(3, (ff::P - 1) >> 12);
    if (d) w12 = ff::inv(w12);
    u32 v = 1;
    for (int r = 0; r < 128; r++) row12[d].push_back(wpair{v, ff::shoup_of(v)}), v = ff::mul(v, w12);
  }
  PassParams A, B;
  memset(&A, 0, sizeof A);
  A.logN = 12, A.log_tiles = 0, A.roots = {g_lo.data(), g_hi.data()}, A.inverse = d;
  for (int k = 0; k < 4; k++) A.w8[k] = g_w8[d][k];
  B = A;
  const u32 g = 3;   // the geometric scale of main(): c * 3^i
  A.in = in, A.in_batch = N, A.n_valid = n_valid, A.logS = 0;
  A.tw_in = g_twin[d].data() + inner_twiddle_offset(7), A.row_tab = row12[d].data();
  A.pre_mode = pre_mode, A.pre_geo = pre_geo;
  { const u32 gj = ff::pow(g, N >> 1); A.pre_g1 = wpair{g, ff::shoup_of(g)}, A.pre_gj = wpair{gj, ff::shoup_of(gj)}; }
  B.out = out, B.out_batch = N, B.n_valid = N, B.logS = 7;
  B.tw_in = g_twin[d].data() + inner_twiddle_offset(5);
  B.post_mode = post_mode, B.post_const = wpair{ff::from_mont(post_c), ff::shoup_of(ff::from_mont(post_c))}, B.post_geo = post_geo;
  { const u32 gk = ff::pow(g, N >> 3); B.post_g1 = wpair{g, ff::shoup_of(g)}, B.post_gk = wpair{gk, ff::shoup_of(gk)}; }
  const int fm = (n_valid < N ? 1 : 0) | (pre_mode == ntt::SCALE_GEO ? 2 : 0);
#define S12(F_) { if (post_mode == ntt::SCALE_NONE) run_small12<F_, 0>(A, B, batch); else if (post_mode == ntt::SCALE_CONST) run_small12<F_, 1>(A, B, batch); else run_small12<F_, 2>(A, B, batch); }
  if (fm == 0) S12(0) else if (fm == 1) S12(1) else S12(3)
#undef S12
}

static bool g_big = false;   // the two-pass plans on 16384-element tiles (STARK_NTT_BIG in ntt.cu)

// mirrors the N >= 2^13 branch of ntt_transform() in ntt.cu
static void transform(const u32 *in, u32 *out, int log_n, int d, u32 batch, u64 n_valid, int post_mode, u32 post_c,
                      GeoTables post_geo, int pre_mode, GeoTables pre_geo) {
  const u64 N = 1ull << log_n;
  int plan[3];
  int n_pass = pass_plan(log_n, plan);
  const bool big = g_big && log_n >= 20 && log_n <= 22;
  if (big) plan[0] = (log_n + 1) / 2, plan[1] = log_n / 2, plan[2] = 0, n_pass = 2;
  const int tile_log = big ? 14 : TILE_LOG;
  std::vector<u32> tmp((size_t)batch * N);
  const bool need_tmp = n_pass == 3 || in == out;
  PassParams B;
  memset(&B, 0, sizeof B);
  B.logN = log_n, B.log_tiles = log_n - tile_log;
  B.roots = {g_lo.data(), g_hi.data()}, B.inverse = d;
  for (int k = 0; k < 4; k++) B.w8[k] = g_w8[d][k];
  const u32 grid = batch << B.log_tiles;
  const u32 *src = in;
  int logS = 0;
  for (int i = 0; i < n_pass; i++) {
    const int r = plan[i];
    const int kind = i == 0 ? FIRST : (i == n_pass - 1 ? LAST : MIDDLE);
    u32 *dst = (kind == LAST || kind == MIDDLE || (n_pass == 2 && !need_tmp)) ? out : tmp.data();
    B.in = src, B.out = dst, B.in_batch = N, B.out_batch = N;
    B.n_valid = kind == FIRST ? n_valid : N;
    B.logS = logS;
    B.tw_in = g_twin[d].data() + inner_twiddle_offset(r);
    B.otw_tab = g_otw[d].data(), B.otw_shift = 16 - (log_n - logS);
    B.row_tab = g_row[d].data() + (log_n - 13) * 2048;
    B.pre_mode = kind == FIRST ? pre_mode : 0, B.pre_geo = pre_geo;
    B.post_mode = kind == LAST ? post_mode : 0, B.post_const = wpair{ff::from_mont(post_c), ff::shoup_of(ff::from_mont(post_c))}, B.post_geo = post_geo;
    { const u32 g = 3, gk = ff::pow(g, N >> 3);   // the geometric scale of main(): c * 3^i
      B.post_g1 = wpair{g, ff::shoup_of(g)}, B.post_gk = wpair{gk, ff::shoup_of(gk)};
      const int lr0 = plan[0] % 3 ? plan[0] % 3 : 3;
      const u32 gj = ff::pow(g, N >> lr0);
      B.pre_g1 = B.post_g1, B.pre_gj = wpair{gj, ff::shoup_of(gj)}; }
    const int mode = kind == FIRST ? ((B.n_valid < N ? 1 : 0) | (B.pre_mode == ntt::SCALE_GEO ? 2 : 0)) : (kind == LAST ? B.post_mode : 0);
#define BIGCASE(R_, K_) if (big && r == R_ && kind == K_) { if (mode == 0) run_pass<R_, K_, 0, 14>(B, grid); else if (mode == 1) run_pass<R_, K_, 1, 14>(B, grid); else if (K_ == FIRST) run_pass<R_, K_, 3, 14>(B, grid); else run_pass<R_, K_, 2, 14>(B, grid); } else
    BIGCASE(10, FIRST) BIGCASE(11, FIRST) BIGCASE(10, LAST) BIGCASE(11, LAST)
#define CASE(R_, K_) if (r == R_ && kind == K_) { if (K_ == MIDDLE || mode == 0) run_pass<R_, K_, 0>(B, grid); else if (mode == 1) run_pass<R_, K_, 1>(B, grid); else if (K_ == FIRST) run_pass<R_, K_, 3>(B, grid); else run_pass<R_, K_, 2>(B, grid); } else
    CASE(6, FIRST) CASE(7, FIRST) CASE(8, FIRST) CASE(6, MIDDLE) CASE(7, MIDDLE) CASE(8, MIDDLE)
    CASE(5, LAST) CASE(6, LAST) CASE(7, LAST) CASE(8, LAST) abort();
    src = dst;
    logS += r;
  }
}

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

// bank conflicts: for every round's load / store pattern, per quarter-warp the 8 slots must be distinct mod 8
template <int LOGR, int KIND, int ROUND, int TL = TILE_LOG>
static long conflicts() {
  typedef Plan<LOGR> PL;
  constexpr int LR = PL::lr(ROUND), LOGS = PL::logs(ROUND), RAD = 1 << LR, NTASK = 8 >> LR, LOGC4 = TL - LOGR - 2;
  constexpr u32 NT = 1u << (TL - 5);
  constexpr bool LASTR = ROUND == PL::NR - 1, ROWFAST = LASTR && KIND == FIRST;
  long bad = 0;
  for (int i = 0; i < NTASK; i++) for (u32 qw = 0; qw < NT / 8; qw++) for (int j = 0; j < RAD; j++) {
    u32 seenL = 0, seenS = 0;
    for (u32 lane = 0; lane < 8; lane++) {
      u32 up, c4; decode<LOGR, LR, ROWFAST, TL>(qw * 8 + lane + i * NT, up, c4);
      if (ROUND > 0) seenL |= 1u << (slot<LOGC4, TL>(up + ((u32)j << (LOGR - LR)), c4) & 7);
      if (!LASTR) { u32 qp = up & ((1u << LOGS) - 1), pp = up >> LOGS; seenS |= 1u << (slot<LOGC4, TL>(qp + (((pp << LR) + j) << LOGS), c4) & 7); }
    }
    if (ROUND > 0 && seenL != 0xff) bad++;
    if (!LASTR && seenS != 0xff) bad++;
  }
  return bad;
}
template <int LOGR, int KIND, int TL = TILE_LOG> static long conflicts_all() {
  long b = conflicts<LOGR, KIND, 0, TL>() + conflicts<LOGR, KIND, Plan<LOGR>::NR - 1, TL>();
  if constexpr (Plan<LOGR>::NR >= 3) b += conflicts<LOGR, KIND, 1, TL>();
  if constexpr (Plan<LOGR>::NR == 4) b += conflicts<LOGR, KIND, 2, TL>();
  return b;
}

int main(int argc, char **argv) {
  init();
  int max_log = argc > 1 ? atoi(argv[1]) : 19;
  int fails = 0;
  long bc = conflicts_all<6, FIRST>() + conflicts_all<7, FIRST>() + conflicts_all<8, FIRST>() + conflicts_all<5, LAST>() +
            conflicts_all<6, LAST>() + conflicts_all<7, LAST>() + conflicts_all<8, LAST>();
  printf("bank-conflicted quarter-warp accesses: %ld\n", bc);
  if (bc) fails++;
  const long bc_big = conflicts_all<10, FIRST, 14>() + conflicts_all<11, FIRST, 14>() + conflicts_all<10, LAST, 14>() +
                      conflicts_all<11, LAST, 14>();
  printf("bank-conflicted quarter-warp accesses (16384-element tiles): %ld\n", bc_big);
  if (bc_big) fails++;
  g_big = argc > 2 && atoi(argv[2]) != 0;
  int min_log = g_big ? 20 : 12;   // 12 = the fused small-transform kernel (transform12)
  // geometric table for scale tests: c * g^i
  const u32 g = 3, c = ff::inv(1u << 10);
  std::vector<u32> glo(4096), ghi(4096);
  for (u32 i = 0; i < 4096; i++) glo[i] = ff::canon(ff::mont_mul(ff::mont_pow(ff::to_mont(g), i), ff::to_mont(c)));
  for (u32 i = 0; i < 4096; i++) ghi[i] = ff::mont_pow(ff::to_mont(g), (u64)i << 12);
  GeoTables G = {glo.data(), ghi.data()};
  for (int log_n = min_log; log_n <= max_log; log_n++) for (int d = 0; d < 2; d++) for (int mode = 0; mode < 4; mode++) {
    if (mode >= 2 && log_n > 16 && log_n != max_log && !g_big) continue;
    const u64 N = 1ull << log_n; const u32 batch = log_n <= 15 ? 2 : 1;
    std::vector<u32> in(batch * N), out(batch * N);
    u64 s = 12345 + log_n;
    for (auto &x : in) { s = s * 6364136223846793005ull + 1442695040888963407ull; x = (u32)((s >> 33) % ff::P); }
    u64 n_valid = mode == 1 ? (N / 4 + 3) : (mode == 3 ? N / 4 : N);
    const bool inplace = mode == 3;
    int post_mode = mode == 2 ? ntt::SCALE_GEO : (mode == 3 ? ntt::SCALE_CONST : ntt::SCALE_NONE);
    int pre_mode = mode == 1 ? ntt::SCALE_GEO : ntt::SCALE_NONE;
    std::vector<u32> keep = in;
    // extra mode combinations for the small kernel: pre-scale without padding, padding without pre-scale
    if (log_n == 12 && mode == 1 && d == 1) n_valid = N;
    if (log_n == 12 && mode == 3 && d == 1) pre_mode = ntt::SCALE_GEO;
    auto run = [&](const u32 *i_, u32 *o_) {
      if (log_n == 12) transform12(i_, o_, d, batch, n_valid, post_mode, ff::to_mont(c), G, pre_mode, G);
      else transform(i_, o_, log_n, d, batch, n_valid, post_mode, ff::to_mont(c), G, pre_mode, G);
    };
    if (inplace) { run(in.data(), in.data()); out = in; in = keep; }
    else run(in.data(), out.data());
    u64 root = ff::pow(3, (ff::P - 1) >> log_n); if (d) root = ff::inv((u32)root);
    for (u32 b = 0; b < batch; b++) {
      std::vector<u64> r(N);
      for (u64 i = 0; i < N; i++) {
        u64 v = i < n_valid ? in[b * N + i] : 0;
        if (pre_mode) v = v * ff::mul(c, ff::pow(g, i)) % ff::P;
        r[i] = v;
      }
      ref_ntt(r, root);
      for (u64 i = 0; i < N; i++) {
        u64 want = r[i];
        if (post_mode == ntt::SCALE_CONST) want = want * c % ff::P;
        if (post_mode == ntt::SCALE_GEO) want = want * ff::mul(c, ff::pow(g, i)) % ff::P;
        if (want != out[b * N + i]) { if (fails < 8) printf("MISMATCH log_n=%d d=%d mode=%d b=%u i=%llu got %u want %llu\n", log_n, d, mode, b, (unsigned long long)i, out[b * N + i], (unsigned long long)want); fails++; break; }
      }
    }
  }
  printf("smem range violations (>= 2p): %ld\n", g_range_viol);
  if (g_range_viol) fails++;
  printf(fails ? "FAIL %d\n" : "OK %d\n", fails ? fails : max_log);
  return fails != 0;
}
