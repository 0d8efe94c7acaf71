// CPU check of hash.cuh (host+device inline code) against the oracle.  TEST INFRASTRUCTURE.
// Build: g++ -O2 -std=c++17 -I stark-rs_b200/csrc tests/emul/hash_emul.cpp oracle/liboracle.so
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "hash.cuh"
extern "C" int oracle_hash_from_bytes(const uint8_t *, size_t, uint8_t *);
extern "C" int oracle_hash_combine(const uint8_t *, const uint8_t *, uint8_t *);
int main() {
  int fails = 0;
  uint64_t s = 99;
  auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); };
  for (int n = 0; n < 200; n++) {
    std::vector<uint8_t> m(n);
    for (auto &x : m) x = rnd();
    uint8_t a[32], b[32];
    hs::from_bytes(m.data(), n, a);
    oracle_hash_from_bytes(m.data(), n, b);
    if (memcmp(a, b, 32)) { fails++; printf("from_bytes mismatch n=%d\n", n); }
  }
  for (int it = 0; it < 200; it++) {
    uint32_t l[8], r[8], o[8]; uint8_t b[32];
    for (int i = 0; i < 8; i++) l[i] = rnd() * 65537u + rnd(), r[i] = rnd() * 65537u + rnd();
    hs::combine(l, r, o);
    oracle_hash_combine((uint8_t *)l, (uint8_t *)r, b);
    if (memcmp(o, b, 32)) { fails++; printf("combine mismatch\n"); }
    uint32_t v = rnd() % 998244353u; uint64_t v64 = v;
    hs::leaf1(v, o);
    oracle_hash_from_bytes((uint8_t *)&v64, 8, b);
    if (memcmp(o, b, 32)) { fails++; printf("leaf mismatch\n"); }
  }
  printf(fails ? "FAIL %d\n" : "OK\n", fails);
  return fails != 0;
}
// ---- hs2 (two hashes per thread) appended check
static int check_hs2() {
  int fails = 0;
  uint64_t s = 7;
  auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 32); };
  for (int it = 0; it < 300; it++) {
    uint32_t la[8], ra[8], lb[8], rb[8], oa[8], ob[8]; uint8_t ea[32], eb[32];
    for (int i = 0; i < 8; i++) la[i] = rnd(), ra[i] = rnd(), lb[i] = rnd(), rb[i] = rnd();
    if (it == 0) for (int i = 0; i < 8; i++) la[i] = ra[i] = 0xffffffffu, lb[i] = rb[i] = 0;
    hs2::combine2(la, ra, lb, rb, oa, ob, 1);
    oracle_hash_combine((uint8_t *)la, (uint8_t *)ra, ea);
    oracle_hash_combine((uint8_t *)lb, (uint8_t *)rb, eb);
    if (memcmp(oa, ea, 32) || memcmp(ob, eb, 32)) { fails++; if (fails < 3) printf("combine2 mismatch it=%d\n", it); }
    uint32_t va = rnd() % 998244353u, vb = rnd() % 998244353u; uint64_t a64 = va, b64 = vb;
    hs2::leaf2(va, vb, oa, ob, 1);
    oracle_hash_from_bytes((uint8_t *)&a64, 8, ea);
    oracle_hash_from_bytes((uint8_t *)&b64, 8, eb);
    if (memcmp(oa, ea, 32) || memcmp(ob, eb, 32)) { fails++; if (fails < 3) printf("leaf2 mismatch it=%d\n", it); }
  }
  printf(fails ? "hs2 FAIL %d\n" : "hs2 OK\n", fails);
  return fails;
}
static int dummy_hs2 = check_hs2();
// ---- hso (one hash on eight lanes): the eight lanes run as eight host threads, a shuffle is a write to a shared slot,
// a barrier, a read of the source lane's slot, a barrier
#include <atomic>
#include <thread>
struct EmuBarrier {
  int lanes = 8;
  std::atomic<int> count{0}, gen{0};
  void wait() {
    const int g = gen.load();
    if (count.fetch_add(1) + 1 == lanes) { count.store(0); gen.fetch_add(1); }
    else while (gen.load() == g) std::this_thread::yield();
  }
};
struct EmuOct {
  uint32_t q;
  uint32_t *slots;
  EmuBarrier *bar;
  uint32_t shfl(uint32_t v, uint32_t src) const {
    slots[q] = v;
    bar->wait();
    const uint32_t r = slots[src & 7u];
    bar->wait();
    return r;
  }
};
static int check_hso() {
  int fails = 0;
  uint64_t s = 31337;
  auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 32); };
  for (int it = 0; it < 200; it++) {
    uint32_t l[8], r[8]; uint8_t want[32], got[32];
    for (int i = 0; i < 8; i++) l[i] = rnd(), r[i] = rnd();
    if (it == 0) for (int i = 0; i < 8; i++) l[i] = r[i] = 0xffffffffu;
    if (it == 1) for (int i = 0; i < 8; i++) l[i] = r[i] = 0;
    oracle_hash_combine((uint8_t *)l, (uint8_t *)r, want);
    uint32_t slots[8];
    EmuBarrier bar;
    std::thread th[8];
    for (uint32_t q = 0; q < 8; q++)
      th[q] = std::thread([&, q]() {
        EmuOct w{q, slots, &bar};
        const uint32_t o = hso::combine(w, (const uint8_t *)l, (const uint8_t *)r);
        memcpy(got + 4 * q, &o, 4);
      });
    for (auto &t : th) t.join();
    if (memcmp(got, want, 32)) { fails++; if (fails < 3) printf("hso combine mismatch it=%d\n", it); }
  }
  printf(fails ? "hso FAIL %d\n" : "hso OK\n", fails);
  return fails;
}
static int dummy_hso = check_hso();
// ---- hsq (one hash on four lanes, eight state bytes each): four host threads
struct EmuQuad {
  uint32_t q;
  uint32_t *slots;
  EmuBarrier *bar;
  uint32_t shfl(uint32_t v, uint32_t src) const {
    slots[q] = v;
    bar->wait();
    const uint32_t r = slots[src & 3u];
    bar->wait();
    return r;
  }
};
static int check_hsq() {
  int fails = 0;
  uint64_t s = 424242;
  auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 32); };
  for (int it = 0; it < 200; it++) {
    alignas(16) uint32_t l[8], r[8];
    uint8_t want[32], got[32];
    for (int i = 0; i < 8; i++) l[i] = rnd(), r[i] = rnd();
    if (it == 0) for (int i = 0; i < 8; i++) l[i] = r[i] = 0xffffffffu;
    if (it == 1) for (int i = 0; i < 8; i++) l[i] = r[i] = 0;
    oracle_hash_combine((uint8_t *)l, (uint8_t *)r, want);
    uint32_t slots[4];
    EmuBarrier bar;
    bar.lanes = 4;
    std::thread th[4];
    for (uint32_t q = 0; q < 4; q++)
      th[q] = std::thread([&, q]() {
        EmuQuad w{q, slots, &bar};
        uint32_t o0, o1;
        hsq::combine(w, (const uint8_t *)l, (const uint8_t *)r, o0, o1);
        memcpy(got + 8 * q, &o0, 4);
        memcpy(got + 8 * q + 4, &o1, 4);
      });
    for (auto &t : th) t.join();
    if (memcmp(got, want, 32)) { fails++; if (fails < 3) printf("hsq combine mismatch it=%d\n", it); }
  }
  printf(fails ? "hsq FAIL %d\n" : "hsq OK\n", fails);
  return fails;
}
static int dummy_hsq = check_hsq();
