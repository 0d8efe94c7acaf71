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
