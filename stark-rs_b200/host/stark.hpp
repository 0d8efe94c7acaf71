// stark.hpp -- C++ host-side mirror of the reference's public interface for the hot path, on top of the C ABI
// (include/stark_b200.h).  The reference is Rust (src/ff.rs, src/univariate, src/hash.rs, src/merkle.rs,
// src/fiat_shamir.rs, src/stream.rs, src/fri.rs); there is no Rust toolchain in this image, so this header is the
// executable stand-in for stark-rs_b200/rust/: same type and method names, same argument meaning, and a reference
// assert!/panic! becomes stark::Panic carrying the reference's message.  Header-only; link with -lstark_b200.
//
// Scalar FieldElement arithmetic and single digests stay on the host (a GPU call per scalar is meaningless);
// everything that touches a whole vector / tree / codeword forwards to the GPU.  There is no CPU fallback for those.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/stark_b200.h"
#include "../csrc/hash.cuh"  // host-compilable: Hash::from_bytes for single digests / transcripts

namespace stark {

struct Panic : std::runtime_error {
  using std::runtime_error::runtime_error;
};
inline void check(int status) {
  if (status == STARK_OK) return;
  const std::string msg = stark_last_error();
  if (status == STARK_ERR_ARG) throw Panic(msg);
  throw std::runtime_error("stark_b200: " + msg);
}
// one context per thread of use
inline stark_ctx *ctx() {
  thread_local stark_ctx *c = nullptr;
  if (!c) check(stark_ctx_create(0, &c));
  return c;
}

// ------------------------------------------------------------------------------------------------ ff.rs
struct FieldElement;
struct FiniteField {
  uint64_t p;
  explicit FiniteField(uint64_t p_) : p(p_) {}
  uint64_t modulus() const { return p; }
  inline FieldElement new_element(uint64_t v) const;  // not reduced (ff.rs:113-118)
  inline FieldElement zero() const;
  inline FieldElement one() const;
  inline FieldElement mul(const FieldElement &l, const FieldElement &r) const;
  inline FieldElement add(const FieldElement &l, const FieldElement &r) const;
  inline FieldElement sub(const FieldElement &l, const FieldElement &r) const;
  inline FieldElement neg(const FieldElement &x) const;
  inline FieldElement inv(const FieldElement &x) const;
  inline FieldElement div(const FieldElement &l, const FieldElement &r) const;
  inline FieldElement exp(const FieldElement &b, uint64_t e) const;
  inline FieldElement g() const;
  inline FieldElement prim_nth_root(uint64_t n) const;
  inline FieldElement sample(const std::vector<uint8_t> &salt) const;
  bool operator==(const FiniteField &o) const { return p == o.p; }
};
struct FieldElement {
  uint64_t value;
  FiniteField field;
  FieldElement operator+(const FieldElement &r) const { return field.add(*this, r); }
  FieldElement operator-(const FieldElement &r) const { return field.sub(*this, r); }
  FieldElement operator*(const FieldElement &r) const { return field.mul(*this, r); }
  FieldElement operator/(const FieldElement &r) const { return field.div(*this, r); }
  FieldElement operator-() const { return field.neg(*this); }
  FieldElement operator^(uint64_t e) const { return field.exp(*this, e); }
  FieldElement pow(uint64_t e) const { return field.exp(*this, e); }
  bool operator==(const FieldElement &o) const { return value == o.value && field == o.field; }
  bool operator!=(const FieldElement &o) const { return !(*this == o); }
};
typedef unsigned __int128 u128_t;
inline FieldElement FiniteField::new_element(uint64_t v) const { return FieldElement{v, *this}; }
inline FieldElement FiniteField::zero() const { return new_element(0); }
inline FieldElement FiniteField::one() const { return new_element(1); }
inline FieldElement FiniteField::mul(const FieldElement &l, const FieldElement &r) const {
  return new_element((uint64_t)(((u128_t)l.value * r.value) % p));
}
inline FieldElement FiniteField::add(const FieldElement &l, const FieldElement &r) const {
  return new_element((uint64_t)(((u128_t)l.value + r.value) % p));
}
inline FieldElement FiniteField::sub(const FieldElement &l, const FieldElement &r) const {
  return new_element((uint64_t)(((u128_t)p + l.value - r.value) % p));
}
inline FieldElement FiniteField::neg(const FieldElement &x) const { return new_element((p - x.value) % p); }
inline FieldElement FiniteField::inv(const FieldElement &x) const {
  // extended Euclid, iterative (utils.rs:3-13 is the recursive form); g != 1 <=> "no inverse" (ff.rs:171)
  __int128 r0 = x.value, r1 = p, s0 = 1, s1 = 0;
  while (r1 != 0) {
    const __int128 q = r0 / r1, r2 = r0 - q * r1, s2 = s0 - q * s1;
    r0 = r1, r1 = r2, s0 = s1, s1 = s2;
  }
  if (r0 != 1) throw Panic("no inverse");
  const __int128 pp = p;
  return new_element((uint64_t)(((s0 % pp) + pp) % pp));
}
inline FieldElement FiniteField::div(const FieldElement &l, const FieldElement &r) const {
  if (r.value == 0) throw Panic("no division by zero");
  return mul(l, inv(r));
}
inline FieldElement FiniteField::exp(const FieldElement &b, uint64_t e) const {
  FieldElement acc = one(), base = b;
  for (; e; e >>= 1) {
    if (e & 1) acc = mul(acc, base);
    base = mul(base, base);
  }
  return acc;
}
inline FieldElement FiniteField::g() const {
  if (p != STARK_P) throw Panic("assertion failed: self.p == 998244353");
  return new_element(3);
}
inline FieldElement FiniteField::prim_nth_root(uint64_t n) const {
  if (p != STARK_P) throw Panic("assertion failed: self.p == 998244353");
  uint64_t w = 0;
  check(stark_ff_prim_nth_root(n, &w));
  return new_element(w);
}
inline FieldElement FiniteField::sample(const std::vector<uint8_t> &salt) const {
  u128_t acc = 0;
  for (uint8_t b : salt) acc = ((acc << 8) % p), acc = (acc ^ b) % p;
  return new_element((uint64_t)acc);
}

inline std::vector<uint64_t> raw(const std::vector<FieldElement> &v) {
  std::vector<uint64_t> r(v.size());
  for (size_t i = 0; i < v.size(); i++) r[i] = v[i].value;
  return r;
}
inline std::vector<FieldElement> wrap(const std::vector<uint64_t> &v, const FiniteField &f) {
  std::vector<FieldElement> r;
  r.reserve(v.size());
  for (uint64_t x : v) r.push_back(f.new_element(x));
  return r;
}

// ------------------------------------------------------------------------------------------ univariate
struct Polynomial {
  std::vector<FieldElement> coeffs;  // low -> high (mod.rs:8-11)
  FiniteField field;
  Polynomial(std::vector<FieldElement> c, FiniteField f) : coeffs(std::move(c)), field(f) {}
  long long deg() const {  // mod.rs:54-68
    for (size_t i = coeffs.size(); i-- > 0;)
      if (coeffs[i].value != 0) return (long long)i;
    return -1;
  }
  bool is_zero() const { return deg() == -1; }
  FieldElement leading_coeff() const {
    if (is_zero()) throw Panic("Zero polynomial has no leading coefficient");
    return coeffs[(size_t)deg()];
  }
  // mul.rs:6-29 -> stark_poly_mul
  static Polynomial mul(const Polynomial &l, const Polynomial &r) {
    const auto a = raw(l.coeffs), b = raw(r.coeffs);
    std::vector<uint64_t> out(a.size() + b.size() + 1);
    size_t n = 0;
    check(stark_poly_mul(ctx(), a.data(), a.size(), b.data(), b.size(), out.data(), &n));
    out.resize(n);
    return Polynomial(wrap(out, l.field), l.field);
  }
  // exp.rs:6-33: square and multiply over mul (each product one stark_poly_mul); exp 0 -> [1], zero base -> []
  static Polynomial exp(const Polynomial &base, uint64_t e) {
    if (e == 0) return Polynomial({base.field.one()}, base.field);
    if (base.is_zero()) return Polynomial({}, base.field);
    Polynomial result({base.field.one()}, base.field), bpower = base;
    for (; e; e >>= 1) {
      if (e & 1) result = mul(result, bpower);
      bpower = mul(bpower, bpower);
    }
    return result;
  }
  FieldElement eval(const FieldElement &x) const {  // eval.rs:6-14
    FieldElement xi = x.field.one(), val = x.field.zero();
    for (const auto &c : coeffs) val = val + c * xi, xi = xi * x;
    return val;
  }
  // (offset, log_n) when domain[i] = offset * w_N^i (fri.rs:575-578)
  static bool as_coset(const FiniteField &f, const std::vector<FieldElement> &d, uint64_t *off, uint32_t *lg) {
    const size_t n = d.size();
    if (n == 0 || (n & (n - 1)) || n > (1u << 23) || d[0].value == 0 || f.p != STARK_P) return false;
    const FieldElement w = f.prim_nth_root(n);
    FieldElement x = d[0];
    for (const auto &e : d) {
      if (e.value != x.value) return false;
      x = x * w;
    }
    *off = d[0].value, *lg = (uint32_t)__builtin_ctzll(n);
    return true;
  }
  std::vector<FieldElement> eval_domain(const std::vector<FieldElement> &domain) const {  // eval.rs:16-21
    const auto c = raw(coeffs);
    std::vector<uint64_t> out(domain.size());
    uint64_t off;
    uint32_t lg;
    if (as_coset(field, domain, &off, &lg) && c.size() <= domain.size()) {
      check(stark_poly_eval_coset(ctx(), c.data(), c.size(), off, lg, out.data()));
    } else {
      const auto d = raw(domain);
      check(stark_poly_eval_domain(ctx(), c.data(), c.size(), d.data(), d.size(), out.data()));
    }
    return wrap(out, field);
  }
  static Polynomial interpolate_domain(const std::vector<FieldElement> &domain, const std::vector<FieldElement> &values) {
    if (domain.size() != values.size()) throw Panic("assertion failed: domain.len() == values.len()");
    if (domain.empty()) throw Panic("assertion failed: domain.len() > 0");
    const FiniteField f = domain[0].field;
    const auto v = raw(values);
    std::vector<uint64_t> out(domain.size());
    size_t n = 0;
    uint64_t off;
    uint32_t lg;
    if (as_coset(f, domain, &off, &lg)) {
      check(stark_poly_interpolate_coset(ctx(), v.data(), off, lg, out.data(), &n));
    } else {
      const auto d = raw(domain);
      check(stark_poly_interpolate_domain(ctx(), d.data(), v.data(), d.size(), out.data(), &n));
    }
    out.resize(n);
    return Polynomial(wrap(out, f), f);
  }
  // div.rs:6-53: (quotient, remainder); panics "No division by zero"
  static std::pair<Polynomial, Polynomial> div(const Polynomial &numer, const Polynomial &denom) {
    const auto a = raw(numer.coeffs), b = raw(denom.coeffs);
    std::vector<uint64_t> q(a.size() + 1), r(a.size() + b.size() + 1);
    size_t nq = 0, nr = 0;
    check(stark_poly_div(ctx(), a.data(), a.size(), b.data(), b.size(), q.data(), &nq, r.data(), &nr));
    q.resize(nq), r.resize(nr);
    return {Polynomial(wrap(q, numer.field), numer.field), Polynomial(wrap(r, numer.field), numer.field)};
  }
  static Polynomial intdiv(const Polynomial &numer, const Polynomial &denom) {  // div.rs:43-47
    auto qr = div(numer, denom);
    for (const auto &c : qr.second.coeffs)
      if (c.value != 0) throw Panic("assertion failed: r.is_zero()");
    return qr.first;
  }
  static Polynomial modulo(const Polynomial &numer, const Polynomial &denom) { return div(numer, denom).second; }  // div.rs:49-52
  static Polynomial zerofier(const std::vector<FieldElement> &domain) {  // mod.rs:77-96
    const auto d = raw(domain);
    std::vector<uint64_t> out(d.size() + 1);
    check(stark_poly_zerofier_domain(ctx(), d.data(), d.size(), out.data()));
    return Polynomial(wrap(out, domain.at(0).field), domain[0].field);
  }
  Polynomial scale(const FieldElement &factor) const {  // mod.rs:99-113
    const auto c = raw(coeffs);
    std::vector<uint64_t> out(c.size());
    check(stark_poly_scale(ctx(), c.data(), c.size(), factor.value, out.data()));
    return Polynomial(wrap(out, field), field);
  }
};

// -------------------------------------------------------------------------------------- hash / merkle
struct Hash {
  uint8_t b[32];
  static Hash from_bytes(const uint8_t *m, size_t n) {  // hash.rs:7-30 (single digest: host)
    Hash h;
    hs::from_bytes(m, n, h.b);
    return h;
  }
  static Hash from_bytes(const std::vector<uint8_t> &m) { return from_bytes(m.data(), m.size()); }
  static Hash from_field_elements(const std::vector<uint64_t> &e) {  // hash.rs:32-35
    std::vector<uint8_t> m(8 * e.size());
    for (size_t i = 0; i < e.size(); i++)
      for (int k = 0; k < 8; k++) m[8 * i + k] = (uint8_t)(e[i] >> (8 * k));
    return from_bytes(m);
  }
  static Hash from_u64(uint64_t v) { return from_field_elements({v}); }
  static Hash combine(const Hash &l, const Hash &r) {  // hash.rs:41-46
    uint8_t m[64];
    memcpy(m, l.b, 32), memcpy(m + 32, r.b, 32);
    return from_bytes(m, 64);
  }
  // bulk: leaf i = from_field_elements(vals[i*width ..]) on the GPU (fri.rs:118-121)
  static std::vector<Hash> leaves(const std::vector<uint64_t> &vals, uint32_t width = 1) {
    std::vector<Hash> out(vals.size() / width);
    check(stark_hash_leaves(ctx(), vals.data(), out.size(), width, out.empty() ? nullptr : out[0].b));
    return out;
  }
  std::string to_hex() const {
    static const char *d = "0123456789abcdef";
    std::string s;
    for (uint8_t x : b) s += d[x >> 4], s += d[x & 15];
    return s;
  }
  bool operator==(const Hash &o) const { return memcmp(b, o.b, 32) == 0; }
  bool operator!=(const Hash &o) const { return !(*this == o); }
};
static_assert(sizeof(Hash) == 32, "Hash must be 32 contiguous bytes");

struct MerkleTree {  // merkle.rs:4-97; all levels live on the device
  stark_tree *h = nullptr;
  std::vector<Hash> leaves;
  Hash root;
  explicit MerkleTree(const std::vector<Hash> &lv) : leaves(lv) {
    check(stark_merkle_build(ctx(), lv.empty() ? nullptr : lv[0].b, lv.size(), &h));
    check(stark_merkle_root(h, root.b));
  }
  MerkleTree(const MerkleTree &) = delete;
  ~MerkleTree() { stark_merkle_free(h); }
  const Hash &get_root() const { return root; }
  static Hash commit(const std::vector<Hash> &lv) {
    Hash r;
    check(stark_merkle_commit(ctx(), lv.empty() ? nullptr : lv[0].b, lv.size(), r.b));
    return r;
  }
  std::vector<Hash> level(uint32_t l) const {  // nodes[l]
    std::vector<Hash> out(leaves.size() >> l);
    check(stark_merkle_level(h, l, out[0].b));
    return out;
  }
  std::vector<Hash> open(size_t index) const {
    std::vector<Hash> out(stark_merkle_num_levels(h));
    size_t n = 0;
    check(stark_merkle_open(h, index, out[0].b, &n));
    out.resize(n);
    return out;
  }
  // MerkleTree::open for many leaves in one device gather (stark_merkle_open_batch)
  std::vector<std::vector<Hash>> open_batch(const std::vector<size_t> &indices) const {
    const size_t depth = stark_merkle_num_levels(h) - 1;
    std::vector<uint64_t> idx(indices.begin(), indices.end());
    std::vector<Hash> flat(depth * idx.size() + 1);
    check(stark_merkle_open_batch(h, idx.data(), idx.size(), flat[0].b));
    std::vector<std::vector<Hash>> out;
    for (size_t q = 0; q < idx.size(); q++) out.emplace_back(flat.begin() + q * depth, flat.begin() + (q + 1) * depth);
    return out;
  }
  static bool verify(const Hash &leaf, size_t index, const std::vector<Hash> &proof, const Hash &root) {
    Hash cur = leaf;
    for (const Hash &s : proof) {
      cur = (index & 1) ? Hash::combine(s, cur) : Hash::combine(cur, s);
      index >>= 1;
    }
    return cur == root;
  }
};

// --------------------------------------------------------------------------- fiat_shamir.rs / stream.rs
struct FiatShamir {
  std::vector<uint8_t> transcript;
  void absorb(const uint8_t *d, size_t n) { transcript.insert(transcript.end(), d, d + n); }
  FieldElement challenge(const FiniteField &f) const {  // fiat_shamir.rs:19-25: unreduced
    const Hash h = Hash::from_bytes(transcript);
    uint64_t v = 0;
    for (int k = 0; k < 8; k++) v |= (uint64_t)h.b[k] << (8 * k);
    return f.new_element(v);
  }
};

struct ProofObject {
  enum Kind { MerkleRoot = 0, Element = 1, Elements = 2, MerklePath = 3 } kind;
  std::vector<uint64_t> values;  // Element / Elements
  std::vector<Hash> hashes;      // MerkleRoot / MerklePath
};
struct ProofStream {
  std::vector<ProofObject> objects;
  void push(ProofObject o) { objects.push_back(std::move(o)); }
  std::vector<uint8_t> serialize() const {  // stream.rs:35-64
    std::vector<uint8_t> b;
    auto u64le = [&](uint64_t v) { for (int k = 0; k < 8; k++) b.push_back((uint8_t)(v >> (8 * k))); };
    for (const auto &o : objects) {
      b.push_back((uint8_t)o.kind);
      if (o.kind == ProofObject::MerkleRoot) b.insert(b.end(), o.hashes[0].b, o.hashes[0].b + 32);
      if (o.kind == ProofObject::Element) u64le(o.values[0]);
      if (o.kind == ProofObject::Elements) { u64le(o.values.size()); for (uint64_t v : o.values) u64le(v); }
      if (o.kind == ProofObject::MerklePath) { u64le(o.hashes.size()); for (const Hash &h : o.hashes) b.insert(b.end(), h.b, h.b + 32); }
    }
    return b;
  }
  static ProofStream deserialize(const std::vector<uint8_t> &b) {  // stream.rs:66-168 (lenient)
    ProofStream s;
    auto rd = [&](size_t i) { uint64_t v = 0; for (int k = 0; k < 8; k++) v |= (uint64_t)b[i + k] << (8 * k); return v; };
    size_t i = 0;
    while (i < b.size()) {
      const uint8_t tag = b[i++];
      ProofObject o;
      if (tag == 0) {
        if (i + 32 > b.size()) continue;
        o.kind = ProofObject::MerkleRoot, o.hashes.resize(1), memcpy(o.hashes[0].b, &b[i], 32), i += 32;
      } else if (tag == 1) {
        if (i + 8 > b.size()) continue;
        o.kind = ProofObject::Element, o.values = {rd(i)}, i += 8;
      } else if (tag == 2 || tag == 3) {
        if (i + 8 > b.size()) continue;
        const uint64_t n = rd(i);
        i += 8;
        o.kind = tag == 2 ? ProofObject::Elements : ProofObject::MerklePath;
        for (uint64_t k = 0; k < n; k++) {
          if (tag == 2 && i + 8 <= b.size()) o.values.push_back(rd(i)), i += 8;
          if (tag == 3 && i + 32 <= b.size()) { Hash h; memcpy(h.b, &b[i], 32); o.hashes.push_back(h); i += 32; }
        }
      } else {
        break;
      }
      s.push(std::move(o));
    }
    return s;
  }
};

// ------------------------------------------------------------------------------------------------ fri.rs
struct Fri {
  FieldElement offset, omega;
  size_t domain_length;
  FiniteField field;
  size_t expansion_factor, num_colinearity_tests;
  Fri(FieldElement omega_, FieldElement offset_, size_t n, size_t ef, size_t nq)
      : offset(offset_), omega(omega_), domain_length(n), field(omega_.field), expansion_factor(ef), num_colinearity_tests(nq) {
    uint32_t r;
    check(stark_fri_num_rounds(n, (uint32_t)ef, (uint32_t)nq, &r));  // the three Fri::new asserts (fri.rs:37-45)
  }
  uint64_t num_rounds() const {
    uint32_t r = 0;
    check(stark_fri_num_rounds(domain_length, (uint32_t)expansion_factor, (uint32_t)num_colinearity_tests, &r));
    return r;
  }
  std::vector<FieldElement> fold_codeword(const std::vector<FieldElement> &cw, const FieldElement &alpha,
                                          const FieldElement &off, const FieldElement &om) const {  // fri.rs:57-91
    const auto v = raw(cw);
    std::vector<uint64_t> out(v.size() / 2);
    check(stark_fri_fold(ctx(), v.data(), v.size(), alpha.value, off.value, om.value, out.data()));
    return wrap(out, field);
  }
  // fri.rs:105-156: the round loop runs on the device (stark_fri_commit); roots and the last codeword are pushed /
  // absorbed in the reference's order and every intermediate codeword is returned
  std::vector<std::vector<FieldElement>> commit(const std::vector<FieldElement> &initial_codeword, ProofStream &proof_stream,
                                                FiatShamir &fiat_shamir) const {
    const auto v = raw(initial_codeword);
    stark_fri_state *st = nullptr;
    check(stark_fri_commit(ctx(), v.data(), v.size(), offset.value, omega.value, (uint32_t)expansion_factor,
                           (uint32_t)num_colinearity_tests, fiat_shamir.transcript.data(), fiat_shamir.transcript.size(), &st));
    const size_t R = stark_fri_rounds(st);
    std::vector<uint8_t> roots(32 * (R ? R : 1));
    check(stark_fri_roots(st, roots.data()));
    for (size_t r = 0; r < R; r++) {
      ProofObject o{ProofObject::MerkleRoot, {}, {Hash()}};
      memcpy(o.hashes[0].b, &roots[32 * r], 32);
      proof_stream.push(o);                     // fri.rs:129-131
      fiat_shamir.absorb(&roots[32 * r], 32);
    }
    std::vector<std::vector<FieldElement>> codewords;
    for (size_t r = 0; r < (R ? R : 1); r++) {
      size_t len = 0;
      check(stark_fri_codeword_len(st, (uint32_t)r, &len));
      std::vector<uint64_t> cw(len);
      check(stark_fri_codeword(st, (uint32_t)r, cw.data()));
      codewords.push_back(wrap(cw, field));
    }
    proof_stream.push(ProofObject{ProofObject::Elements, raw(codewords.back()), {}});   // fri.rs:151
    stark_fri_free(st);
    return codewords;
  }
  // fri.rs:176-213 (same asserts, same text)
  std::vector<size_t> sample_indices(const std::vector<uint8_t> &seed, size_t size, size_t reduced_size, size_t number) const {
    std::vector<uint64_t> out(number ? number : 1);
    check(stark_fri_sample_indices(seed.data(), seed.size(), size, reduced_size, number, out.data()));
    return std::vector<size_t>(out.begin(), out.begin() + number);
  }
  // fri.rs:215-248: triples, then the authentication paths (one batched device gather per tree)
  std::vector<size_t> query(const std::vector<FieldElement> &current_codeword, const std::vector<FieldElement> &next_codeword,
                            const std::vector<size_t> &c_indices, ProofStream &proof_stream, const MerkleTree &current_tree,
                            const MerkleTree &next_tree) const {
    const size_t half = current_codeword.size() / 2;
    std::vector<size_t> a_indices = c_indices, b_indices;
    for (size_t i : a_indices) b_indices.push_back(i + half);
    for (size_t s = 0; s < num_colinearity_tests; s++)
      proof_stream.push(ProofObject{ProofObject::Elements,
                                    {current_codeword[a_indices[s]].value, current_codeword[b_indices[s]].value, next_codeword[c_indices[s]].value},
                                    {}});
    const auto pa = current_tree.open_batch(a_indices), pb = current_tree.open_batch(b_indices), pc = next_tree.open_batch(c_indices);
    for (size_t s = 0; s < num_colinearity_tests; s++) {
      proof_stream.push(ProofObject{ProofObject::MerklePath, {}, pa[s]});
      proof_stream.push(ProofObject{ProofObject::MerklePath, {}, pb[s]});
      proof_stream.push(ProofObject{ProofObject::MerklePath, {}, pc[s]});
    }
    a_indices.insert(a_indices.end(), b_indices.begin(), b_indices.end());
    return a_indices;
  }
  // fri.rs:250-311: returns top_level_indices; proof_stream receives the objects, fiat_shamir the roots
  std::vector<size_t> prove(const std::vector<FieldElement> &initial_codeword, FiatShamir &fiat_shamir,
                            ProofStream &proof_stream) const {
    const auto v = raw(initial_codeword);
    size_t cap = 0, len = 0;
    check(stark_fri_proof_size(domain_length, (uint32_t)expansion_factor, (uint32_t)num_colinearity_tests, &cap));
    std::vector<uint8_t> proof(cap ? cap : 1);
    std::vector<uint64_t> top(num_colinearity_tests ? num_colinearity_tests : 1);
    check(stark_fri_prove(ctx(), v.data(), v.size(), domain_length, offset.value, omega.value, (uint32_t)expansion_factor,
                          (uint32_t)num_colinearity_tests, fiat_shamir.transcript.data(), fiat_shamir.transcript.size(),
                          proof.data(), cap, &len, top.data()));
    proof.resize(len);
    for (auto &o : ProofStream::deserialize(proof).objects) {
      if (o.kind == ProofObject::MerkleRoot) fiat_shamir.absorb(o.hashes[0].b, 32);
      proof_stream.push(std::move(o));
    }
    return std::vector<size_t>(top.begin(), top.begin() + num_colinearity_tests);
  }
  // fri.rs:313-505 on the device (stark_fri_verify).  On success the stream objects the reference pops are consumed, the
  // roots are absorbed into fiat_shamir (fri.rs:327) and polynomial_values receives the top-layer pairs (fri.rs:437-441);
  // on failure the reference's println! line is printed and false returned.
  bool verify(ProofStream &proof_stream, FiatShamir &fiat_shamir,
              std::vector<std::pair<size_t, FieldElement>> &polynomial_values) const {
    const std::vector<uint8_t> bytes = proof_stream.serialize();
    const size_t nq = num_colinearity_tests, R = (size_t)num_rounds();
    int ok = 0;
    uint32_t why = 0;
    std::vector<uint8_t> roots(32 * (R ? R : 1));
    std::vector<uint64_t> top(nq ? nq : 1), pi(2 * nq + 1), pv(2 * nq + 1);
    check(stark_fri_verify(ctx(), bytes.data(), bytes.size(), domain_length, offset.value, omega.value,
                           (uint32_t)expansion_factor, (uint32_t)nq, fiat_shamir.transcript.data(),
                           fiat_shamir.transcript.size(), &ok, &why, roots.data(), top.data(), pi.data(), pv.data()));
    if (!ok) {
      printf("%s\n", stark_fri_verify_reason(why));
      return false;
    }
    for (size_t r = 0; r < R; r++) fiat_shamir.absorb(&roots[32 * r], 32);
    for (size_t i = 0; R > 1 && i < 2 * nq; i++) polynomial_values.push_back({(size_t)pi[i], field.new_element(pv[i])});
    const size_t consumed = R + 1 + (R - 1) * 4 * nq;
    proof_stream.objects.erase(proof_stream.objects.begin(), proof_stream.objects.begin() + consumed);
    return true;
  }
};

// trace.rs:21-34 columns -> low-degree extension, column-major
inline std::vector<std::vector<uint64_t>> lde(const std::vector<std::vector<uint64_t>> &cols, uint32_t log_blowup,
                                              uint64_t offset) {
  const size_t n = cols.at(0).size(), N = n << log_blowup;
  std::vector<uint64_t> flat, out(N * cols.size());
  for (const auto &c : cols) flat.insert(flat.end(), c.begin(), c.end());
  check(stark_lde(ctx(), flat.data(), (uint32_t)cols.size(), (uint32_t)__builtin_ctzll(n), log_blowup, offset, out.data()));
  std::vector<std::vector<uint64_t>> r;
  for (size_t c = 0; c < cols.size(); c++) r.emplace_back(out.begin() + c * N, out.begin() + (c + 1) * N);
  return r;
}

// trace.rs:4-49: the trace container (row-major i128) with ingestion + prove on the device
struct Trace {
  std::vector<std::vector<__int128>> trace;
  size_t num_columns;
  explicit Trace(const std::vector<std::vector<__int128>> &rows) : trace(rows), num_columns(rows.at(0).size()) {}
  const std::vector<__int128> *get_row(size_t i) const { return i < trace.size() ? &trace[i] : nullptr; }
  std::vector<__int128> get_col(size_t j) const {
    std::vector<__int128> c;
    for (const auto &r : trace) c.push_back(r.at(j));
    return c;
  }
  std::vector<std::vector<FieldElement>> to_field_elements(FiniteField field) const {   // `e as u64`, not reduced
    std::vector<std::vector<FieldElement>> out;
    for (const auto &r : trace) {
      out.emplace_back();
      for (__int128 e : r) out.back().push_back(field.new_element((uint64_t)e));
    }
    return out;
  }
  static Trace fibonacci(size_t length) {
    std::vector<std::vector<__int128>> rows;
    __int128 a = 1, b = 1;
    for (size_t i = 0; i < length; i++) {
      rows.push_back({a});
      __int128 next;
      if (__builtin_add_overflow(a, b, &next)) throw Panic("attempt to add with overflow");
      a = b, b = next;
    }
    return Trace(rows);
  }
  // LDE + per-column Merkle roots + Fri::prove of column 0 from the rows as they lie: stark_prove_trace_rows
  std::pair<std::vector<Hash>, std::vector<uint8_t>> prove(uint32_t log_blowup, uint64_t offset, uint32_t num_colinearity_tests) const {
    const size_t n = trace.size();
    if (n == 0 || (n & (n - 1))) throw Panic("n must be a power of two");
    std::vector<__int128> flat;
    for (const auto &r : trace) {
      if (r.size() != num_columns) throw Panic("ragged trace");
      flat.insert(flat.end(), r.begin(), r.end());
    }
    size_t cap = 0, len = 0;
    check(stark_fri_proof_size(n << log_blowup, 1u << log_blowup, num_colinearity_tests, &cap));
    std::vector<Hash> roots(num_columns);
    std::vector<uint8_t> proof(cap);
    check(stark_prove_trace_rows(ctx(), flat.data(), (uint32_t)num_columns, (uint32_t)__builtin_ctzll(n), log_blowup, offset,
                                 num_colinearity_tests, roots.data()->b, proof.data(), cap, &len));
    proof.resize(len);
    return {roots, proof};
  }
};

}  // namespace stark
