// merkle_dev.cuh -- device building blocks of the Merkle kernels (merkle.cu) and of the single-CTA FRI tail (fri.cu).
#pragma once
#include "hash.cuh"
#include "transcript.cuh"

__device__ __forceinline__ void load_hash(const uint8_t *p, uint32_t *w) {
  const uint4 a = reinterpret_cast<const uint4 *>(p)[0], b = reinterpret_cast<const uint4 *>(p)[1];
  w[0] = a.x, w[1] = a.y, w[2] = a.z, w[3] = a.w, w[4] = b.x, w[5] = b.y, w[6] = b.z, w[7] = b.w;
}
__device__ __forceinline__ void store_hash(uint8_t *p, const uint32_t *w) {
  reinterpret_cast<uint4 *>(p)[0] = make_uint4(w[0], w[1], w[2], w[3]);
  reinterpret_cast<uint4 *>(p)[1] = make_uint4(w[4], w[5], w[6], w[7]);
}
// hash offset of level l in the flattened tree of n leaves (MerkleTree.nodes, merkle.rs:18-29)
__device__ __forceinline__ size_t level_off(size_t n, uint32_t l) { return 2 * n - 2 * (n >> l); }

// One CTA of NT threads climbs up to `levels` levels of the tree (merkle.rs:21-27) starting from the `cnt` nodes
// (power of two, 2 <= cnt <= 4 NT) [first, first + cnt) of level `level_in`; every level is written to the tree array
// (MerkleTree::open needs them) and kept in shared memory for the next step.  Children come from `src0` in the first
// step (global memory, or shared memory when the caller already holds them there) and from `sm` afterwards.
//   parents >  NT/4:   two parents per thread (hs2, ALU-pipe efficient)
//   parents <= NT/4:   one parent per FOUR lanes (hsq): these steps are pure hash latency
// sm: at least 16 cnt bytes.  Returns with the last computed level in sm (thread 0 can read the root there).
// STARK_HSO=0 at compile time keeps the four-lane form everywhere (A/B measurement)
#ifndef STARK_HSO
#define STARK_HSO 1
#endif
#ifdef STARK_CLIMB_TIMING
__device__ unsigned long long g_climb_clk[64];
__device__ unsigned int g_climb_n;
#define CLIMB_TICK()                                                                       \
  do {                                                                                     \
    if (threadIdx.x == 0 && blockIdx.x == gridDim.x - 1 && g_climb_n < 64) g_climb_clk[g_climb_n++] = clock64(); \
  } while (0)
#else
#define CLIMB_TICK() do {} while (0)
#endif
template <int NT>
__device__ __forceinline__ void cta_climb(uint8_t *nodes, size_t n, uint32_t level_in, size_t first, uint32_t cnt,
                                          uint32_t levels, const uint8_t *src0, uint8_t *sm, uint32_t one) {
  const uint32_t t = threadIdx.x;
  uint32_t cur = cnt;
  const uint8_t *src = src0;
  CLIMB_TICK();
  for (uint32_t s = 0; s < levels && cur > 1; s++) {
    const uint32_t parents = cur >> 1;
    uint8_t *dst = nodes + 32 * (level_off(n, level_in + s + 1) + (first >> (s + 1)));
    if (parents <= 16 && 8 * parents <= (uint32_t)NT && STARK_HSO) {
      // octet path (hso): parent p on lanes 8p .. 8p+7, four output bytes per lane; whole warps without a parent skip.
      // Only while the live warps are at most one per scheduler (<= 128 lanes): a lone warp runs hso in 2180 cycles
      // against 2560 for hsq, but two hso warps sharing a scheduler are no faster than one hsq warp (measured: the
      // 64-parent step of the FRI tail got slower with it).
      const uint32_t p = t >> 3, q = t & 7u;
      const bool warp_on = (t & ~31u) < 8 * parents, act = p < parents;
      uint32_t o = 0;
      if (warp_on) {
        const uint32_t pc = act ? p : 0;   // idle octets of a live warp hash parent 0 again (shuffles need all lanes)
        const hso::Dev w;
        o = hso::combine(w, src + 64 * pc, src + 64 * pc + 32);
      }
      CLIMB_TICK();
      __syncthreads();
      if (act) {
        *reinterpret_cast<uint32_t *>(sm + 32 * p + 4 * q) = o;
        *reinterpret_cast<uint32_t *>(dst + 32 * p + 4 * q) = o;
      }
      __syncthreads();
    } else if (4 * parents <= (uint32_t)NT) {
      // quad path: parent p on lanes 4p .. 4p+3; whole warps without a parent skip (warp-uniform)
      const uint32_t p = t >> 2, q = t & 3u;
      const bool warp_on = (t & ~31u) < 4 * parents, act = p < parents;
      uint32_t o0 = 0, o1 = 0;
      if (warp_on) {
        const uint32_t pc = act ? p : 0;   // idle quads of a live warp hash parent 0 again (shuffles need all lanes)
        hsq::combine(src + 64 * pc, src + 64 * pc + 32, o0, o1);
      }
      CLIMB_TICK();
      __syncthreads();
      if (act) {
        *reinterpret_cast<uint2 *>(sm + 32 * p + 8 * q) = make_uint2(o0, o1);
        *reinterpret_cast<uint2 *>(dst + 32 * p + 8 * q) = make_uint2(o0, o1);
      }
      __syncthreads();
    } else {
      // two parents per thread (hs2): a lone warp hashes two nodes in about the time of one, so the one-per-thread
      // form never pays -- it only doubles the warps competing for an SM's issue slots
      uint32_t wa[8], wb[8];
      const uint32_t p = 2 * t;
      const bool act = p < parents, two = p + 1 < parents;
      if (act) {
        const uint32_t pb = two ? p + 1 : p;
        uint32_t la[8], ra[8], lb[8], rb[8];
        load_hash(src + 64 * p, la), load_hash(src + 64 * p + 32, ra);
        load_hash(src + 64 * pb, lb), load_hash(src + 64 * pb + 32, rb);
        hs2::combine2(la, ra, lb, rb, wa, wb, one);
      }
      __syncthreads();   // everybody has read its children: sm may be overwritten
      if (act) {
        store_hash(sm + 32 * p, wa), store_hash(dst + 32 * p, wa);
        if (two) store_hash(sm + 32 * p + 32, wb), store_hash(dst + 32 * p + 32, wb);
      }
      __syncthreads();
    }
    src = sm;
    cur = parents;
    CLIMB_TICK();
  }
}

// ------------------------------------------------------------------------------------ multi-GPU root exchange
// (mgpu.h)  Called by ALL threads of the CTA that holds a rank's 32-byte subtree root in sm[0..32): store it into every
// peer's window, raise this round's epoch flag there, wait until every peer's flag has arrived here, then climb the
// `world` roots to the tree root (the top log2(world) levels of MerkleTree::new, merkle.rs:21-27, replicated on every
// rank) and leave it in sm[0..32).  Returns false when this launch only signals (lock-step groups).
#include "mgpu.h"

__device__ __forceinline__ void mg_st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t mg_ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long mg_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until *flag has reached `epoch` (epochs only grow; compared modulo 2^32).  A peer that never arrives must not hang
// the GPU: after MG_TIMEOUT_NS the wait gives up and records the failure (the host returns STARK_ERR_NCCL).
constexpr unsigned long long MG_TIMEOUT_NS = 4000000000ull;
__device__ __forceinline__ void mg_wait_flag(const uint32_t *flag, uint32_t epoch, uint32_t *err) {
  if ((int32_t)(mg_ld_acquire_sys(flag) - epoch) >= 0) return;
  const unsigned long long t0 = mg_now_ns();
  while ((int32_t)(mg_ld_acquire_sys(flag) - epoch) < 0) {
    if (mg_now_ns() - t0 > MG_TIMEOUT_NS) {
      atomicOr(err, 1u);
      return;
    }
    __nanosleep(64);
  }
}

template <int NT>
__device__ __forceinline__ bool mg_exchange_top(const MgExchange &X, uint8_t *sm, uint32_t one) {
  const uint32_t t = threadIdx.x;
  const int G = X.world;
  if (X.mode != MG_X_WAIT) {
    // the root as 8 words to each of the G ranks (own window included: the top tree reads all roots from one place)
    if (t < 8u * G) {
      const uint32_t g = t >> 3, w = t & 7u;
      reinterpret_cast<uint32_t *>(X.slot_peer[g] + 32 * X.rank)[w] = reinterpret_cast<const uint32_t *>(sm)[w];
      __threadfence_system();
    }
    __syncthreads();
    if (t < (uint32_t)G) mg_st_release_sys(X.flag_peer[t] + X.rank, X.epoch);
    if (X.mode == MG_X_SIGNAL) return false;
  }
  if (t < (uint32_t)G) mg_wait_flag(X.flag_local + t, X.epoch, X.err_local);
  __syncthreads();
  // leaves of the top tree: the G roots, written by the peers into this rank's window (read past L1)
  if (t < 2u * G) {
    const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(X.slot_local) + t);
    reinterpret_cast<uint4 *>(sm)[t] = v;
    reinterpret_cast<uint4 *>(X.top_nodes)[t] = v;
  }
  __syncthreads();
  uint32_t levels = 0;
  for (int c = G; c > 1; c >>= 1) levels++;
  cta_climb<NT>(X.top_nodes, (size_t)G, 0, 0, (uint32_t)G, levels, sm, sm, one);
  return true;
}
