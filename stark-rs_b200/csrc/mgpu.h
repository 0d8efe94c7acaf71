// mgpu.h -- one rank's membership in a group of GPUs: the peer-memory WINDOW every rank exposes to the
// others over NVLink, the epoch flags that order the exchanges, and the NCCL communicator used for the bootstrap and
// for the exchanges that are not latency-critical.  Shared by mgpu.cu (setup, barriers, C ABI) and fri.cu / merkle.cu
// (the kernels that store into peers).
//
// Exchanges of the sharded prover (north star: "NCCL over NVLink used only to gather subtree roots and folded
// codewords"):
//   * subtree roots   -- the CTA that finishes a rank's subtree stores its 32-byte root into every peer's window and
//                        raises an epoch flag there, waits for the peers' flags, climbs the G roots to the tree root and
//                        runs the Fiat-Shamir round: one kernel, no host round trip (mg_exchange_top, merkle_dev.cuh)
//   * folded codeword -- the fold kernel stores its output slice into every peer's replica (fused fold + all-gather);
//                        the slice is complete on the peer before the NEXT root flag is raised, so it needs no flag of
//                        its own
//   * group / column roots of independent commitments -- ncclAllGather (multi-process groups) or the same peer stores
// A window is one cudaMalloc allocation: peers map it with CUDA IPC (one process per GPU) or address it directly (one
// process driving several GPUs with peer access enabled; or several "virtual ranks" on ONE device, which is how the
// 1-GPU tests exercise the sharded path in lock step).
#pragma once
#include "common.cuh"

constexpr int MG_MAX_RANKS = 8;
constexpr int MG_MAX_ROUNDS = 24;          // = MAX_FRI_ROUNDS
constexpr u32 MG_EPOCH_STRIDE = 64;        // epochs reserved per collective operation (> MG_MAX_ROUNDS + barriers)
constexpr int MG_FLAG_KINDS = 4;           // 0: round roots, 1: end-of-operation barrier, 2: auxiliary barrier, 3: spare

// window layout (byte offsets from the window base, identical on every rank)
struct MgLayout {
  size_t flags;      // u32 [MG_FLAG_KINDS][MG_MAX_RANKS]: flags[k][g] = last epoch rank g raised here
  size_t err;        // u32: set by a wait that timed out
  size_t slots;      // u8  [MG_MAX_ROUNDS][MG_MAX_RANKS][32]: subtree roots of the current operation
  size_t colroots;   // u8  [max_cols][32]: column / group roots of the current operation
  size_t proof;      // u8  [proof_cap]: ProofStream::serialize bytes, assembled by all ranks
  size_t arena;      // u32 [arena_elems]: replicas of the folded codewords
  size_t bcast;      // u32 [arena_elems / 2] + one flag word: a column one rank uploaded for everybody (mg_bcast_column)
  size_t total;
  size_t max_cols, proof_cap, arena_elems;
};

struct stark_mgpu {
  stark_ctx *ctx;
  int rank, world;
  int mode;                 // MG_PROC: one process (or thread) per rank, NCCL + CUDA IPC;  MG_LOCAL: handles created together
  void *nccl;               // ncclComm_t (MG_PROC only)
  u8 *win;                  // this rank's window (device memory of ctx->device)
  u8 *peer[MG_MAX_RANKS];   // every rank's window as addressable from THIS device (peer[rank] == win)
  bool ipc_opened[MG_MAX_RANKS];
  MgLayout L;
  u32 op;                   // collective operations started so far (same on every rank): epoch base = op * MG_EPOCH_STRIDE
  u64 bytes_sent;           // bytes this rank stored into peers or handed to NCCL (bench.py "comm")
  int shard_log;            // FRI rounds with at least 2^shard_log elements are sharded (STARK_MGPU_SHARD_LOG, default 17)
  bool lockstep;            // virtual ranks on one device: signal and wait are separate launches, driven in lock step
  stark_mgpu **group;       // MG_LOCAL: all handles of the group, indexed by rank (owned by rank 0's handle)
};
enum { MG_PROC = 0, MG_LOCAL = 1 };

// What the root-producing kernel of a sharded tree does with its subtree root (merkle_dev.cuh: mg_exchange_top).
struct MgExchange {
  int world, rank;          // world <= 1: no exchange (plain single-GPU tree)
  int mode;                 // MG_X_FUSED: signal, wait, top, transcript;  MG_X_SIGNAL: signal only;  MG_X_WAIT: wait, top, transcript
  u32 epoch;                // value raised in flags[0][rank] on every peer
  u8 *slot_peer[MG_MAX_RANKS];    // peer g's slot array of this round (MG_MAX_RANKS x 32 bytes); this rank writes entry `rank`
  u32 *flag_peer[MG_MAX_RANKS];   // peer g's flags[0]; this rank writes entry `rank`
  const u8 *slot_local;           // this rank's slot array of this round
  const u32 *flag_local;          // this rank's flags[0]
  u32 *err_local;
  u8 *top_nodes;                  // (2 world - 1) hashes: the replicated top tree over the world subtree roots
};
enum { MG_X_FUSED = 0, MG_X_SIGNAL = 1, MG_X_WAIT = 2 };

// ---- host helpers (mgpu.cu)
// one host thread may drive ranks on several devices: make the rank's device current before queueing its work
static inline void mg_use(const stark_mgpu *m) { cudaSetDevice(m->ctx->device); }
// Group calls walk over the ranks of a one-process group and make each rank's device current in turn: the device that was
// current when the caller entered is current again when it gets its answer (a caller that goes on with a single-GPU
// call on another context must not find itself on the last rank's device).
struct MgDeviceScope {
  int dev = -1;
  MgDeviceScope() { if (cudaGetDevice(&dev) != cudaSuccess) dev = -1; }
  MgDeviceScope(const MgDeviceScope &) = delete;
  MgDeviceScope &operator=(const MgDeviceScope &) = delete;
  ~MgDeviceScope() { if (dev >= 0) cudaSetDevice(dev); }
};
static inline u32 *mg_flags(const stark_mgpu *m, int g, int kind) {
  return reinterpret_cast<u32 *>(m->peer[g] + m->L.flags) + kind * MG_MAX_RANKS;
}
static inline u32 mg_epoch(const stark_mgpu *m, u32 k) { return m->op * MG_EPOCH_STRIDE + k + 1; }
// start a collective operation (every rank calls the same operations in the same order)
static inline void mg_begin_op(stark_mgpu *m) { m->op++; }
// all-to-all barrier on the stream: every rank raises flags[kind][rank] = epoch on every peer, then waits for all of them.
// In lock-step groups the two halves are separate calls (signal on every rank first, then wait on every rank).
int mg_barrier_signal(stark_mgpu *m, int kind, u32 epoch);
int mg_barrier_wait(stark_mgpu *m, int kind, u32 epoch);
int mg_barrier(stark_mgpu *m, int kind, u32 epoch);          // signal + wait in one launch (not for lock-step groups)
// after the final synchronisation of an operation: STARK_ERR_NCCL if a wait timed out on this rank
int mg_check_err(stark_mgpu *m);
// ncclAllGather of `bytes` per rank (device buffers) on the context's stream; MG_LOCAL groups copy peer to peer instead
int mg_all_gather(stark_mgpu *m, const void *send_dev, void *recv_dev, size_t bytes);
// Column broadcast: rank `root` stores n u32 values (src, on its device) and its canonical-input flag into every rank's
// bcast region and raises flags[2]; the other ranks wait for that flag.  PCIe is the scarce link when every rank copies
// from the host at once: a column all ranks need is uploaded by ONE of them and travels on over NVLink.
int mg_bcast_column(stark_mgpu *m, int root, const u32 *src, size_t n, u32 epoch);
static inline u32 *mg_bcast_ptr(const stark_mgpu *m, int g) { return reinterpret_cast<u32 *>(m->peer[g] + m->L.bcast); }
// this rank's n roots (device, 32 bytes each) into entries idx[] of EVERY rank's column-root table (peer stores).  The
// table is double-buffered by operation parity: a rank may be one operation ahead of a peer that is still copying the
// previous table to its host.
int mg_put_roots(stark_mgpu *m, const u8 *src_dev, const u32 *idx, u32 n);
static inline u8 *mg_colroots(const stark_mgpu *m, int g) {
  return m->peer[g] + m->L.colroots + (m->op & 1u) * 32 * (m->L.max_cols / 2);
}
