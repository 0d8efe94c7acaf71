// mgpu.cu -- groups of GPUs behind the C ABI (the reference has no parallel path at all): the
// library itself owns the NCCL communicator and the peer-memory windows, so a host in any language reaches the sharded
// prover through plain extern "C" calls (stark_mgpu_*).  See mgpu.h for the exchange protocol.
//
//   stark_mgpu_init          one process (or thread) per GPU: ncclCommInitRank with a caller-distributed unique id; the
//                            window handles are exchanged with ncclAllGather and mapped with CUDA IPC
//   stark_mgpu_create_local  one host thread driving several contexts: distinct devices with peer access, or several
//                            virtual ranks on ONE device (the 1-GPU tests), which run in lock step
// NCCL is loaded with dlopen (libnccl.so.2 -- inside a PyTorch process that is the copy torch already loaded), so the
// library has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>

#include <vector>

#include "common.cuh"
#include "merkle.h"
#include "merkle_dev.cuh"
#include "mgpu.h"

// ------------------------------------------------------------------------------------------------ NCCL
namespace {
struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
};
NcclApi g_nccl;
int nccl_load() {
  if (g_nccl.h) return STARK_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names)
    if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
  if (!h) return stark_fail(nullptr, STARK_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                                       \
  *(void **)(&g_nccl.field) = dlsym(h, name);                                                  \
  if (!g_nccl.field) return stark_fail(nullptr, STARK_ERR_NCCL, "libnccl has no symbol %s", name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllGather, "ncclAllGather")
  SYM(GetErrorString, "ncclGetErrorString")
  SYM(GetVersion, "ncclGetVersion")
#undef SYM
  g_nccl.h = h;
  return STARK_OK;
}
}  // namespace
#define NCCL_TRY(ctx, expr)                                                                                \
  do {                                                                                                     \
    ncclResult_t r__ = (expr);                                                                             \
    if (r__ != ncclSuccess)                                                                                \
      return stark_fail((ctx), STARK_ERR_NCCL, "%s failed: %s", #expr, g_nccl.GetErrorString(r__));        \
  } while (0)

// ---------------------------------------------------------------------------------------------- kernels

// all-to-all barrier between the ranks, on the device: raise flags[kind][rank] = epoch in every peer's window, wait until
// every peer's flag has arrived in this rank's window.  mode: MG_X_FUSED both, MG_X_SIGNAL / MG_X_WAIT one half.
struct BarrierArgs {
  int world, rank, mode;
  u32 epoch;
  u32 *flag_peer[MG_MAX_RANKS];   // peer g's flags[kind]
  const u32 *flag_local;
  u32 *err_local;
};
__global__ void k_mg_barrier(const __grid_constant__ BarrierArgs A) {
  pdl_entry();
  const u32 t = threadIdx.x;
  if (t >= (u32)A.world) return;
  if (A.mode != MG_X_WAIT) {
    __threadfence_system();
    mg_st_release_sys(A.flag_peer[t] + A.rank, A.epoch);
  }
  if (A.mode != MG_X_SIGNAL) mg_wait_flag(A.flag_local + t, A.epoch, A.err_local);
}

// column broadcast (mg_bcast_column): 128-bit loads, one 128-bit store per peer; the word after the column carries the
// root's canonical-input flag (api.cu k_narrow) so that every rank rejects a non-canonical column, as the root does
struct BcastArgs {
  int world, rank;
  u32 *dst[MG_MAX_RANKS];
  const u32 *src;
  const u32 *flag_src;
  size_t n;     // multiple of 4
};
__global__ void __launch_bounds__(256) k_mg_bcast(const __grid_constant__ BcastArgs A) {
  pdl_entry();
  const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < A.n; i += stride) {
    const uint4 v = *reinterpret_cast<const uint4 *>(A.src + i);
#pragma unroll 1
    for (int g = 0; g < A.world; g++) *reinterpret_cast<uint4 *>(A.dst[g] + i) = v;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const u32 f = A.flag_src ? *A.flag_src : 0u;
    for (int g = 0; g < A.world; g++) A.dst[g][A.n] = f;
  }
  __threadfence_system();
}
// one rank raises flags[kind][rank] on every peer / every rank waits for ONE rank's flag
struct SignalArgs {
  int world, rank, from;
  u32 epoch;
  u32 *flag_peer[MG_MAX_RANKS];
  const u32 *flag_local;
  u32 *err_local;
};
__global__ void k_mg_signal_from(const __grid_constant__ SignalArgs A) {
  pdl_entry();
  const u32 t = threadIdx.x;
  if (A.rank == A.from) {
    __threadfence_system();
    if (t < (u32)A.world) mg_st_release_sys(A.flag_peer[t] + A.rank, A.epoch);
  } else if (t == 0) {
    mg_wait_flag(A.flag_local + A.from, A.epoch, A.err_local);
  }
}

// entries of `src` (32 bytes each) into slot idx[i] of every rank's column-root table
struct PutRootsArgs {
  int world;
  u32 n;
  u32 idx[64];
  u8 *dst[MG_MAX_RANKS];
};
__global__ void k_mg_put_roots(const u8 *__restrict__ src, const __grid_constant__ PutRootsArgs A) {
  pdl_entry();
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (entry, rank, word)
  const u32 w = t & 7u, g = (t >> 3) % (u32)A.world, i = (t >> 3) / (u32)A.world;
  if (i >= A.n) return;
  reinterpret_cast<u32 *>(A.dst[g] + 32 * (size_t)A.idx[i])[w] = reinterpret_cast<const u32 *>(src + 32 * (size_t)i)[w];
  __threadfence_system();
}

// ------------------------------------------------------------------------------------------ host helpers

static BarrierArgs barrier_args(stark_mgpu *m, int kind, u32 epoch, int mode) {
  BarrierArgs A;
  memset(&A, 0, sizeof A);
  A.world = m->world, A.rank = m->rank, A.mode = mode, A.epoch = epoch;
  for (int g = 0; g < m->world; g++) A.flag_peer[g] = mg_flags(m, g, kind);
  A.flag_local = mg_flags(m, m->rank, kind);
  A.err_local = reinterpret_cast<u32 *>(m->win + m->L.err);
  return A;
}
int mg_barrier_signal(stark_mgpu *m, int kind, u32 epoch) {
  LAUNCH_PDL(m->ctx, "mg_barrier", 0, k_mg_barrier, 1u, 32, barrier_args(m, kind, epoch, MG_X_SIGNAL));
  return STARK_OK;
}
int mg_barrier_wait(stark_mgpu *m, int kind, u32 epoch) {
  LAUNCH_PDL(m->ctx, "mg_barrier", 0, k_mg_barrier, 1u, 32, barrier_args(m, kind, epoch, MG_X_WAIT));
  return STARK_OK;
}
int mg_barrier(stark_mgpu *m, int kind, u32 epoch) {
  if (m->lockstep) return stark_fail(m->ctx, STARK_ERR_ARG, "a lock-step group signals and waits in separate phases");
  LAUNCH_PDL(m->ctx, "mg_barrier", 0, k_mg_barrier, 1u, 32, barrier_args(m, kind, epoch, MG_X_FUSED));
  return STARK_OK;
}
int mg_check_err(stark_mgpu *m) {
  u32 e = 0;
  CU_TRY(m->ctx, cudaMemcpyAsync(&e, m->win + m->L.err, 4, cudaMemcpyDeviceToHost, m->ctx->stream));
  CU_TRY(m->ctx, cudaStreamSynchronize(m->ctx->stream));
  if (e) {
    cudaMemsetAsync(m->win + m->L.err, 0, 4, m->ctx->stream);
    return stark_fail(m->ctx, STARK_ERR_NCCL, "rank %d: a peer did not arrive within %.0f s (exchange timed out)", m->rank,
                      MG_TIMEOUT_NS * 1e-9);
  }
  return STARK_OK;
}
int mg_all_gather(stark_mgpu *m, const void *send_dev, void *recv_dev, size_t bytes) {
  if (m->mode != MG_PROC || !m->nccl) return stark_fail(m->ctx, STARK_ERR_ARG, "no NCCL communicator in this group");
  NCCL_TRY(m->ctx, g_nccl.AllGather(send_dev, recv_dev, bytes, ncclUint8, (ncclComm_t)m->nccl, m->ctx->stream));
  m->bytes_sent += bytes * (size_t)(m->world - 1);
  return STARK_OK;
}
int mg_bcast_column(stark_mgpu *m, int root, const u32 *src, size_t n, u32 epoch) {
  if (n % 4 || n > m->L.arena_elems / 2) return stark_fail(m->ctx, STARK_ERR_ARG, "column does not fit the group's window");
  if (m->rank == root) {
    BcastArgs A;
    memset(&A, 0, sizeof A);
    A.world = m->world, A.rank = m->rank, A.src = src, A.n = n, A.flag_src = m->ctx->flag + 2;
    for (int g = 0; g < m->world; g++) A.dst[g] = mg_bcast_ptr(m, g);
    size_t blocks = (n / 4 + 255) / 256, cap = (size_t)m->ctx->sm_count * 4;
    LAUNCH_PDL(m->ctx, "mg_bcast", 4ull * n * m->world, k_mg_bcast, (u32)(blocks < cap ? blocks : cap), 256, A);
    m->bytes_sent += 4ull * n * (size_t)(m->world - 1);
  }
  SignalArgs S;
  memset(&S, 0, sizeof S);
  S.world = m->world, S.rank = m->rank, S.from = root, S.epoch = epoch;
  for (int g = 0; g < m->world; g++) S.flag_peer[g] = mg_flags(m, g, 2);
  S.flag_local = mg_flags(m, m->rank, 2), S.err_local = reinterpret_cast<u32 *>(m->win + m->L.err);
  LAUNCH_PDL(m->ctx, "mg_signal", 0, k_mg_signal_from, 1u, 32, S);
  return STARK_OK;
}
int mg_put_roots(stark_mgpu *m, const u8 *src_dev, const u32 *idx, u32 n) {
  for (u32 i0 = 0; i0 < n; i0 += 64) {
    PutRootsArgs A;
    memset(&A, 0, sizeof A);
    A.world = m->world, A.n = n - i0 < 64 ? n - i0 : 64;
    for (u32 i = 0; i < A.n; i++) {
      if (idx[i0 + i] >= m->L.max_cols / 2) return stark_fail(m->ctx, STARK_ERR_ARG, "more roots than the window holds");
      A.idx[i] = idx[i0 + i];
    }
    for (int g = 0; g < m->world; g++) A.dst[g] = mg_colroots(m, g);
    const u32 threads = A.n * (u32)m->world * 8;
    LAUNCH_PDL(m->ctx, "mg_put_roots", 0, k_mg_put_roots, (threads + 127) / 128, 128, src_dev + 32 * (size_t)i0, A);
    m->bytes_sent += 32ull * A.n * (size_t)(m->world - 1);
  }
  return STARK_OK;
}

static void layout(MgLayout *L, size_t max_codeword) {
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t off = 0;
  L->flags = off, off = up(off + 4 * MG_FLAG_KINDS * MG_MAX_RANKS);
  L->err = off, off = up(off + 4);
  L->slots = off, off = up(off + 32 * (size_t)MG_MAX_ROUNDS * MG_MAX_RANKS);
  L->max_cols = 1024;
  L->colroots = off, off = up(off + 32 * L->max_cols);
  // proof of the largest codeword the window serves: 32 queries x 3 paths x 24 levels x 24 rounds is ~1.8 MB; sized from
  // the arena so that small groups stay small.  fri_proof_size is checked against it at prove time.
  L->proof_cap = (4u << 20);
  L->proof = off, off = up(off + L->proof_cap);
  L->arena_elems = max_codeword < 1024 ? 1024 : max_codeword;
  L->arena = off, off = up(off + 4 * L->arena_elems);
  L->bcast = off, off = up(off + 4 * (L->arena_elems / 2) + 256);   // a trace column (<= max_codeword / 2 rows) + flag word
  L->total = off;
}

static int alloc_window(stark_mgpu *m, size_t max_codeword) {
  layout(&m->L, max_codeword);
  CU_TRY(m->ctx, cudaSetDevice(m->ctx->device));
  CU_TRY(m->ctx, cudaMalloc((void **)&m->win, m->L.total));
  CU_TRY(m->ctx, cudaMemset(m->win, 0, m->L.arena));   // flags, slots, tables (the arena needs no initial value)
  CU_TRY(m->ctx, cudaDeviceSynchronize());
  return STARK_OK;
}

static stark_mgpu *new_handle(stark_ctx *ctx, int rank, int world, int mode) {
  stark_mgpu *m = new stark_mgpu();
  memset((void *)m, 0, sizeof *m);
  m->ctx = ctx, m->rank = rank, m->world = world, m->mode = mode;
  m->shard_log = 17;
  if (const char *e = getenv("STARK_MGPU_SHARD_LOG")) {
    const int v = atoi(e);
    m->shard_log = v < 4 ? 4 : (v > 30 ? 30 : v);
  }
  return m;
}

static bool pow2_world(int w) { return w >= 1 && w <= MG_MAX_RANKS && (w & (w - 1)) == 0; }

// ----------------------------------------------------------------------------------------------- C ABI
extern "C" {

int stark_mgpu_unique_id(uint8_t id[STARK_MGPU_ID_BYTES]) {
  if (!id) return stark_fail(nullptr, STARK_ERR_ARG, "null argument");
  ST_TRY(nccl_load());
  static_assert(sizeof(ncclUniqueId) == STARK_MGPU_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId u;
  NCCL_TRY(nullptr, g_nccl.GetUniqueId(&u));
  memcpy(id, &u, sizeof u);
  return STARK_OK;
}

struct PeerRecord {
  cudaIpcMemHandle_t handle;   // 64 bytes
  u64 pid, ptr;
  int device, pad;
};

int stark_mgpu_init(stark_ctx *ctx, const uint8_t id[STARK_MGPU_ID_BYTES], int rank, int world, size_t max_codeword,
                    stark_mgpu **out) {
  MgDeviceScope restore_device;
  if (!ctx || !id || !out) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (!pow2_world(world) || rank < 0 || rank >= world)
    return stark_fail(ctx, STARK_ERR_ARG, "world size must be 1, 2, 4 or 8 and 0 <= rank < world");
  ST_TRY(nccl_load());
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  stark_mgpu *m = new_handle(ctx, rank, world, MG_PROC);
  int rc = alloc_window(m, max_codeword);
  ncclComm_t comm = nullptr;
  PeerRecord *d_rec = nullptr;
  std::vector<PeerRecord> rec(world);
  if (rc == STARK_OK) {
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclResult_t r = g_nccl.CommInitRank(&comm, world, u, rank);
    if (r != ncclSuccess) rc = stark_fail(ctx, STARK_ERR_NCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
    m->nccl = comm;
  }
  // exchange the window handles: one ncclAllGather of a 96-byte record per rank
  if (rc == STARK_OK) {
    PeerRecord mine;
    memset(&mine, 0, sizeof mine);
    mine.pid = (u64)getpid(), mine.ptr = (u64)(uintptr_t)m->win, mine.device = ctx->device;
    if (cudaIpcGetMemHandle(&mine.handle, m->win) != cudaSuccess) {
      cudaGetLastError();
      memset(&mine.handle, 0, sizeof mine.handle);   // same-process peers do not need it; others will fail to open it
    }
    if (cudaMalloc((void **)&d_rec, sizeof(PeerRecord) * (size_t)(world + 1)) != cudaSuccess)
      rc = stark_fail(ctx, STARK_ERR_OOM, "window exchange buffer");
    if (rc == STARK_OK && cudaMemcpyAsync(d_rec + world, &mine, sizeof mine, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
      rc = stark_fail(ctx, STARK_ERR_CUDA, "H2D copy failed");
    if (rc == STARK_OK) rc = mg_all_gather(m, d_rec + world, d_rec, sizeof(PeerRecord));
    if (rc == STARK_OK && (cudaMemcpyAsync(rec.data(), d_rec, sizeof(PeerRecord) * (size_t)world, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                           cudaStreamSynchronize(ctx->stream) != cudaSuccess))
      rc = stark_fail(ctx, STARK_ERR_CUDA, "window exchange failed: %s", cudaGetErrorString(cudaGetLastError()));
    m->bytes_sent = 0;   // the bootstrap is not data-path traffic
  }
  for (int g = 0; g < world && rc == STARK_OK; g++) {
    if (g == rank) {
      m->peer[g] = m->win;
    } else if (rec[g].pid == (u64)getpid()) {
      // another rank of this process (one thread per GPU): address its allocation directly
      if (rec[g].device != ctx->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(rec[g].device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          rc = stark_fail(ctx, STARK_ERR_CUDA, "no peer access from device %d to device %d: %s", ctx->device, rec[g].device, cudaGetErrorString(e));
        cudaGetLastError();
      }
      m->peer[g] = (u8 *)(uintptr_t)rec[g].ptr;
    } else {
      void *p = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&p, rec[g].handle, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess)
        rc = stark_fail(ctx, STARK_ERR_CUDA, "cudaIpcOpenMemHandle for rank %d failed: %s", g, cudaGetErrorString(e));
      else
        m->peer[g] = (u8 *)p, m->ipc_opened[g] = true;
    }
  }
  if (d_rec) cudaFree(d_rec);
  // every rank has mapped every window before anyone uses it
  if (rc == STARK_OK) {
    mg_begin_op(m);
    rc = mg_barrier(m, 1, mg_epoch(m, 0));
    if (rc == STARK_OK) rc = mg_check_err(m);
  }
  if (rc != STARK_OK) {
    stark_mgpu_destroy(m);
    return rc;
  }
  *out = m;
  return STARK_OK;
}

int stark_mgpu_create_local(stark_ctx *const *ctxs, int world, size_t max_codeword, stark_mgpu **out) {
  MgDeviceScope restore_device;
  if (!ctxs || !out) return stark_fail(nullptr, STARK_ERR_ARG, "null argument");
  if (!pow2_world(world)) return stark_fail(nullptr, STARK_ERR_ARG, "world size must be 1, 2, 4 or 8");
  stark_mgpu **group = new stark_mgpu *[world];
  int rc = STARK_OK;
  bool same_device = false;
  for (int g = 0; g < world; g++) {
    group[g] = nullptr;
    if (!ctxs[g]) rc = stark_fail(nullptr, STARK_ERR_ARG, "null context");
    for (int k = 0; k < g && rc == STARK_OK; k++) same_device |= ctxs[k]->device == ctxs[g]->device;
  }
  for (int g = 0; g < world && rc == STARK_OK; g++) {
    group[g] = new_handle(ctxs[g], g, world, MG_LOCAL);
    group[g]->group = group;
    // ranks that share a device cannot wait for one another inside a kernel (nothing guarantees co-residency): the
    // driver runs them in lock step, signal phase on every rank first, then the wait phase
    group[g]->lockstep = same_device;
    rc = alloc_window(group[g], max_codeword);
  }
  for (int g = 0; g < world && rc == STARK_OK; g++) {
    cudaSetDevice(ctxs[g]->device);
    for (int k = 0; k < world && rc == STARK_OK; k++) {
      if (ctxs[k]->device != ctxs[g]->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[k]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          rc = stark_fail(ctxs[g], STARK_ERR_CUDA, "no peer access from device %d to device %d: %s", ctxs[g]->device,
                          ctxs[k]->device, cudaGetErrorString(e));
        cudaGetLastError();
      }
      group[g]->peer[k] = group[k]->win;
    }
  }
  if (rc != STARK_OK) {
    for (int g = 0; g < world; g++)
      if (group[g]) {
        if (group[g]->win) cudaFree(group[g]->win);
        delete group[g];
      }
    delete[] group;
    return rc;
  }
  for (int g = 0; g < world; g++) out[g] = group[g];
  return STARK_OK;
}

void stark_mgpu_destroy(stark_mgpu *m) {
  MgDeviceScope restore_device;
  if (!m) return;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  if (m->mode == MG_LOCAL) {
    // handles of a local group are destroyed together, through any one of them
    stark_mgpu **group = m->group;
    const int world = m->world;
    for (int g = 0; g < world; g++) {
      cudaSetDevice(group[g]->ctx->device);
      cudaStreamSynchronize(group[g]->ctx->stream);
    }
    for (int g = 0; g < world; g++) {
      cudaSetDevice(group[g]->ctx->device);
      if (group[g]->win) cudaFree(group[g]->win);
      delete group[g];
    }
    delete[] group;
    return;
  }
  for (int g = 0; g < m->world; g++)
    if (m->ipc_opened[g]) cudaIpcCloseMemHandle(m->peer[g]);
  if (m->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)m->nccl);
  if (m->win) cudaFree(m->win);
  delete m;
}

int stark_mgpu_rank(const stark_mgpu *m) { return m ? m->rank : -1; }
int stark_mgpu_world(const stark_mgpu *m) { return m ? m->world : 0; }
uint64_t stark_mgpu_bytes_sent(const stark_mgpu *m) { return m ? m->bytes_sent : 0; }
int stark_mgpu_set_shard_log(stark_mgpu *m, uint32_t log_n) {
  if (!m || log_n < 4 || log_n > 30) return stark_fail(m ? m->ctx : nullptr, STARK_ERR_ARG, "shard_log must be in 4..30");
  if (m->mode == MG_LOCAL)
    for (int g = 0; g < m->world; g++) m->group[g]->shard_log = (int)log_n;
  else
    m->shard_log = (int)log_n;
  return STARK_OK;
}

// device-side barrier over the group followed by a host synchronisation of this rank's stream
int stark_mgpu_barrier(stark_mgpu *m) {
  MgDeviceScope restore_device;
  if (!m) return stark_fail(nullptr, STARK_ERR_ARG, "null argument");
  if (m->mode == MG_LOCAL) {
    stark_mgpu **G = m->group;
    for (int g = 0; g < m->world; g++) mg_begin_op(G[g]);
    for (int g = 0; g < m->world; g++) {
      mg_use(G[g]);
      ST_TRY(mg_barrier_signal(G[g], 1, mg_epoch(G[g], 0)));
    }
    for (int g = 0; g < m->world; g++) {
      mg_use(G[g]);
      ST_TRY(mg_barrier_wait(G[g], 1, mg_epoch(G[g], 0)));
    }
    for (int g = 0; g < m->world; g++) {
      mg_use(G[g]);
      ST_TRY(mg_check_err(G[g]));
    }
    return STARK_OK;
  }
  mg_begin_op(m);
  ST_TRY(mg_barrier(m, 1, mg_epoch(m, 0)));
  return mg_check_err(m);
}

}  // extern "C"
