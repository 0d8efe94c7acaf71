// common.cuh -- context, status codes and error plumbing shared by the .cu files of libstark_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/stark_b200.h"
#include "field.cuh"
#include "ntt_core.cuh"

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;

struct GeoCacheEntry {
  u32 g, c;        // canonical base and constant
  u32 *lo, *hi;    // device tables (4096 + hi_len entries)
  u32 hi_len;
  u64 stamp;
  cudaEvent_t ready;   // recorded behind the kernel that fills the tables: a hit from another stream waits for it
};

struct stark_ctx {
  int device;
  cudaStream_t stream;
  bool own_stream;
  int sm_count;
  // w23 power tables (forward), Montgomery form
  u32 *root_lo, *root_hi;
  // per-length sub-transform twiddles: tw_sub[dir][(1 << logL) + e] = w_L^(+-e), logL <= 12
  u32 *tw_sub[2];
  // the same sub-transform twiddles, and the w_8 powers, in Shoup form (plain value + floor(w 2^32 / p)): what the
  // transforms multiply by (tw_sub is only their source)
  ntt::wpair *tw_sh[2];
  ntt::wpair w8_sh[2][4];
  ntt::wpair *tw_in_sh[2]; // inner twiddles per radix and round (ntt_pass.cuh fill_inner_twiddles)
  ntt::wpair *otw_sh[2];   // w_{2^16}^(+-e), e < 2^16 (outer twiddles of the MIDDLE pass)
  ntt::wpair *row12_sh[2]; // w_4096^(+-row), row < 128 (FIRST pass of the fused 2^12 kernel, ntt.cu k_ntt_small12)
  int ntt_small_off;       // STARK_NTT_SMALL_OFF=1: batches of 2^12 keep the one-CTA-per-transform kernel (comparison)
  ntt::wpair *row_sh[2];   // w_{2^logN}^(+-row), row < 2048, at [(logN - 13) * 2048 + row], logN = 13..23 (FIRST pass)
  GeoCacheEntry geo[8];
  u64 geo_stamp;
  u32 *flag;       // device ints: [0], [2] validation flags; [4 .. 4 + 64): last-CTA tickets of the Merkle climb kernel for
                   // launches on the context's stream, [68 .. 132) / [132 .. 196): the same for the two column streams
  u32 *climb_counter;   // the tickets the next climb launch uses (one per tree of a batch): flag + 4, or flag + 68
  u32 *h_flag;     // pinned host mirror
  u64 launches;    // kernels launched through this context (bench.py "gpu_launches")
  // side streams for batched transforms larger than L2 (ntt.cu: column groups run all passes back to back, a few groups
  // in flight): created on first use
  cudaStream_t side[6];   // [0], [1]: NTT column groups; [2], [4]: the two column streams of the config-3 pipeline; [3]: copies
  cudaEvent_t side_done[6], fork_ev;
  int n_side;            // streams created so far
  // the latency chain of a proof (column 0 -> LDE -> Fri::prove: ~50 short dependent kernels) runs on a stream of the
  // HIGHEST priority, so that its kernels are dispatched ahead of the pending CTAs of the throughput kernels the column
  // stream has in flight (fri.cu PrioScope); created on first use
  cudaStream_t prio_stream;
  cudaEvent_t prio_ev;
  int ntt_streams;       // groups in flight (STARK_NTT_STREAMS, default 2; 1 = whole batch per pass)
  int ntt_group_mb;      // bytes of one group's column slice (STARK_NTT_GROUP_MB, default 16)
  int ntt_big;           // STARK_NTT_BIG=1: two-pass plans on 16384-element tiles for 2^20..2^22 (experiment)
  int ntt_l2_persist;    // STARK_NTT_L2_PERSIST=1: mark each pass's destination as L2-persisting (experiment, default off)
  int l2_persist_ready;
  int keep_pdl;          // STARK_KEEP_PDL=1: programmatic dependent launch stays on inside multi-column pipelines (diagnosis)
  int trace_pipe;        // STARK_TRACE_PIPE=1: print device timestamps of the config-3 pipeline's milestones (diagnosis)
  int no_bcast0;         // STARK_NO_BCAST0=1: every rank of a group copies column 0 from the host itself (comparison)
  int no_prio;           // STARK_NO_PRIO=1: keep the latency chain on the caller's stream (diagnosis)
  int colpipe_serial;    // STARK_COLPIPE_SERIAL=1: no column / copy stream (everything on the context's stream; diagnosis)
  int colpipe_group;     // columns per group when the trace is copied from the host (STARK_COLPIPE_GROUP, default 8)
  int climb_log;         // Merkle levels above 2^climb_log nodes get one launch each, the rest one climb launch (merkle.cu)
  char err[512];
  // optional per-kernel timing (stark_ctx_profile_begin/end): CUDA events around every launch, on ctx->stream
  bool prof_on;
  struct ProfRec *prof;   // growing array
  size_t prof_n, prof_cap;
};
constexpr int FLAG_WORDS = 4 + 3 * 64, TICKET_MAIN = 4, TICKET_SIDE = 68, TICKET_SIDE2 = 132;   // layout of stark_ctx::flag
struct ProfRec {
  const char *tag;
  u64 bytes;              // algorithmic HBM bytes of this launch, 0 if not meaningful
  cudaEvent_t e0, e1;
};
void prof_begin(stark_ctx *ctx, const char *tag, u64 bytes);
void prof_end(stark_ctx *ctx);

extern thread_local char g_stark_err[512];

static inline int stark_fail(stark_ctx *ctx, int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_stark_err, sizeof g_stark_err, fmt, ap);
  va_end(ap);
  if (ctx) snprintf(ctx->err, sizeof ctx->err, "%s", g_stark_err);
  return code;
}

#define CU_TRY(ctx, expr)                                                                            \
  do {                                                                                               \
    cudaError_t e__ = (expr);                                                                        \
    if (e__ != cudaSuccess)                                                                          \
      return stark_fail((ctx), e__ == cudaErrorMemoryAllocation ? STARK_ERR_OOM : STARK_ERR_CUDA,    \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

#define ST_TRY(expr)            \
  do {                          \
    int s__ = (expr);           \
    if (s__ != STARK_OK) return s__; \
  } while (0)

#define KERNEL_CHECK(ctx)                \
  do {                                   \
    (ctx)->launches++;                   \
    CU_TRY((ctx), cudaGetLastError());   \
  } while (0)

// LAUNCH(ctx, "tag", algorithmic_bytes, kernel<<<grid, block, smem, ctx->stream>>>(args...));
#define LAUNCH(ctx, tag, bytes, ...)                       \
  do {                                                     \
    if ((ctx)->prof_on) prof_begin((ctx), (tag), (bytes)); \
    __VA_ARGS__;                                           \
    if ((ctx)->prof_on) prof_end((ctx));                   \
    KERNEL_CHECK(ctx);                                     \
  } while (0)

// ---- programmatic dependent launch (sm_90+): a kernel launched with LAUNCH_PDL may be SCHEDULED while its predecessor
// in the stream is still running; it must begin with pdl_entry(), which (1) lets ITS successor be scheduled early and
// (2) blocks until every predecessor grid has completed and its memory is visible.  The prove pipeline is a chain of
// ~50 short dependent kernels, so the ~2 us launch + ramp between two of them is worth hiding.  Kernels launched the
// classic way before or after a PDL kernel keep full stream ordering.
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_entry() {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
// While a multi-column pipeline is in flight the chain's kernels are launched the classic way: a dependent launched early
// sits in griddepcontrol.wait and holds CTA slots for as long as its predecessor runs -- behind a latency-bound climb
// kernel (~40 us, a handful of busy SMs) that idles the slots the column stream's throughput kernels would have used
// (measured: 16-column prove 6.9 -> 8.2 ms once the chain also had stream priority).  Set by PrioScope (fri.cu).
inline thread_local bool t_pdl_off = false;
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at, cfg.numAttrs = t_pdl_off ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif
// LAUNCH_PDL(ctx, "tag", algorithmic_bytes, kernel, grid, block, args...)
#define LAUNCH_PDL(ctx, tag, bytes, kernel, grid, block, ...)                                      \
  do {                                                                                             \
    if ((ctx)->prof_on) prof_begin((ctx), (tag), (bytes));                                         \
    cudaError_t le__ = launch_pdl(kernel, dim3(grid), dim3(block), 0, (ctx)->stream, __VA_ARGS__); \
    if ((ctx)->prof_on) prof_end((ctx));                                                           \
    (ctx)->launches++;                                                                             \
    CU_TRY((ctx), le__);                                                                           \
  } while (0)

// stream-ordered scratch allocation (no device-wide sync on the hot path)
static inline int dev_alloc(stark_ctx *ctx, void **p, size_t bytes) {
  CU_TRY(ctx, cudaMallocAsync(p, bytes ? bytes : 16, ctx->stream));
  return STARK_OK;
}
static inline void dev_free(stark_ctx *ctx, void *p) {
  if (p) cudaFreeAsync(p, ctx->stream);
}
// Scratch blocks owned by ONE call: the TRY / LAUNCH macros return from the middle of a function, so the blocks are
// handed back (stream-ordered, on the context's stream) by the destructor on every exit path.
struct Scratch {
  stark_ctx *ctx;
  void *blk[8];
  int n = 0;
  explicit Scratch(stark_ctx *c) : ctx(c) {}
  Scratch(const Scratch &) = delete;
  Scratch &operator=(const Scratch &) = delete;
  ~Scratch() {
    for (int i = 0; i < n; i++) dev_free(ctx, blk[i]);
  }
  template <class T>
  int get(T **out, size_t bytes) {
    if (n == 8) return stark_fail(ctx, STARK_ERR_CUDA, "scratch holder full");
    int rc = dev_alloc(ctx, (void **)out, bytes);
    if (rc == STARK_OK) blk[n++] = *out;
    return rc;
  }
};

// ---- internal device-pointer entry points (implemented in ntt.cu / poly.cu / merkle.cu / fri.cu)
struct ScaleSpec {
  int mode;   // ntt::ScaleMode
  u32 c;      // canonical constant (SCALE_CONST, SCALE_GEO)
  u32 g;      // canonical base (SCALE_GEO)
};
int ntt_init(stark_ctx *ctx);
int side_streams(stark_ctx *ctx, int n);   // ntt.cu: make sure ctx->side[0 .. n) and their events exist
void ntt_destroy(stark_ctx *ctx);
int geo_tables(stark_ctx *ctx, u32 g, u32 c, u64 max_index, ntt::GeoTables *out);
// batched transform of length 2^log_n: in/out device u32 (canonical), n_valid = leading input elements that are
// read (the rest are zero), batch strides in elements.  in == out is allowed.
int ntt_transform(stark_ctx *ctx, const u32 *in, u32 *out, int log_n, bool inverse, u32 batch, u64 in_batch,
                  u64 out_batch, u64 n_valid, ScaleSpec pre, ScaleSpec post);
