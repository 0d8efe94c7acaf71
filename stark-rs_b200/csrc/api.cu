// api.cu -- context, device buffers, host<->device value marshalling and the ff.rs batch entry points.
//
// Host values are uint64_t (FieldElement.value, reference src/ff.rs:25-28); device values are uint32_t.
// upload_u64 narrows on the device and REJECTS non-canonical input (>= p) instead of reducing it: the
// reference hashes and serialises raw values (hash.rs:32-35, stream.rs:45-51), so silently reducing would
// break bit-exactness.
#include "common.cuh"
#include <stdlib.h>

#include "merkle.h"

thread_local char g_stark_err[512] = "";

// ------------------------------------------------------------------------------------------- kernels

__global__ void k_narrow(const u64 *__restrict__ in, u32 *__restrict__ out, size_t n, u32 *flag) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 v = in[i];
  if (v >= ff::P) atomicOr(flag, 1u);
  out[i] = (u32)v;
}
__global__ void k_widen(const u32 *__restrict__ in, u64 *__restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

// Trace ingestion (trace.rs:21-34): rows[r][c] is an i128 (16 bytes, little-endian); the reference casts it `as u64`
// (low 64 bits, no reduction) into a FieldElement whose value then only ever enters FiniteField::mul / add, which
// reduce (ff.rs:138-152), so the column handed to the LDE is the residue mod p.  32 x 32 tiles through shared memory:
// coalesced 16-byte reads along the row, coalesced 4-byte writes along the column.
__global__ void __launch_bounds__(256) k_trace_ingest(const uint4 *__restrict__ rows, size_t n_rows, u32 n_cols,
                                                      u32 *__restrict__ cols) {
  __shared__ u32 tile[32][33];
  const size_t r0 = (size_t)blockIdx.x * 32;
  const u32 c0 = blockIdx.y * 32;
  for (u32 i = threadIdx.y; i < 32; i += blockDim.y) {
    const size_t r = r0 + i;
    const u32 c = c0 + threadIdx.x;
    u32 v = 0;
    if (r < n_rows && c < n_cols) {
      const uint4 x = rows[r * n_cols + c];
      v = (u32)((((u64)x.y << 32) | x.x) % ff::P);   // i128 as u64, then the residue
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (u32 i = threadIdx.y; i < 32; i += blockDim.y) {
    const u32 c = c0 + i;
    const size_t r = r0 + threadIdx.x;
    if (r < n_rows && c < n_cols) cols[(size_t)c * n_rows + r] = tile[threadIdx.x][i];
  }
}

enum { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_NEG = 3, OP_POW = 4 };
// element-wise ff.rs:138-167, 200-213 on canonical u32
__global__ void k_ff_vec(int op, const u32 *__restrict__ a, const u32 *__restrict__ b, u64 e, u32 *__restrict__ out,
                         size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const u32 x = a[i];
    u32 r;
    switch (op) {
      case OP_ADD: r = ff::add(x, b[i]); break;
      case OP_SUB: r = ff::sub(x, b[i]); break;
      case OP_MUL: r = ff::mul(x, b[i]); break;
      case OP_NEG: r = ff::neg(x); break;
      default: r = ff::pow(x, e); break;
    }
    out[i] = r;
  }
}
// ff.rs:169-178 for a whole vector: Montgomery batch inversion, 8 elements per thread (one Fermat
// exponentiation per 8 inverses).  Sets *flag when an element is zero ("no inverse").
__global__ void k_ff_vec_inv(const u32 *__restrict__ a, u32 *__restrict__ out, size_t n, u32 *flag) {
  constexpr int K = 8;
  const size_t T = (size_t)gridDim.x * blockDim.x;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  u32 x[K], pre[K];
  u32 acc = ff::R1;  // Montgomery one
  bool zero = false;
#pragma unroll
  for (int j = 0; j < K; j++) {
    const size_t i = t + (size_t)j * T;
    u32 v = i < n ? a[i] : 1u;
    if (v == 0) zero = true, v = 1u;
    x[j] = ff::to_mont(v);
    pre[j] = acc;
    acc = ff::canon(ff::mont_mul(acc, x[j]));
  }
  if (zero) atomicOr(flag, 2u);
  u32 inv = ff::mont_pow(acc, (u64)ff::P - 2);  // (prod)^-1, Montgomery form
#pragma unroll
  for (int j = K - 1; j >= 0; j--) {
    const size_t i = t + (size_t)j * T;
    const u32 r = ff::canon(ff::mont_mul(inv, pre[j]));  // x[j]^-1 (Montgomery)
    inv = ff::canon(ff::mont_mul(inv, x[j]));
    if (i < n) out[i] = ff::from_mont(r);
  }
}

// --------------------------------------------------------------------------------------- profiling

void prof_begin(stark_ctx *ctx, const char *tag, u64 bytes) {
  if (ctx->prof_n == ctx->prof_cap) {
    ctx->prof_cap = ctx->prof_cap ? ctx->prof_cap * 2 : 1024;
    ctx->prof = (ProfRec *)realloc(ctx->prof, ctx->prof_cap * sizeof(ProfRec));
  }
  ProfRec &r = ctx->prof[ctx->prof_n];
  r.tag = tag, r.bytes = bytes;
  cudaEventCreate(&r.e0), cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, ctx->stream);
}
void prof_end(stark_ctx *ctx) { cudaEventRecord(ctx->prof[ctx->prof_n++].e1, ctx->stream); }

// ------------------------------------------------------------------------------------- marshalling

static int read_flag(stark_ctx *ctx, u32 *value) {
  CU_TRY(ctx, cudaMemcpyAsync(ctx->h_flag, ctx->flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *value = *ctx->h_flag;
  return STARK_OK;
}

int upload_u64(stark_ctx *ctx, const uint64_t *host, size_t n, u32 *dst) {
  if (n == 0) return STARK_OK;
  u64 *tmp = nullptr;
  Scratch sc(ctx);
  ST_TRY(sc.get(&tmp, n * 8));
  CU_TRY(ctx, cudaMemsetAsync(ctx->flag, 0, 4, ctx->stream));
  CU_TRY(ctx, cudaMemcpyAsync(tmp, host, n * 8, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(ctx, "narrow_u64", 12ull * n, k_narrow<<<(u32)((n + 255) / 256), 256, 0, ctx->stream>>>(tmp, dst, n, ctx->flag));
  u32 f = 0;
  ST_TRY(read_flag(ctx, &f));
  if (f) return stark_fail(ctx, STARK_ERR_ARG, "non-canonical field element (value >= p) in input");
  return STARK_OK;
}

// the same without the host round trip: the canonical-input flag is copied to pinned host memory behind the narrowing
// kernel and inspected by upload_u64_check() after the caller's final synchronisation (pipelines that end with a sync
// anyway, e.g. stark_prove_trace, save one host-device bubble)
int upload_flag_reset(stark_ctx *ctx) {   // once before a group of upload_u64_nosync calls
  ctx->h_flag[2] = 0;
  CU_TRY(ctx, cudaMemsetAsync(ctx->flag + 2, 0, 4, ctx->stream));
  return STARK_OK;
}
int upload_u64_nosync(stark_ctx *ctx, const uint64_t *host, size_t n, u32 *dst) {
  if (n == 0) return STARK_OK;
  u64 *tmp = nullptr;
  Scratch sc(ctx);
  ST_TRY(sc.get(&tmp, n * 8));
  CU_TRY(ctx, cudaMemcpyAsync(tmp, host, n * 8, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(ctx, "narrow_u64", 12ull * n, k_narrow<<<(u32)((n + 255) / 256), 256, 0, ctx->stream>>>(tmp, dst, n, ctx->flag + 2));
  CU_TRY(ctx, cudaMemcpyAsync(ctx->h_flag + 2, ctx->flag + 2, 4, cudaMemcpyDeviceToHost, ctx->stream));
  return STARK_OK;
}
// pipelines that copy the host values themselves (column groups on a copy stream, fri.cu): narrowing of values already
// on the device, on the context's current stream, and the flag copy queued once at the end
int narrow_dev(stark_ctx *ctx, const uint64_t *staging_dev, size_t n, u32 *dst) {
  if (n == 0) return STARK_OK;
  LAUNCH(ctx, "narrow_u64", 12ull * n, k_narrow<<<(u32)((n + 255) / 256), 256, 0, ctx->stream>>>(staging_dev, dst, n, ctx->flag + 2));
  return STARK_OK;
}
int upload_flag_fetch(stark_ctx *ctx) {
  CU_TRY(ctx, cudaMemcpyAsync(ctx->h_flag + 2, ctx->flag + 2, 4, cudaMemcpyDeviceToHost, ctx->stream));
  return STARK_OK;
}
int upload_u64_check(stark_ctx *ctx) {   // call after the stream has been synchronised
  if (ctx->h_flag[2]) return stark_fail(ctx, STARK_ERR_ARG, "non-canonical field element (value >= p) in input");
  return STARK_OK;
}

int download_u64(stark_ctx *ctx, const u32 *src, size_t n, uint64_t *host) {
  if (n == 0) return STARK_OK;
  u64 *tmp = nullptr;
  Scratch sc(ctx);
  ST_TRY(sc.get(&tmp, n * 8));
  LAUNCH(ctx, "widen_u64", 12ull * n, k_widen<<<(u32)((n + 255) / 256), 256, 0, ctx->stream>>>(src, tmp, n));
  CU_TRY(ctx, cudaMemcpyAsync(host, tmp, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return STARK_OK;
}

// trace ingestion: H2D of the row-major i128 matrix + k_trace_ingest (no synchronisation)
int trace_to_columns_dev(stark_ctx *ctx, const void *rows_i128, size_t n_rows, u32 n_cols, u32 *cols) {
  const size_t bytes = n_rows * (size_t)n_cols * 16;
  if (bytes == 0) return STARK_OK;
  uint4 *tmp = nullptr;
  Scratch sc(ctx);
  ST_TRY(sc.get(&tmp, bytes));
  CU_TRY(ctx, cudaMemcpyAsync(tmp, rows_i128, bytes, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(ctx, "trace_ingest", bytes + 4 * n_rows * (size_t)n_cols,
         k_trace_ingest<<<dim3((u32)((n_rows + 31) / 32), (n_cols + 31) / 32), dim3(32, 8), 0, ctx->stream>>>(tmp, n_rows, n_cols, cols));
  return STARK_OK;
}

// ----------------------------------------------------------------------------------------------- C ABI

extern "C" {

const char *stark_last_error(void) { return g_stark_err; }
const char *stark_version(void) { return "stark_b200 0.1 (sm_100a)"; }

static int ctx_create(int device, cudaStream_t borrowed, bool borrow, stark_ctx **out) {
  if (!out) return stark_fail(nullptr, STARK_ERR_ARG, "null argument");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return stark_fail(nullptr, STARK_ERR_CUDA, "no CUDA device available (%s); libstark_b200 has no CPU fallback",
                      cudaGetErrorString(e));
  if (device < 0 || device >= count) return stark_fail(nullptr, STARK_ERR_ARG, "device %d out of range", device);
  CU_TRY(nullptr, cudaSetDevice(device));
  stark_ctx *ctx = new stark_ctx();
  memset(ctx, 0, sizeof *ctx);
  ctx->device = device;
  if (borrow) {
    ctx->stream = borrowed, ctx->own_stream = false;
  } else {
    cudaError_t se = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) {
      delete ctx;
      return stark_fail(nullptr, STARK_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(se));
    }
    ctx->own_stream = true;
  }
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  // keep freed scratch memory in the stream-ordered pool instead of returning it to the driver
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  int rc = STARK_OK;
  if (cudaMalloc(&ctx->flag, 4 * FLAG_WORDS) != cudaSuccess || cudaMallocHost(&ctx->h_flag, 16) != cudaSuccess)
    rc = stark_fail(nullptr, STARK_ERR_OOM, "context allocation failed");
  if (rc == STARK_OK && cudaMemset(ctx->flag, 0, 4 * FLAG_WORDS) != cudaSuccess) rc = stark_fail(nullptr, STARK_ERR_CUDA, "context initialisation failed");
  ctx->climb_counter = ctx->flag + TICKET_MAIN;
  // k_merkle_climb<256> takes at most 256 chunks of 512 / 1024 nodes (its fused top holds the chunk roots in 16 KB of
  // shared memory): the tuning knob is clamped to the range the kernel supports
  ctx->colpipe_serial = getenv("STARK_COLPIPE_SERIAL") && atoi(getenv("STARK_COLPIPE_SERIAL")) != 0;
  ctx->keep_pdl = getenv("STARK_KEEP_PDL") && atoi(getenv("STARK_KEEP_PDL")) != 0;
  ctx->trace_pipe = getenv("STARK_TRACE_PIPE") && atoi(getenv("STARK_TRACE_PIPE")) != 0;
  ctx->no_bcast0 = getenv("STARK_NO_BCAST0") && atoi(getenv("STARK_NO_BCAST0")) != 0;
  ctx->no_prio = getenv("STARK_NO_PRIO") && atoi(getenv("STARK_NO_PRIO")) != 0;
  ctx->colpipe_group = 8;
  if (const char *e = getenv("STARK_COLPIPE_GROUP")) {
    const int v = atoi(e);
    ctx->colpipe_group = v < 1 ? 1 : (v > 64 ? 64 : v);
  }
  ctx->climb_log = 18;
  if (const char *e = getenv("STARK_CLIMB_LOG")) {
    const int v = atoi(e);
    ctx->climb_log = v < 11 ? 11 : (v > 18 ? 18 : v);
  }
  if (rc == STARK_OK) rc = ntt_init(ctx);
  if (rc == STARK_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess)
    rc = stark_fail(nullptr, STARK_ERR_CUDA, "context initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (rc != STARK_OK) {
    delete ctx;
    return rc;
  }
  ctx->launches = 0;
  *out = ctx;
  return STARK_OK;
}
int stark_ctx_create(int device, stark_ctx **out) { return ctx_create(device, nullptr, false, out); }
int stark_ctx_create_on_stream(int device, void *cuda_stream, stark_ctx **out) {
  return ctx_create(device, (cudaStream_t)cuda_stream, true, out);
}
void stark_ctx_destroy(stark_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ntt_destroy(ctx);
  cudaFree(ctx->flag);
  cudaFreeHost(ctx->h_flag);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}
// A host thread that drives contexts on several devices makes a context's device current before using it (the group
// entry points do this themselves for every rank they drive).
int stark_ctx_make_current(stark_ctx *ctx) {
  if (!ctx) return stark_fail(nullptr, STARK_ERR_ARG, "null context");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  return STARK_OK;
}
int stark_ctx_sync(stark_ctx *ctx) {
  if (!ctx) return stark_fail(nullptr, STARK_ERR_ARG, "null context");
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return STARK_OK;
}
void *stark_ctx_stream(stark_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
uint64_t stark_ctx_launches(stark_ctx *ctx) { return ctx ? ctx->launches : 0; }

int stark_ctx_profile_begin(stark_ctx *ctx) {
  if (!ctx) return stark_fail(nullptr, STARK_ERR_ARG, "null context");
  ctx->prof_on = true;
  return STARK_OK;
}
// writes a JSON array [{"kernel": tag, "launches": n, "ms": total, "bytes": total algorithmic}, ...]
int stark_ctx_profile_end(stark_ctx *ctx, char *json, size_t cap) {
  if (!ctx || !json || cap < 3) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  ctx->prof_on = false;
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  struct Agg { const char *tag; u64 n, bytes; double ms; } agg[64];
  int na = 0;
  for (size_t i = 0; i < ctx->prof_n; i++) {
    ProfRec &r = ctx->prof[i];
    float ms = 0;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    cudaEventDestroy(r.e0), cudaEventDestroy(r.e1);
    int k = 0;
    while (k < na && strcmp(agg[k].tag, r.tag)) k++;
    if (k == na && na < 64) agg[na++] = Agg{r.tag, 0, 0, 0.0};
    if (k < 64) agg[k].n++, agg[k].bytes += r.bytes, agg[k].ms += ms;
  }
  ctx->prof_n = 0;
  size_t w = 0;
  w += snprintf(json + w, cap - w, "[");
  for (int k = 0; k < na && w + 160 < cap; k++)
    w += snprintf(json + w, cap - w, "%s{\"kernel\": \"%s\", \"launches\": %llu, \"ms\": %.6f, \"bytes\": %llu}", k ? ", " : "",
                  agg[k].tag, (unsigned long long)agg[k].n, agg[k].ms, (unsigned long long)agg[k].bytes);
  snprintf(json + w, cap - w, "]");
  return STARK_OK;
}

// ---- buffers
int stark_buf_alloc(stark_ctx *ctx, size_t n, stark_buf **out) {
  if (!ctx || !out) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  stark_buf *b = new stark_buf{ctx, nullptr, n, true};
  int rc = dev_alloc(ctx, (void **)&b->ptr, n * 4);
  if (rc != STARK_OK) {
    delete b;
    return rc;
  }
  *out = b;
  return STARK_OK;
}
int stark_buf_upload(stark_ctx *ctx, const uint64_t *host, size_t n, stark_buf **out) {
  if (!ctx || !out || (!host && n)) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  stark_buf *b = nullptr;
  ST_TRY(stark_buf_alloc(ctx, n, &b));
  int rc = upload_u64(ctx, host, n, b->ptr);
  if (rc != STARK_OK) {
    stark_buf_free(b);
    return rc;
  }
  *out = b;
  return STARK_OK;
}
int stark_buf_upload_into(stark_ctx *ctx, const uint64_t *host, size_t n, stark_buf *dst, size_t dst_off) {
  if (!ctx || !dst || (!host && n)) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (dst_off + n > dst->n) return stark_fail(ctx, STARK_ERR_ARG, "buffer too small");
  return upload_u64(ctx, host, n, dst->ptr + dst_off);
}
int stark_buf_download(stark_ctx *ctx, const stark_buf *buf, size_t off, size_t n, uint64_t *host) {
  if (!ctx || !buf || (!host && n)) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (off + n > buf->n) return stark_fail(ctx, STARK_ERR_ARG, "range out of bounds");
  return download_u64(ctx, buf->ptr + off, n, host);
}
int stark_buf_wrap(stark_ctx *ctx, void *device_u32, size_t n, stark_buf **out) {
  if (!ctx || !out || (!device_u32 && n)) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  // the kernels read and write field elements with 128-bit accesses (stark_b200.h, ALIGNMENT)
  if ((uintptr_t)device_u32 & 15u) return stark_fail(ctx, STARK_ERR_ARG, "device pointer must be 16-byte aligned");
  *out = new stark_buf{ctx, (u32 *)device_u32, n, false};
  return STARK_OK;
}
void *stark_buf_ptr(const stark_buf *buf) { return buf ? buf->ptr : nullptr; }
size_t stark_buf_len(const stark_buf *buf) { return buf ? buf->n : 0; }
void stark_buf_free(stark_buf *buf) {
  if (!buf) return;
  if (buf->owns && buf->ptr) dev_free(buf->ctx, buf->ptr);
  delete buf;
}

// ---- trace ingestion
int stark_trace_to_columns(stark_ctx *ctx, const void *rows_i128, size_t n_rows, uint32_t n_cols, stark_buf **out) {
  if (!ctx || !out || (!rows_i128 && n_rows && n_cols)) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  stark_buf *b = nullptr;
  ST_TRY(stark_buf_alloc(ctx, n_rows * (size_t)n_cols, &b));
  int rc = trace_to_columns_dev(ctx, rows_i128, n_rows, n_cols, b->ptr);
  if (rc == STARK_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = stark_fail(ctx, STARK_ERR_CUDA, "trace ingestion failed");
  if (rc != STARK_OK) {
    stark_buf_free(b);
    return rc;
  }
  *out = b;
  return STARK_OK;
}

// ---- ff.rs batch ops
static int ff_vec(stark_ctx *ctx, int op, const uint64_t *a, const uint64_t *b, u64 e, uint64_t *out, size_t n) {
  if (!ctx || (n && (!a || !out)) || (n && op <= OP_MUL && !b)) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  if (n == 0) return STARK_OK;
  u32 *da = nullptr, *db = nullptr, *dout = nullptr;
  Scratch sc(ctx);
  ST_TRY(sc.get(&da, n * 4));
  ST_TRY(sc.get(&dout, n * 4));
  int rc = upload_u64(ctx, a, n, da);
  if (rc == STARK_OK && op <= OP_MUL) {
    rc = sc.get(&db, n * 4);
    if (rc == STARK_OK) rc = upload_u64(ctx, b, n, db);
  }
  if (rc == STARK_OK) {
    if (op == 5) {
      cudaMemsetAsync(ctx->flag, 0, 4, ctx->stream);
      const size_t threads = (n + 7) / 8;
      k_ff_vec_inv<<<(u32)((threads + 127) / 128), 128, 0, ctx->stream>>>(da, dout, n, ctx->flag);
      ctx->launches++;
      u32 f = 0;
      rc = read_flag(ctx, &f);
      if (rc == STARK_OK && f) rc = stark_fail(ctx, STARK_ERR_ARG, "no inverse");  // ff.rs:171
    } else {
      const u32 blocks = (u32)((n + 255) / 256 < (size_t)ctx->sm_count * 16 ? (n + 255) / 256 : (size_t)ctx->sm_count * 16);
      k_ff_vec<<<blocks, 256, 0, ctx->stream>>>(op, da, db, e, dout, n);
      ctx->launches++;
    }
    if (rc == STARK_OK && cudaGetLastError() != cudaSuccess) rc = stark_fail(ctx, STARK_ERR_CUDA, "kernel launch failed");
  }
  if (rc == STARK_OK) rc = download_u64(ctx, dout, n, out);
  return rc;
}
int stark_ff_vec_add(stark_ctx *c, const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) { return ff_vec(c, OP_ADD, a, b, 0, o, n); }
int stark_ff_vec_sub(stark_ctx *c, const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) { return ff_vec(c, OP_SUB, a, b, 0, o, n); }
int stark_ff_vec_mul(stark_ctx *c, const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) { return ff_vec(c, OP_MUL, a, b, 0, o, n); }
int stark_ff_vec_neg(stark_ctx *c, const uint64_t *a, uint64_t *o, size_t n) { return ff_vec(c, OP_NEG, a, nullptr, 0, o, n); }
int stark_ff_vec_pow(stark_ctx *c, const uint64_t *a, uint64_t e, uint64_t *o, size_t n) { return ff_vec(c, OP_POW, a, nullptr, e, o, n); }
int stark_ff_vec_inv(stark_ctx *c, const uint64_t *a, uint64_t *o, size_t n) { return ff_vec(c, 5, a, nullptr, 0, o, n); }

int stark_ff_prim_nth_root(uint64_t n, uint64_t *out) {
  if (!out) return stark_fail(nullptr, STARK_ERR_ARG, "null argument");
  if (n == 0 || (n & (n - 1))) return stark_fail(nullptr, STARK_ERR_ARG, "n must be a power of two");       // ff.rs:217
  if (n > (1ull << 23)) return stark_fail(nullptr, STARK_ERR_ARG, "n > 2^23 not supported by this modulus");  // ff.rs:218
  *out = ff::pow(ff::GEN, (ff::P - 1) / n);
  return STARK_OK;
}

}  // extern "C"
