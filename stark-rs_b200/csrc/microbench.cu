// microbench.cu -- register-only integer issue-rate microbenchmark (measurement infrastructure).
// The roofline of the hash kernels is the integer pipe, for which MEASURED_PEAKS.json has no number
// (SURVEY 8(d)).  Three kernels with 8 independent chains per thread: IMAD only (fma pipe), LOP3 only
// (alu pipe), and a 1:1 mix (both pipes, the issue-slot limit).
#include "common.cuh"

// one dependent step of each flavour, pinned with inline PTX so that ptxas can neither fuse nor strength-reduce it
__device__ __forceinline__ u32 op_imad(u32 a, u32 m, u32 c) {
  u32 d;
  asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(m), "r"(c));
  return d;
}
__device__ __forceinline__ u32 op_lop3(u32 a, u32 m, u32 c) {
  u32 d;
  asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(m), "r"(c));
  return d;
}
__device__ __forceinline__ u32 op_iadd3(u32 a, u32 m, u32 c) {
  u32 d;
  asm volatile("{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3; }" : "=r"(d) : "r"(a), "r"(m), "r"(c));
  return d;
}
__device__ __forceinline__ u32 op_imad_hi(u32 a, u32 m, u32 c) {
  u32 d;
  asm volatile("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(m), "r"(c));
  return d;
}
// a 32 x 32 -> 64 multiply whose two halves are both consumed (by one LOP3 on the other pipe)
__device__ __forceinline__ u32 op_imad_wide(u32 a, u32 m) {
  u32 d;
  asm volatile("{ .reg .u64 w; .reg .u32 lo, hi; mul.wide.u32 w, %1, %2; mov.b64 {lo, hi}, w; xor.b32 %0, lo, hi; }" : "=r"(d) : "r"(a), "r"(m));
  return d;
}
// MODE 0: IMAD only (FMA pipe); MODE 1: LOP3 only (ALU pipe); MODE 2: IMAD + LOP3 alternating (both pipes)
// MODE 3: IMAD.WIDE.U32 (+ one LOP3 on the other pipe); MODE 4: IMAD.HI.U32; MODE 5: Montgomery products (field.cuh: IMAD.WIDE +
// IMAD + IMAD.HI each); MODE 6: the NTT butterfly mix -- one Montgomery product, one lazy add, one range reduction
template <int MODE>
__global__ void __launch_bounds__(256) k_int_peak(u32 *out, u32 seed, int iters) {
  u32 a[8];
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = seed + threadIdx.x * 8 + k;
  const u32 m = seed | 1u, c = seed ^ 0x9e3779b9u;
  const u32 tw = (seed * 2654435761u) % ff::P;   // a canonical twiddle for the Montgomery modes
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (MODE == 3) {
          a[k] = op_imad_wide(a[k], m);
        } else if (MODE == 4) {
          a[k] = op_imad_hi(a[k], m, c);
        } else if (MODE == 5) {
          a[k] = ff::mont_mul(a[k], tw);
        } else if (MODE == 6) {
          a[k] = ff::red2p(a[(k + 1) & 7] + ff::mont_mul(a[k], tw));
        } else if (MODE == 0) {
          a[k] = op_imad(a[k], m, c);
        } else if (MODE == 1) {
          a[k] = op_lop3(a[k], m, c);
        } else {
          a[k] = op_imad(a[k], m, c);
          a[k] = op_lop3(a[k], m, c);
        }
      }
    }
  }
  u32 r = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) r ^= a[k];
  if (r == 0x12345678u) out[0] = r;  // keep the chains alive
}

template <int MODE>
static int run_mode(stark_ctx *ctx, u32 *d_out, double ops_per_inner, double *per_s) {
  const int iters = 2048, blocks = ctx->sm_count * 8, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  k_int_peak<MODE><<<blocks, threads, 0, ctx->stream>>>(d_out, 12345u, 64);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0, ctx->stream);
    k_int_peak<MODE><<<blocks, threads, 0, ctx->stream>>>(d_out, 12345u + rep, iters);
    cudaEventRecord(e1, ctx->stream);
    CU_TRY(ctx, cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  ctx->launches += 6;
  const double instr = (double)blocks * threads * iters * 64.0 * ops_per_inner;
  *per_s = instr / (best * 1e-3);
  return STARK_OK;
}

extern "C" int stark_bench_int_peak(stark_ctx *ctx, double *imad_per_s, double *alu_per_s, double *mixed_per_s) {
  if (!ctx || !imad_per_s || !alu_per_s || !mixed_per_s) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  u32 *d_out = nullptr;
  Scratch sc(ctx);
  ST_TRY(sc.get(&d_out, 16));
  // instructions per innermost statement (pinned with inline PTX): MODE 0: 1 IMAD; MODE 1: 1 LOP3; MODE 2: IMAD + LOP3
  int rc = run_mode<0>(ctx, d_out, 1.0, imad_per_s);
  if (rc == STARK_OK) rc = run_mode<1>(ctx, d_out, 1.0, alu_per_s);
  if (rc == STARK_OK) rc = run_mode<2>(ctx, d_out, 2.0, mixed_per_s);
  return rc;
}

// rates of the multiply forms a Montgomery product is made of (thread-operations per second): out[0] IMAD.WIDE.U32 (each with one LOP3 beside it),
// out[1] IMAD.HI.U32, out[2] Montgomery products, out[3] butterfly steps (product + add + reduction)
extern "C" int stark_bench_mul_peak(stark_ctx *ctx, double *out) {
  if (!ctx || !out) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  u32 *d_out = nullptr;
  Scratch sc(ctx);
  ST_TRY(sc.get(&d_out, 16));
  int rc = run_mode<3>(ctx, d_out, 1.0, out + 0);
  if (rc == STARK_OK) rc = run_mode<4>(ctx, d_out, 1.0, out + 1);
  if (rc == STARK_OK) rc = run_mode<5>(ctx, d_out, 1.0, out + 2);
  if (rc == STARK_OK) rc = run_mode<6>(ctx, d_out, 1.0, out + 3);
  return rc;
}

// ---- dependent-hash latency of ONE warp alone on an SM (the narrow Merkle levels and the FRI tail are chains of such
// hashes): K chained Hash::combine(x, x) with the one-hash-per-thread (hs), two-per-thread (hs2) and four-lanes-per-hash
// (hsq) forms; clock64 cycles per hash.
#include "hash.cuh"
__global__ void k_hash_latency(int mode, int K, unsigned long long *cycles, u32 *sink) {
  __shared__ __align__(16) u8 buf[8 * 32];   // hsq reads its children from memory: per-quad scratch
  u32 x[8], y[8];
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 8 + i, y[i] = ~x[i];
  u8 *mine = buf + 32 * (threadIdx.x >> 2);
  const u32 q = threadIdx.x & 3u;
  *reinterpret_cast<uint2 *>(mine + 8 * q) = make_uint2(x[0], x[1]);
  __syncwarp();
  const unsigned long long t0 = clock64();
  for (int k = 0; k < K; k++) {
    if (mode == 0) {
      u32 o[8];
      hs::combine(x, x, o);
      for (int i = 0; i < 8; i++) x[i] = o[i];
    } else if (mode == 1) {
      u32 oa[8], ob[8];
      hs2::combine2(x, x, y, y, oa, ob, blockDim.y);
      for (int i = 0; i < 8; i++) x[i] = oa[i], y[i] = ob[i];
    } else if (mode == 2) {
      u32 o0, o1;
      hsq::combine(mine, mine, o0, o1);
      __syncwarp();
      *reinterpret_cast<uint2 *>(mine + 8 * q) = make_uint2(o0, o1);
      __syncwarp();
      x[0] = o0;
    } else {
      // eight lanes per hash (hso): octet o of the warp uses scratch hash o
      const hso::Dev w;
      u8 *mine8 = buf + 32 * (threadIdx.x >> 3);
      const u32 o = hso::combine(w, mine8, mine8);
      __syncwarp();
      *reinterpret_cast<u32 *>(mine8 + 4 * w.q) = o;
      __syncwarp();
      x[0] = o;
    }
  }
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) *cycles = (t1 - t0) / (unsigned long long)K;
  if (x[0] == 0x12345678u && y[0] == 1u) *sink = x[1];
}
// the same for the eight-lanes-per-hash form (hso), which the narrow steps of the Merkle climb and the transcript use
extern "C" int stark_bench_hash_latency_hso(stark_ctx *ctx, double *hso_cycles) {
  if (!ctx || !hso_cycles) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  unsigned long long *d = nullptr, h = 0;
  Scratch sc(ctx);
  ST_TRY(sc.get(&d, 64));
  k_hash_latency<<<1, 32, 0, ctx->stream>>>(3, 4, d, (u32 *)(d + 4));    // warm-up (instruction cache)
  k_hash_latency<<<1, 32, 0, ctx->stream>>>(3, 64, d, (u32 *)(d + 4));
  ctx->launches += 2;
  CU_TRY(ctx, cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *hso_cycles = (double)h;
  return STARK_OK;
}
extern "C" int stark_bench_hash_latency(stark_ctx *ctx, double *hs_cycles, double *hs2_cycles, double *hsq_cycles) {
  if (!ctx || !hs_cycles || !hs2_cycles || !hsq_cycles) return stark_fail(ctx, STARK_ERR_ARG, "null argument");
  unsigned long long *d = nullptr, h[3];
  Scratch sc(ctx);
  ST_TRY(sc.get(&d, 64));
  for (int mode = 0; mode < 3; mode++) {
    k_hash_latency<<<1, 32, 0, ctx->stream>>>(mode, 4, d + mode, (u32 *)(d + 4));    // warm-up (instruction cache)
    k_hash_latency<<<1, 32, 0, ctx->stream>>>(mode, 64, d + mode, (u32 *)(d + 4));
    ctx->launches += 2;
  }
  CU_TRY(ctx, cudaMemcpyAsync(h, d, 24, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *hs_cycles = (double)h[0], *hs2_cycles = (double)h[1], *hsq_cycles = (double)h[2];
  return STARK_OK;
}
