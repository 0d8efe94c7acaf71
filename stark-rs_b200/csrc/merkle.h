// merkle.h -- internal handle layouts and device-pointer entry points shared by the .cu files.
#pragma once
#include "common.cuh"
#include "transcript.cuh"

struct stark_buf {
  stark_ctx *ctx;
  u32 *ptr;
  size_t n;
  bool owns;
};

struct stark_tree {
  stark_ctx *ctx;
  size_t n;      // leaves
  u32 levels;    // nodes.len() of merkle.rs:18-29 = log2(n) + 1
  u8 *nodes;     // device, (2n-1) hashes, level l at hash offset 2n - 2(n >> l)
};

int merkle_check_n(stark_ctx *ctx, size_t n);
int merkle_tree_alloc(stark_ctx *ctx, size_t n, stark_tree **out);
int merkle_leaves_dev(stark_ctx *ctx, const u32 *vals, size_t n, u32 width, size_t row_stride, size_t col_stride,
                      u8 *out);
struct MgExchange;
int merkle_climb_dev(stark_ctx *ctx, u8 *nodes, size_t n, const TranscriptArgs *tr = nullptr, const MgExchange *mx = nullptr);
int merkle_mg_top_dev(stark_ctx *ctx, const TranscriptArgs *tr, const MgExchange *mx);
constexpr int CLIMB_TICKETS = 64;   // trees per batched climb launch (one last-CTA ticket each)
int merkle_climb_batch_dev(stark_ctx *ctx, u8 *nodes, size_t n, u32 batch, size_t tree_stride, const TranscriptArgs *tr,
                           const MgExchange *mx);
int merkle_build_batch_dev(stark_ctx *ctx, const u32 *vals, size_t n, u32 batch, size_t val_stride, u8 *nodes,
                           size_t tree_stride);
int merkle_build_from_dev_values(stark_ctx *ctx, const u32 *vals, size_t n, u32 width, size_t row_stride,
                                 size_t col_stride, stark_tree **out, const TranscriptArgs *tr = nullptr);
int merkle_open_dev(stark_ctx *ctx, const u8 *nodes, size_t n, const u64 *idx_dev, u32 n_idx, u8 *out_dev);

// poly.cu / api.cu helpers
int upload_u64(stark_ctx *ctx, const uint64_t *host, size_t n, u32 *dst);     // H2D + canonical check + narrow
int upload_flag_reset(stark_ctx *ctx);                                         // before a group of nosync uploads
int upload_u64_nosync(stark_ctx *ctx, const uint64_t *host, size_t n, u32 *dst);  // no host round trip ...
int upload_u64_check(stark_ctx *ctx);                                          // ... flag read after the final sync
int narrow_dev(stark_ctx *ctx, const uint64_t *staging_dev, size_t n, u32 *dst);   // values already on the device
int upload_flag_fetch(stark_ctx *ctx);                                             // queue the flag's D2H copy
int download_u64(stark_ctx *ctx, const u32 *src, size_t n, uint64_t *host);   // widen + D2H (synchronises)
int trace_to_columns_dev(stark_ctx *ctx, const void *rows_i128, size_t n_rows, u32 n_cols, u32 *cols);
int lde_dev(stark_ctx *ctx, const u32 *cols, u32 n_cols, u32 log_n, u32 log_blowup, u32 offset, u32 *out);
