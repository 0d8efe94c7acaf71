/*
 * stark_b200.h -- C ABI of libstark_b200.so, the B200 (sm_100a) back end for the data-parallel hot
 * path of 0xSooki/stark-rs.  This is the drop-in boundary: the reference has no FFI of its own, so each
 * entry point below names the reference function (file:line under the reference's src/) whose results
 * it reproduces bit for bit.  The Rust shim in stark-rs_b200/rust/ and the C++ mirror in
 * stark-rs_b200/host/ forward the reference's public methods to these symbols.
 *
 * Conventions
 *   - Every function returns a stark_status (0 = ok) and never aborts; stark_last_error() gives the
 *     thread-local message.  STARK_ERR_ARG mirrors a reference assert/panic and carries ITS message text
 *     ("no inverse", "Number of leaves must be power of 2", ...) so the host shim can re-raise it.
 *   - Field values cross the boundary as contiguous little-endian uint64_t (FieldElement.value, ff.rs:25-28)
 *     and must be canonical (< p = 998244353, ff.rs:191-197) unless a parameter says "raw"; results are
 *     canonical.  A non-canonical input is STARK_ERR_ARG, never silently reduced.
 *   - Hashes are uint8_t[32] (hash.rs:2), contiguous.
 *   - Host buffers are caller-owned.  Device objects are opaque handles created and freed by the library and
 *     valid only on the context that made them.  A context owns one CUDA device + one stream and is not
 *     thread-safe; use one per thread / per process (one process per GPU).
 *   - On the device an element is ONE uint32_t (canonical).  stark_buf wraps such an array.
 *   - ALIGNMENT: device arrays handed to the library (stark_buf_wrap, peer / multicast addresses) must be 16-byte aligned -- the kernels use 128-bit accesses -- and are rejected with STARK_ERR_ARG
 *     otherwise.  Element offsets into a stark_buf (out_off, i0 of the range entry points) may be arbitrary: a range
 *     whose addresses are not 16-byte aligned takes the scalar kernel.
 *   - There is no CPU fallback: every compute entry point fails with STARK_ERR_CUDA when no device is usable.
 */
#ifndef STARK_B200_H
#define STARK_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STARK_P 998244353ull /* ff.rs:191-197, main.rs:6 */

typedef enum {
  STARK_OK = 0,
  STARK_ERR_ARG = 1,  /* precondition = a reference assert!/panic! */
  STARK_ERR_CUDA = 2,
  STARK_ERR_NCCL = 3,
  STARK_ERR_OOM = 4
} stark_status;

typedef struct stark_ctx stark_ctx;
typedef struct stark_buf stark_buf;             /* device array of uint32_t field elements */
typedef struct stark_tree stark_tree;           /* MerkleTree (merkle.rs:4-8): all levels on device */
typedef struct stark_fri_state stark_fri_state; /* result of Fri::commit (fri.rs:105-156) kept on device */

/* ---- context ------------------------------------------------------------------------------------------ */
int stark_ctx_create(int device, stark_ctx **out);
/* borrow an existing CUDA stream (e.g. torch.cuda.current_stream().cuda_stream) */
int stark_ctx_create_on_stream(int device, void *cuda_stream, stark_ctx **out);
void stark_ctx_destroy(stark_ctx *ctx);
int stark_ctx_sync(stark_ctx *ctx);
/* cudaSetDevice(the context's device): needed only by a host thread that uses contexts on DIFFERENT devices in turn */
int stark_ctx_make_current(stark_ctx *ctx);
void *stark_ctx_stream(stark_ctx *ctx);
uint64_t stark_ctx_launches(stark_ctx *ctx); /* kernels launched so far through this context */
/* per-kernel device timing with CUDA events on the context's stream (measurement only; bench.py roofline) */
int stark_ctx_profile_begin(stark_ctx *ctx);
int stark_ctx_profile_end(stark_ctx *ctx, char *json, size_t cap);
/* register-only integer issue-rate microbenchmark: thread-instructions per second of IMAD, LOP3/IADD3 and a
 * 1:1 mix, measured with CUDA events (no integer peak is recorded in MEASURED_PEAKS.json) */
int stark_bench_int_peak(stark_ctx *ctx, double *imad_per_s, double *alu_per_s, double *mixed_per_s);
/* rates of the pieces of a Montgomery product (ff.rs:138-144 replacement), thread-operations per second: out[0]
 * IMAD.WIDE.U32 (32 x 32 -> 64, both halves consumed), out[1] IMAD.HI.U32, out[2] Montgomery products (3 multiplies each), out[3] NTT
 * butterfly steps (product + lazy add + range reduction) */
int stark_bench_mul_peak(stark_ctx *ctx, double *out4);
/* dependent-hash latency (clock cycles per Hash::combine) of one warp alone on an SM, for the one-hash-per-thread,
 * two-per-thread and four-lanes-per-hash kernels forms (measurement only) */
int stark_bench_hash_latency(stark_ctx *ctx, double *hs_cycles, double *hs2_cycles, double *hsq_cycles);
int stark_bench_hash_latency_hso(stark_ctx *ctx, double *hso_cycles); /* eight lanes per hash (round 2) */
const char *stark_last_error(void);
const char *stark_version(void);

/* ---- device buffers ----------------------------------------------------------------------------------- */
int stark_buf_alloc(stark_ctx *ctx, size_t n, stark_buf **out);
int stark_buf_upload(stark_ctx *ctx, const uint64_t *host, size_t n, stark_buf **out); /* checks canonical */
int stark_buf_upload_into(stark_ctx *ctx, const uint64_t *host, size_t n, stark_buf *dst, size_t dst_off);
int stark_buf_download(stark_ctx *ctx, const stark_buf *buf, size_t off, size_t n, uint64_t *host);
int stark_buf_wrap(stark_ctx *ctx, void *device_u32, size_t n, stark_buf **out); /* non-owning view */
void *stark_buf_ptr(const stark_buf *buf);
size_t stark_buf_len(const stark_buf *buf);
void stark_buf_free(stark_buf *buf);

/* ---- K0: ff.rs batch arithmetic (ff.rs:138-213) ------------------------------------------------------- */
int stark_ff_vec_add(stark_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n); /* ff.rs:146 */
int stark_ff_vec_sub(stark_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n); /* ff.rs:154 */
int stark_ff_vec_mul(stark_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n); /* ff.rs:138 */
int stark_ff_vec_neg(stark_ctx *ctx, const uint64_t *a, uint64_t *out, size_t n);                    /* ff.rs:162 */
/* ff.rs:169-178; STARK_ERR_ARG "no inverse" if any a[i] == 0 (batch inversion on device) */
int stark_ff_vec_inv(stark_ctx *ctx, const uint64_t *a, uint64_t *out, size_t n);
int stark_ff_vec_pow(stark_ctx *ctx, const uint64_t *a, uint64_t e, uint64_t *out, size_t n);        /* ff.rs:200 */
/* ff.rs:215-223; STARK_ERR_ARG "n must be a power of two" / "n > 2^23 not supported by this modulus" */
int stark_ff_prim_nth_root(uint64_t n, uint64_t *out);

/* ---- K1/K3: univariate (src/univariate) --------------------------------------------------------------- */
/* Polynomial::mul (mul.rs:6-29).  *out_len = 0 if either side is the zero polynomial, else na+nb-1
 * (vector length, not degree).  out must hold na+nb-1 values. */
int stark_poly_mul(stark_ctx *ctx, const uint64_t *a, size_t na, const uint64_t *b, size_t nb, uint64_t *out,
                   size_t *out_len);
/* Polynomial::div (div.rs:6-53): quotient and remainder with the reference's vector lengths (intdiv = q with a zero
 * remainder, modulo = r).  STARK_ERR_ARG "No division by zero" when the divisor is the zero polynomial.  q must hold
 * na values, r must hold na + nb values.  O(n log n): Newton inversion of the reversed divisor over the NTT. */
int stark_poly_div(stark_ctx *ctx, const uint64_t *a, size_t na, const uint64_t *b, size_t nb, uint64_t *q,
                   size_t *q_len, uint64_t *r, size_t *r_len);
/* Polynomial::eval_domain (eval.rs:16-21) on domain[i] = offset * w_N^i, N = 2^log_n, w_N = prim_nth_root(N)
 * (the pattern of fri.rs:575-578); nc <= N coefficients; natural order. */
int stark_poly_eval_coset(stark_ctx *ctx, const uint64_t *coeffs, size_t nc, uint64_t offset, uint32_t log_n,
                          uint64_t *out);
/* Polynomial::interpolate_domain (interpolate.rs:6-44) on the same domain.  coeffs gets N values; *out_len
 * follows the reference's shape rule: N if any value is non-zero, else 0 (N >= 2) or 1 (N == 1). */
int stark_poly_interpolate_coset(stark_ctx *ctx, const uint64_t *vals, uint64_t offset, uint32_t log_n,
                                 uint64_t *coeffs, size_t *out_len);
/* eval_domain on an ARBITRARY domain (eval.rs:16-21), O(nc*m) on device */
int stark_poly_eval_domain(stark_ctx *ctx, const uint64_t *coeffs, size_t nc, const uint64_t *domain, size_t m,
                           uint64_t *out);
/* interpolate_domain on an ARBITRARY domain (interpolate.rs:6-44), O(n^2) on device.
 * STARK_ERR_ARG "no inverse" on duplicate points (mod.rs:613-625).  Same *out_len rule as above. */
int stark_poly_interpolate_domain(stark_ctx *ctx, const uint64_t *domain, const uint64_t *vals, size_t n,
                                  uint64_t *coeffs, size_t *out_len);
/* Polynomial::scale, f(cX) (mod.rs:99-113) */
int stark_poly_scale(stark_ctx *ctx, const uint64_t *coeffs, size_t n, uint64_t factor, uint64_t *out);
/* Polynomial::zerofier (mod.rs:77-96) of the coset {offset * w_N^i}: X^N - offset^N, N+1 coefficients */
int stark_poly_zerofier_coset(stark_ctx *ctx, uint64_t offset, uint32_t log_n, uint64_t *out);
/* Polynomial::zerofier of an arbitrary domain (mod.rs:77-96), n+1 coefficients */
int stark_poly_zerofier_domain(stark_ctx *ctx, const uint64_t *domain, size_t n, uint64_t *out);

/* ---- K2: trace low-degree extension (trace.rs:21-34 columns, pattern fri.rs:575-578) ------
 * out column c = eval_domain(interpolate_domain([w_n^i], col c), [offset * w_{bn}^i]); column-major in and
 * out (column c at cols + c*n, out + c*n*b).  n = 2^log_n, b = 2^log_blowup, log_n + log_blowup <= 23. */
int stark_lde(stark_ctx *ctx, const uint64_t *cols, uint32_t n_cols, uint32_t log_n, uint32_t log_blowup,
              uint64_t offset, uint64_t *out);
int stark_lde_dev(stark_ctx *ctx, const stark_buf *cols, uint32_t n_cols, uint32_t log_n, uint32_t log_blowup,
                  uint64_t offset, stark_buf *out);
/* device-resident transforms used by the pipelines above (natural order in and out) */
int stark_ntt_dev(stark_ctx *ctx, const stark_buf *in, stark_buf *out, uint32_t log_n, uint32_t batch, int inverse);

/* ---- K5: hash.rs / merkle.rs -------------------------------------------------------------------------- */
/* Hash::from_bytes (hash.rs:7-30) of n_msgs messages of msg_len bytes each (contiguous) */
int stark_hash_bytes(stark_ctx *ctx, const uint8_t *msgs, size_t n_msgs, size_t msg_len, uint8_t *out);
/* leaf i = Hash::from_field_elements(&vals[i*width .. (i+1)*width]) (hash.rs:32-35; fri.rs:118-121 has width 1) */
int stark_hash_leaves(stark_ctx *ctx, const uint64_t *vals, size_t n_leaves, uint32_t width, uint8_t *out);
/* MerkleTree::new (merkle.rs:11-38).  STARK_ERR_ARG "Cannot create tree from empty leaves" /
 * "Number of leaves must be power of 2". */
int stark_merkle_build(stark_ctx *ctx, const uint8_t *leaves, size_t n, stark_tree **out);
/* leaf hashing + MerkleTree::new in one go (fri.rs:118-127); row-major values, width per leaf */
int stark_merkle_build_from_values(stark_ctx *ctx, const uint64_t *vals, size_t n_leaves, uint32_t width,
                                   stark_tree **out);
/* same, values already on device; column-major [width][n_leaves] when width > 1 (LDE output layout) */
int stark_merkle_build_from_buf(stark_ctx *ctx, const stark_buf *vals, size_t n_leaves, uint32_t width,
                                stark_tree **out);
/* MerkleTree::new over leaf hashes already on the device (sharded prover: the gathered subtree roots) */
int stark_merkle_build_dev(stark_ctx *ctx, const void *leaves_dev, size_t n, stark_tree **out);
/* device address of the flattened MerkleTree.nodes (merkle.rs:18-29): level l starts at hash 2n - 2(n >> l) */
void *stark_merkle_nodes_ptr(const stark_tree *t);
/* MerkleTree::open (merkle.rs:67-80) for n_idx leaves in one call: out[(q*log2(n) + l)*32 ..] */
int stark_merkle_open_batch(stark_tree *t, const uint64_t *idx, size_t n_idx, uint8_t *out);
/* MerkleTree::commit (merkle.rs:44-65): root only */
int stark_merkle_commit(stark_ctx *ctx, const uint8_t *leaves, size_t n, uint8_t root[32]);
int stark_merkle_root(stark_tree *t, uint8_t root[32]);                          /* merkle.rs:40-42 */
size_t stark_merkle_num_leaves(const stark_tree *t);
uint32_t stark_merkle_num_levels(const stark_tree *t);                           /* nodes.len(), merkle.rs:18-29 */
int stark_merkle_level(stark_tree *t, uint32_t level, uint8_t *out);             /* nodes[level] */
/* MerkleTree::open (merkle.rs:67-80): log2(n) sibling hashes; STARK_ERR_ARG "Index out of bounds" */
int stark_merkle_open(stark_tree *t, size_t index, uint8_t *out, size_t *n_hashes);
void stark_merkle_free(stark_tree *t);

/* ---- K4/K6/K7: fri.rs --------------------------------------------------------------------------------- */
/* Fri::num_rounds (fri.rs:93-103); also checks the Fri::new asserts (fri.rs:37-45) */
int stark_fri_num_rounds(size_t domain_length, uint32_t expansion_factor, uint32_t num_colinearity_tests,
                         uint32_t *rounds);
/* Fri::fold_codeword (fri.rs:57-91); alpha_raw may be unreduced (fiat_shamir.rs:21-24); n even */
int stark_fri_fold(stark_ctx *ctx, const uint64_t *codeword, size_t n, uint64_t alpha_raw, uint64_t offset,
                   uint64_t omega, uint64_t *out);
int stark_fri_fold_dev(stark_ctx *ctx, const stark_buf *codeword, size_t n, uint64_t alpha_raw, uint64_t offset,
                       uint64_t omega, stark_buf *out);
/* one rank's share of Fri::fold_codeword (fri.rs:57-91), SURVEY 8(e): outputs [i0, i0+count) -> out[out_off ..] */
int stark_fri_fold_range_dev(stark_ctx *ctx, const stark_buf *codeword, size_t n, uint64_t alpha_raw, uint64_t offset,
                             uint64_t omega, size_t i0, size_t count, stark_buf *out, size_t out_off);
/* the same fused with the replication of the result: every output is stored straight into all ranks' replicas of the
 * next codeword (peers[g], device addresses mapped on this device) or once through the NVSwitch multicast address */
int stark_fri_fold_bcast_dev(stark_ctx *ctx, const stark_buf *codeword, size_t n, uint64_t alpha_raw, uint64_t offset,
                             uint64_t omega, size_t i0, size_t count, void *const *peers, int n_peers, void *multicast);
/* host-side transcript helpers of the sharded prover: FiatShamir::challenge (fiat_shamir.rs:19-25, raw u64) of a
 * host-held transcript and Hash::from_u64 (hash.rs:37-39, the index seed of fri.rs:272) */
int stark_fiat_shamir_challenge(const uint8_t *transcript, size_t len, uint64_t *challenge_raw);
int stark_hash_from_u64(uint64_t value, uint8_t out[32]);
/* Fri::commit (fri.rs:105-156) entirely on device, Fiat-Shamir included (fiat_shamir.rs:15-25); the
 * transcript starts with `transcript` (may be NULL/0 = FiatShamir::new()).  Keeps every codeword and tree. */
int stark_fri_commit(stark_ctx *ctx, const uint64_t *codeword, size_t n, uint64_t offset, uint64_t omega,
                     uint32_t expansion_factor, uint32_t num_colinearity_tests, const uint8_t *transcript,
                     size_t transcript_len, stark_fri_state **out);
int stark_fri_commit_dev(stark_ctx *ctx, const stark_buf *codeword, size_t n, uint64_t offset, uint64_t omega,
                         uint32_t expansion_factor, uint32_t num_colinearity_tests, const uint8_t *transcript,
                         size_t transcript_len, stark_fri_state **out);
uint32_t stark_fri_rounds(const stark_fri_state *s);
int stark_fri_roots(stark_fri_state *s, uint8_t *out /* 32*rounds */);
int stark_fri_alphas(stark_fri_state *s, uint64_t *out /* rounds-1 raw challenges */);
int stark_fri_codeword_len(const stark_fri_state *s, uint32_t round, size_t *len);
int stark_fri_codeword(stark_fri_state *s, uint32_t round, uint64_t *out);
int stark_fri_open(stark_fri_state *s, uint32_t round, size_t index, uint8_t *out, size_t *n_hashes);
void stark_fri_free(stark_fri_state *s);
/* Fri::sample_indices (fri.rs:176-213); STARK_ERR_ARG with the reference's two messages */
int stark_fri_sample_indices(const uint8_t *seed, size_t seed_len, size_t size, size_t reduced_size, size_t number,
                             uint64_t *out);
/* Fri::prove (fri.rs:250-311) + ProofStream::serialize (stream.rs:35-64): the exact proof bytes.
 * top_indices (may be NULL) receives the num_colinearity_tests returned indices.  STARK_ERR_ARG
 * "initial codeword length does not match domain length" when n != domain_length. */
int stark_fri_proof_size(size_t domain_length, uint32_t expansion_factor, uint32_t num_colinearity_tests,
                         size_t *bytes);
int stark_fri_prove(stark_ctx *ctx, const uint64_t *codeword, size_t n, size_t domain_length, uint64_t offset,
                    uint64_t omega, uint32_t expansion_factor, uint32_t num_colinearity_tests,
                    const uint8_t *transcript, size_t transcript_len, uint8_t *proof, size_t proof_cap,
                    size_t *proof_len, uint64_t *top_indices);
int stark_fri_prove_dev(stark_ctx *ctx, const stark_buf *codeword, size_t n, size_t domain_length, uint64_t offset,
                        uint64_t omega, uint32_t expansion_factor, uint32_t num_colinearity_tests,
                        const uint8_t *transcript, size_t transcript_len, uint8_t *proof, size_t proof_cap,
                        size_t *proof_len, uint64_t *top_indices);

/* ---- pipeline: BASELINE config 3 ("2^20-row trace: coset LDE + Merkle commit + FRI") ------------------
 * LDE of n_cols columns, one Merkle tree per column (leaf rule fri.rs:118-121), Fri::prove on column 0's
 * codeword with omega = prim_nth_root(n*blowup).  column_roots gets 32*n_cols bytes. */
int stark_prove_trace(stark_ctx *ctx, const uint64_t *cols, uint32_t n_cols, uint32_t log_n, uint32_t log_blowup,
                      uint64_t offset, uint32_t num_colinearity_tests, uint8_t *column_roots, uint8_t *proof,
                      size_t proof_cap, size_t *proof_len);
int stark_prove_trace_dev(stark_ctx *ctx, const stark_buf *cols, uint32_t n_cols, uint32_t log_n,
                          uint32_t log_blowup, uint64_t offset, uint32_t num_colinearity_tests,
                          uint8_t *column_roots, uint8_t *proof, size_t proof_cap, size_t *proof_len);
/* Fri::verify (fri.rs:313-505) with test_colinearity (fri.rs:507-525) and MerkleTree::verify (merkle.rs:82-97) as a
 * batch verifier on the device: `proof` is ProofStream::serialize's bytes (stream.rs:35-64), `transcript` the
 * caller's FiatShamir state before the call (fiat_shamir.rs:4-13; empty for a fresh one).  *ok = the reference's
 * return value; *reason (optional) = 0 or the index of the reference's `println!` text, see stark_fri_verify_reason.
 * Optional outputs: roots_out = the num_rounds() Merkle roots popped from the stream (the shim absorbs them into its
 * FiatShamir like fri.rs:327), top_indices = the num_colinearity_tests sampled indices (fri.rs:401-407), poly_indices /
 * poly_values = the 2 * num_colinearity_tests (index, value) pairs of `polynomial_values` (fri.rs:437-441); the last
 * three are written only when *ok (polynomial_values only when num_rounds() > 1: the reference fills it in query
 * round 0 and there are num_rounds() - 1 query rounds; otherwise the two arrays are left untouched).  Status 1 mirrors the reference's panics: Fri::new asserts (fri.rs:37-45),
 * MerkleTree::new on an empty / non-power-of-two last codeword (merkle.rs:12-16), the sample_indices asserts
 * (fri.rs:183-192), u128 underflow in FiniteField::sub on non-canonical stream values (ff.rs:154-160).  omega must be
 * FiniteField::prim_nth_root(domain_length) (what every Fri in the reference is built with). */
int stark_fri_verify(stark_ctx *ctx, const uint8_t *proof, size_t proof_len, size_t domain_length, uint64_t offset,
                     uint64_t omega, uint32_t expansion_factor, uint32_t num_colinearity_tests,
                     const uint8_t *transcript, size_t transcript_len, int *ok, uint32_t *reason, uint8_t *roots_out,
                     uint64_t *top_indices, uint64_t *poly_indices, uint64_t *poly_values);
const char *stark_fri_verify_reason(uint32_t reason);

/* Trace ingestion (trace.rs:4-34: Trace { trace: Vec<Vec<i128>> }, get_col, to_field_elements): `rows_i128` is the
 * row-major matrix, n_rows x n_cols values of 16 little-endian bytes each.  Every value is cast `as u64` like the
 * reference (trace.rs:29-34) and taken mod p -- the reference's FieldElement keeps the raw cast, but the trace only
 * enters the LDE through FiniteField::mul/add, which reduce (ff.rs:138-152).  Result: the column-major device matrix
 * stark_lde_dev / stark_prove_trace_dev take (column c at c * n_rows). */
int stark_trace_to_columns(stark_ctx *ctx, const void *rows_i128, size_t n_rows, uint32_t n_cols, stark_buf **out);
int stark_prove_trace_rows(stark_ctx *ctx, const void *rows_i128, uint32_t n_cols, uint32_t log_n, uint32_t log_blowup,
                           uint64_t offset, uint32_t num_colinearity_tests, uint8_t *column_roots, uint8_t *proof,
                           size_t proof_cap, size_t *proof_len);

/* ---- groups of GPUs ---------------------------------------------------------------------------------
 * The reference is single-threaded and has no parallel path; these entry points shard ITS functions over
 * the GPUs of one node with byte-identical results for every group size: independent trace-column LDEs and column trees
 * (no exchange), Merkle trees split by leaf range (MerkleTree::new, merkle.rs:11-38: the ranks' subtree roots are
 * exchanged and the top levels replicated), fold_codeword split by output range (fri.rs:57-91: every rank stores its
 * slice into all replicas of the next codeword).  The library owns the NCCL communicator (loaded with dlopen) and the
 * peer-memory windows; a host in any language drives it through these calls.  A stark_mgpu_* call makes the ranks' devices
 * current while it runs and leaves the calling thread on the device it was on (single-context calls do NOT switch
 * devices: a thread that uses contexts on different devices in turn calls stark_ctx_make_current, above).
 *
 * A stark_mgpu is ONE RANK's membership in a group of `world` (1, 2, 4 or 8) GPUs.  Two ways to make a group:
 *   stark_mgpu_init          one process or thread per GPU.  Rank 0 calls stark_mgpu_unique_id and hands the 128 bytes
 *                            to the others (any channel); every rank then calls stark_mgpu_init on its own context.
 *                            Collective: all ranks must call it, and every later operation, in the same order.
 *   stark_mgpu_create_local  one host thread driving `world` contexts (distinct devices with peer access -- or the same
 *                            device, which is how single-GPU tests exercise the sharded path; such a group runs in
 *                            lock step).  Fills out[0 .. world).
 * Operations take the ranks THIS CALL DRIVES: one handle (n_here = 1) in a multi-process group, all `world` handles in
 * rank order in a local group; per-rank arguments are arrays of n_here entries.  Errors: STARK_ERR_NCCL when NCCL fails
 * or a peer does not arrive within 4 s (the kernels never spin forever); STARK_ERR_ARG mirrors the reference's panics as
 * everywhere else.  max_codeword = the longest codeword / LDE column the group will fold (sizes the window).          */
typedef struct stark_mgpu stark_mgpu;
#define STARK_MGPU_ID_BYTES 128
int stark_mgpu_unique_id(uint8_t id[STARK_MGPU_ID_BYTES]);
int stark_mgpu_init(stark_ctx *ctx, const uint8_t id[STARK_MGPU_ID_BYTES], int rank, int world, size_t max_codeword,
                    stark_mgpu **out);
int stark_mgpu_create_local(stark_ctx *const *ctxs, int world, size_t max_codeword, stark_mgpu **out);
void stark_mgpu_destroy(stark_mgpu *m); /* a local group is destroyed as a whole, through any one of its handles */
int stark_mgpu_rank(const stark_mgpu *m);
int stark_mgpu_world(const stark_mgpu *m);
/* bytes this rank has stored into its peers or handed to NCCL so far (data path only) */
uint64_t stark_mgpu_bytes_sent(const stark_mgpu *m);
/* FRI rounds of at least 2^log_n elements are sharded, shorter ones run replicated (default 17, STARK_MGPU_SHARD_LOG) */
int stark_mgpu_set_shard_log(stark_mgpu *m, uint32_t log_n);
int stark_mgpu_barrier(stark_mgpu *m); /* device-side barrier over the group + synchronisation of this rank's stream */
/* the trace columns (indices > 0) this rank commits in stark_mgpu_prove_trace; returns their number */
uint32_t stark_mgpu_owned_columns(const stark_mgpu *m, uint32_t n_cols, uint32_t *out);
/* the same partition for any (rank, world) without a handle: host logic only, no device needed */
uint32_t stark_mgpu_columns_of_rank(int rank, int world, uint32_t n_cols, uint32_t *out);

/* BASELINE config 3 on a group = stark_prove_trace with the work sharded: column 0 is LDE'd by every rank (its codeword
 * is the FRI input), columns 1.. are LDE'd and committed round robin (column c by rank (c - 1) % world), Fri::prove runs
 * sharded (fri.rs:250-311).  `cols` is the WHOLE column-major trace on every rank (each rank uploads only what it needs);
 * the _dev form takes per rank a buffer holding column 0 followed by the rank's owned columns.  EVERY rank receives all
 * n_cols column roots and the complete proof bytes, identical to stark_prove_trace's.                               */
int stark_mgpu_prove_trace(stark_mgpu *const *ranks, int n_here, const uint64_t *cols, uint32_t n_cols, uint32_t log_n,
                           uint32_t log_blowup, uint64_t offset, uint32_t num_colinearity_tests,
                           uint8_t *const *column_roots, uint8_t *const *proofs, size_t proof_cap, size_t *proof_len);
int stark_mgpu_prove_trace_dev(stark_mgpu *const *ranks, int n_here, const stark_buf *const *my_cols, uint32_t n_cols,
                               uint32_t log_n, uint32_t log_blowup, uint64_t offset, uint32_t num_colinearity_tests,
                               uint8_t *const *column_roots, uint8_t *const *proofs, size_t proof_cap, size_t *proof_len);
/* Fri::prove (fri.rs:250-311) + ProofStream::serialize on a group; codewords[k] = rank k's replica of the codeword */
int stark_mgpu_fri_prove_dev(stark_mgpu *const *ranks, int n_here, const stark_buf *const *codewords, size_t n,
                             size_t domain_length, uint64_t offset, uint64_t omega, uint32_t expansion_factor,
                             uint32_t num_colinearity_tests, const uint8_t *transcript, size_t transcript_len,
                             uint8_t *const *proofs, size_t proof_cap, size_t *proof_len, uint64_t *const *top_indices);
/* BASELINE config 5: ONE round of Fri::commit (fri.rs:116-147) on a replicated codeword -- leaf hashes + tree (sharded by
 * leaf range), alpha from a fresh transcript, fold (sharded by output range, stored into every replica).  roots[k] (32
 * bytes), alpha_raw[k], folded[k] (a view of rank k's replica of the folded codeword inside the window: valid until the
 * group's next operation; free the view with stark_buf_free) are per driven rank; omega need not have order n
 * (Fri::new never checks it, fri.rs:30-55).                                                                        */
int stark_mgpu_fold_commit_round(stark_mgpu *const *ranks, int n_here, const stark_buf *const *codewords, size_t n,
                                 uint64_t offset, uint64_t omega, uint8_t *const *roots, uint64_t *alpha_raw,
                                 stark_buf **folded);
/* BASELINE config 4: n_groups fixed groups of group_width trace columns (column-major, group k at cols + k *
 * group_width * 2^log_n); rank g owns groups {g, g + world, ...}: coset LDE (eval.rs / interpolate.rs composed, SURVEY
 * 3.4) and one Merkle tree per group with leaf i = Hash::from_field_elements(row i of the group's LDE) (hash.rs:32-35);
 * the group roots are gathered (ncclAllGather in a multi-process group) and the commitment is MerkleTree::new over them
 * (merkle.rs:11-38).  The _dev form takes, per driven rank, n_groups / world device buffers (rank-major array).
 * group_roots[k]: n_groups x 32 bytes; commitments[k]: 32 bytes.                                                    */
int stark_mgpu_lde_commit(stark_mgpu *const *ranks, int n_here, const uint64_t *cols, uint32_t n_groups, uint32_t group_width,
                          uint32_t log_n, uint32_t log_blowup, uint64_t offset, uint8_t *const *group_roots,
                          uint8_t *const *commitments);
int stark_mgpu_lde_commit_dev(stark_mgpu *const *ranks, int n_here, const stark_buf *const *owned_groups, uint32_t n_groups,
                              uint32_t group_width, uint32_t log_n, uint32_t log_blowup, uint64_t offset,
                              uint8_t *const *group_roots, uint8_t *const *commitments);

#ifdef __cplusplus
}
#endif
#endif /* STARK_B200_H */
