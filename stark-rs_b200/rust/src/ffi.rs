//! extern "C" surface of libstark_b200.so -- generated from include/stark_b200.h (tools: see INTEGRATION.md).
//! NOT compiled here: no rustc in the build image.
#![allow(dead_code)]

#[repr(C)] pub struct StarkCtx { _private: [u8; 0] }
#[repr(C)] pub struct StarkBuf { _private: [u8; 0] }
#[repr(C)] pub struct StarkTree { _private: [u8; 0] }
#[repr(C)] pub struct StarkFriState { _private: [u8; 0] }
#[repr(C)] pub struct StarkMgpu { _private: [u8; 0] }

extern "C" {
    pub fn stark_ctx_create(device: i32, out: *mut *mut StarkCtx) -> i32;
    pub fn stark_ctx_create_on_stream(device: i32, cuda_stream: *mut std::ffi::c_void, out: *mut *mut StarkCtx) -> i32;
    pub fn stark_ctx_destroy(ctx: *mut StarkCtx);
    pub fn stark_ctx_sync(ctx: *mut StarkCtx) -> i32;
    pub fn stark_ctx_make_current(ctx: *mut StarkCtx) -> i32;
    pub fn stark_ctx_stream(ctx: *mut StarkCtx) -> *mut std::ffi::c_void;
    pub fn stark_ctx_launches(ctx: *mut StarkCtx) -> u64;
    pub fn stark_ctx_profile_begin(ctx: *mut StarkCtx) -> i32;
    pub fn stark_ctx_profile_end(ctx: *mut StarkCtx, json: *mut std::os::raw::c_char, cap: usize) -> i32;
    pub fn stark_bench_int_peak(ctx: *mut StarkCtx, imad_per_s: *mut f64, alu_per_s: *mut f64, mixed_per_s: *mut f64) -> i32;
    pub fn stark_last_error() -> *const std::os::raw::c_char;
    pub fn stark_version() -> *const std::os::raw::c_char;
    pub fn stark_buf_alloc(ctx: *mut StarkCtx, n: usize, out: *mut *mut StarkBuf) -> i32;
    pub fn stark_buf_upload(ctx: *mut StarkCtx, host: *const u64, n: usize, out: *mut *mut StarkBuf) -> i32;
    pub fn stark_buf_upload_into(ctx: *mut StarkCtx, host: *const u64, n: usize, dst: *mut StarkBuf, dst_off: usize) -> i32;
    pub fn stark_buf_download(ctx: *mut StarkCtx, buf: *const StarkBuf, off: usize, n: usize, host: *mut u64) -> i32;
    pub fn stark_buf_wrap(ctx: *mut StarkCtx, device_u32: *mut std::ffi::c_void, n: usize, out: *mut *mut StarkBuf) -> i32;
    pub fn stark_buf_ptr(buf: *const StarkBuf) -> *mut std::ffi::c_void;
    pub fn stark_buf_len(buf: *const StarkBuf) -> usize;
    pub fn stark_buf_free(buf: *mut StarkBuf);
    pub fn stark_ff_vec_add(ctx: *mut StarkCtx, a: *const u64, b: *const u64, out: *mut u64, n: usize) -> i32;
    pub fn stark_ff_vec_sub(ctx: *mut StarkCtx, a: *const u64, b: *const u64, out: *mut u64, n: usize) -> i32;
    pub fn stark_ff_vec_mul(ctx: *mut StarkCtx, a: *const u64, b: *const u64, out: *mut u64, n: usize) -> i32;
    pub fn stark_ff_vec_neg(ctx: *mut StarkCtx, a: *const u64, out: *mut u64, n: usize) -> i32;
    pub fn stark_ff_vec_inv(ctx: *mut StarkCtx, a: *const u64, out: *mut u64, n: usize) -> i32;
    pub fn stark_ff_vec_pow(ctx: *mut StarkCtx, a: *const u64, e: u64, out: *mut u64, n: usize) -> i32;
    pub fn stark_ff_prim_nth_root(n: u64, out: *mut u64) -> i32;
    pub fn stark_poly_mul(ctx: *mut StarkCtx, a: *const u64, na: usize, b: *const u64, nb: usize, out: *mut u64, out_len: *mut usize) -> i32;
    pub fn stark_poly_div(ctx: *mut StarkCtx, a: *const u64, na: usize, b: *const u64, nb: usize, q: *mut u64, q_len: *mut usize, r: *mut u64, r_len: *mut usize) -> i32;
    pub fn stark_poly_eval_coset(ctx: *mut StarkCtx, coeffs: *const u64, nc: usize, offset: u64, log_n: u32, out: *mut u64) -> i32;
    pub fn stark_poly_interpolate_coset(ctx: *mut StarkCtx, vals: *const u64, offset: u64, log_n: u32, coeffs: *mut u64, out_len: *mut usize) -> i32;
    pub fn stark_poly_eval_domain(ctx: *mut StarkCtx, coeffs: *const u64, nc: usize, domain: *const u64, m: usize, out: *mut u64) -> i32;
    pub fn stark_poly_interpolate_domain(ctx: *mut StarkCtx, domain: *const u64, vals: *const u64, n: usize, coeffs: *mut u64, out_len: *mut usize) -> i32;
    pub fn stark_poly_scale(ctx: *mut StarkCtx, coeffs: *const u64, n: usize, factor: u64, out: *mut u64) -> i32;
    pub fn stark_poly_zerofier_coset(ctx: *mut StarkCtx, offset: u64, log_n: u32, out: *mut u64) -> i32;
    pub fn stark_poly_zerofier_domain(ctx: *mut StarkCtx, domain: *const u64, n: usize, out: *mut u64) -> i32;
    pub fn stark_lde(ctx: *mut StarkCtx, cols: *const u64, n_cols: u32, log_n: u32, log_blowup: u32, offset: u64, out: *mut u64) -> i32;
    pub fn stark_lde_dev(ctx: *mut StarkCtx, cols: *const StarkBuf, n_cols: u32, log_n: u32, log_blowup: u32, offset: u64, out: *mut StarkBuf) -> i32;
    pub fn stark_ntt_dev(ctx: *mut StarkCtx, in_: *const StarkBuf, out: *mut StarkBuf, log_n: u32, batch: u32, inverse: i32) -> i32;
    pub fn stark_hash_bytes(ctx: *mut StarkCtx, msgs: *const u8, n_msgs: usize, msg_len: usize, out: *mut u8) -> i32;
    pub fn stark_hash_leaves(ctx: *mut StarkCtx, vals: *const u64, n_leaves: usize, width: u32, out: *mut u8) -> i32;
    pub fn stark_merkle_build(ctx: *mut StarkCtx, leaves: *const u8, n: usize, out: *mut *mut StarkTree) -> i32;
    pub fn stark_merkle_build_from_values(ctx: *mut StarkCtx, vals: *const u64, n_leaves: usize, width: u32, out: *mut *mut StarkTree) -> i32;
    pub fn stark_merkle_build_from_buf(ctx: *mut StarkCtx, vals: *const StarkBuf, n_leaves: usize, width: u32, out: *mut *mut StarkTree) -> i32;
    pub fn stark_merkle_commit(ctx: *mut StarkCtx, leaves: *const u8, n: usize, root: *mut u8) -> i32;
    pub fn stark_merkle_root(t: *mut StarkTree, root: *mut u8) -> i32;
    pub fn stark_merkle_num_leaves(t: *const StarkTree) -> usize;
    pub fn stark_merkle_num_levels(t: *const StarkTree) -> u32;
    pub fn stark_merkle_level(t: *mut StarkTree, level: u32, out: *mut u8) -> i32;
    pub fn stark_merkle_open(t: *mut StarkTree, index: usize, out: *mut u8, n_hashes: *mut usize) -> i32;
    pub fn stark_merkle_build_dev(ctx: *mut StarkCtx, leaves_dev: *const std::ffi::c_void, n: usize, out: *mut *mut StarkTree) -> i32;
    pub fn stark_merkle_nodes_ptr(t: *const StarkTree) -> *mut std::ffi::c_void;
    pub fn stark_merkle_open_batch(t: *mut StarkTree, idx: *const u64, n_idx: usize, out: *mut u8) -> i32;
    pub fn stark_merkle_free(t: *mut StarkTree);
    pub fn stark_fri_fold_range_dev(ctx: *mut StarkCtx, codeword: *const StarkBuf, n: usize, alpha_raw: u64, offset: u64, omega: u64, i0: usize, count: usize, out: *mut StarkBuf, out_off: usize) -> i32;
    pub fn stark_fri_fold_bcast_dev(ctx: *mut StarkCtx, codeword: *const StarkBuf, n: usize, alpha_raw: u64, offset: u64, omega: u64, i0: usize, count: usize, peers: *const *mut std::ffi::c_void, n_peers: i32, multicast: *mut std::ffi::c_void) -> i32;
    pub fn stark_fiat_shamir_challenge(transcript: *const u8, len: usize, challenge_raw: *mut u64) -> i32;
    pub fn stark_hash_from_u64(value: u64, out: *mut u8) -> i32;
    pub fn stark_bench_mul_peak(ctx: *mut StarkCtx, out4: *mut f64) -> i32;
    pub fn stark_bench_hash_latency(ctx: *mut StarkCtx, hs_cycles: *mut f64, hs2_cycles: *mut f64, hsq_cycles: *mut f64) -> i32;
    pub fn stark_bench_hash_latency_hso(ctx: *mut StarkCtx, hso_cycles: *mut f64) -> i32;
    pub fn stark_fri_verify(ctx: *mut StarkCtx, proof: *const u8, proof_len: usize, domain_length: usize, offset: u64, omega: u64, expansion_factor: u32, num_colinearity_tests: u32, transcript: *const u8, transcript_len: usize, ok: *mut i32, reason: *mut u32, roots_out: *mut u8, top_indices: *mut u64, poly_indices: *mut u64, poly_values: *mut u64) -> i32;
    pub fn stark_fri_verify_reason(reason: u32) -> *const std::os::raw::c_char;
    pub fn stark_trace_to_columns(ctx: *mut StarkCtx, rows_i128: *const std::ffi::c_void, n_rows: usize, n_cols: u32, out: *mut *mut StarkBuf) -> i32;
    pub fn stark_prove_trace_rows(ctx: *mut StarkCtx, rows_i128: *const std::ffi::c_void, n_cols: u32, log_n: u32, log_blowup: u32, offset: u64, num_colinearity_tests: u32, column_roots: *mut u8, proof: *mut u8, proof_cap: usize, proof_len: *mut usize) -> i32;
    pub fn stark_fri_num_rounds(domain_length: usize, expansion_factor: u32, num_colinearity_tests: u32, rounds: *mut u32) -> i32;
    pub fn stark_fri_fold(ctx: *mut StarkCtx, codeword: *const u64, n: usize, alpha_raw: u64, offset: u64, omega: u64, out: *mut u64) -> i32;
    pub fn stark_fri_fold_dev(ctx: *mut StarkCtx, codeword: *const StarkBuf, n: usize, alpha_raw: u64, offset: u64, omega: u64, out: *mut StarkBuf) -> i32;
    pub fn stark_fri_commit(ctx: *mut StarkCtx, codeword: *const u64, n: usize, offset: u64, omega: u64, expansion_factor: u32, num_colinearity_tests: u32, transcript: *const u8, transcript_len: usize, out: *mut *mut StarkFriState) -> i32;
    pub fn stark_fri_commit_dev(ctx: *mut StarkCtx, codeword: *const StarkBuf, n: usize, offset: u64, omega: u64, expansion_factor: u32, num_colinearity_tests: u32, transcript: *const u8, transcript_len: usize, out: *mut *mut StarkFriState) -> i32;
    pub fn stark_fri_rounds(s: *const StarkFriState) -> u32;
    pub fn stark_fri_roots(s: *mut StarkFriState, out: *mut u8) -> i32;
    pub fn stark_fri_alphas(s: *mut StarkFriState, out: *mut u64) -> i32;
    pub fn stark_fri_codeword_len(s: *const StarkFriState, round: u32, len: *mut usize) -> i32;
    pub fn stark_fri_codeword(s: *mut StarkFriState, round: u32, out: *mut u64) -> i32;
    pub fn stark_fri_open(s: *mut StarkFriState, round: u32, index: usize, out: *mut u8, n_hashes: *mut usize) -> i32;
    pub fn stark_fri_free(s: *mut StarkFriState);
    pub fn stark_fri_sample_indices(seed: *const u8, seed_len: usize, size: usize, reduced_size: usize, number: usize, out: *mut u64) -> i32;
    pub fn stark_fri_proof_size(domain_length: usize, expansion_factor: u32, num_colinearity_tests: u32, bytes: *mut usize) -> i32;
    pub fn stark_fri_prove(ctx: *mut StarkCtx, codeword: *const u64, n: usize, domain_length: usize, offset: u64, omega: u64, expansion_factor: u32, num_colinearity_tests: u32, transcript: *const u8, transcript_len: usize, proof: *mut u8, proof_cap: usize, proof_len: *mut usize, top_indices: *mut u64) -> i32;
    pub fn stark_fri_prove_dev(ctx: *mut StarkCtx, codeword: *const StarkBuf, n: usize, domain_length: usize, offset: u64, omega: u64, expansion_factor: u32, num_colinearity_tests: u32, transcript: *const u8, transcript_len: usize, proof: *mut u8, proof_cap: usize, proof_len: *mut usize, top_indices: *mut u64) -> i32;
    pub fn stark_prove_trace(ctx: *mut StarkCtx, cols: *const u64, n_cols: u32, log_n: u32, log_blowup: u32, offset: u64, num_colinearity_tests: u32, column_roots: *mut u8, proof: *mut u8, proof_cap: usize, proof_len: *mut usize) -> i32;
    // groups of GPUs (include/stark_b200.h): one StarkMgpu per rank; `ranks` = the handles this call drives
    pub fn stark_mgpu_unique_id(id: *mut u8) -> i32;
    pub fn stark_mgpu_init(ctx: *mut StarkCtx, id: *const u8, rank: i32, world: i32, max_codeword: usize, out: *mut *mut StarkMgpu) -> i32;
    pub fn stark_mgpu_create_local(ctxs: *const *mut StarkCtx, world: i32, max_codeword: usize, out: *mut *mut StarkMgpu) -> i32;
    pub fn stark_mgpu_destroy(m: *mut StarkMgpu);
    pub fn stark_mgpu_rank(m: *const StarkMgpu) -> i32;
    pub fn stark_mgpu_world(m: *const StarkMgpu) -> i32;
    pub fn stark_mgpu_bytes_sent(m: *const StarkMgpu) -> u64;
    pub fn stark_mgpu_set_shard_log(m: *mut StarkMgpu, log_n: u32) -> i32;
    pub fn stark_mgpu_barrier(m: *mut StarkMgpu) -> i32;
    pub fn stark_mgpu_owned_columns(m: *const StarkMgpu, n_cols: u32, out: *mut u32) -> u32;
    pub fn stark_mgpu_columns_of_rank(rank: i32, world: i32, n_cols: u32, out: *mut u32) -> u32;
    pub fn stark_mgpu_prove_trace(ranks: *const *mut StarkMgpu, n_here: i32, cols: *const u64, n_cols: u32, log_n: u32, log_blowup: u32, offset: u64, num_colinearity_tests: u32, column_roots: *const *mut u8, proofs: *const *mut u8, proof_cap: usize, proof_len: *mut usize) -> i32;
    pub fn stark_mgpu_prove_trace_dev(ranks: *const *mut StarkMgpu, n_here: i32, my_cols: *const *const StarkBuf, n_cols: u32, log_n: u32, log_blowup: u32, offset: u64, num_colinearity_tests: u32, column_roots: *const *mut u8, proofs: *const *mut u8, proof_cap: usize, proof_len: *mut usize) -> i32;
    pub fn stark_mgpu_fri_prove_dev(ranks: *const *mut StarkMgpu, n_here: i32, codewords: *const *const StarkBuf, n: usize, domain_length: usize, offset: u64, omega: u64, expansion_factor: u32, num_colinearity_tests: u32, transcript: *const u8, transcript_len: usize, proofs: *const *mut u8, proof_cap: usize, proof_len: *mut usize, top_indices: *const *mut u64) -> i32;
    pub fn stark_mgpu_fold_commit_round(ranks: *const *mut StarkMgpu, n_here: i32, codewords: *const *const StarkBuf, n: usize, offset: u64, omega: u64, roots: *const *mut u8, alpha_raw: *mut u64, folded: *mut *mut StarkBuf) -> i32;
    pub fn stark_mgpu_lde_commit(ranks: *const *mut StarkMgpu, n_here: i32, cols: *const u64, n_groups: u32, group_width: u32, log_n: u32, log_blowup: u32, offset: u64, group_roots: *const *mut u8, commitments: *const *mut u8) -> i32;
    pub fn stark_mgpu_lde_commit_dev(ranks: *const *mut StarkMgpu, n_here: i32, owned_groups: *const *const StarkBuf, n_groups: u32, group_width: u32, log_n: u32, log_blowup: u32, offset: u64, group_roots: *const *mut u8, commitments: *const *mut u8) -> i32;
    pub fn stark_prove_trace_dev(ctx: *mut StarkCtx, cols: *const StarkBuf, n_cols: u32, log_n: u32, log_blowup: u32, offset: u64, num_colinearity_tests: u32, column_roots: *mut u8, proof: *mut u8, proof_cap: usize, proof_len: *mut usize) -> i32;
}

/// Status 1 mirrors a reference assert!/panic!: re-raise it with the reference's own message text.
pub fn check(status: i32) {
    if status != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(stark_last_error()) }.to_string_lossy().into_owned();
        panic!("{}", msg);
    }
}

thread_local! {
    static CTX: std::cell::Cell<*mut StarkCtx> = std::cell::Cell::new(std::ptr::null_mut());
}

/// One context per thread of use, created on first use (device from STARK_B200_DEVICE, default 0).
pub fn ctx() -> *mut StarkCtx {
    CTX.with(|c| {
        if c.get().is_null() {
            let dev = std::env::var("STARK_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
            let mut p = std::ptr::null_mut();
            check(unsafe { stark_ctx_create(dev, &mut p) });
            c.set(p);
        }
        c.get()
    })
}
