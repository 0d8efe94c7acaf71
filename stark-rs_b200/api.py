"""ctypes binding of libstark_b200.so (include/stark_b200.h) + a host-side mirror of the reference's types.

This is the Python face of the drop-in boundary: thin wrappers over the C ABI, numpy in / numpy out, used by
tests/ and bench.py.  There is NO CPU fallback here: if the CUDA library is missing or no device is usable
every entry point raises.  The mirror classes (FiniteField, Polynomial, Hash, MerkleTree, FiatShamir, Fri)
keep the names, argument meaning and panic messages of the reference (src/ff.rs, src/univariate, src/hash.rs,
src/merkle.rs, src/fiat_shamir.rs, src/fri.rs) so parity tests read like the reference's own tests.
"""
import ctypes as C
import os

import numpy as np

__all__ = ["P", "StarkError", "StarkPanic", "Context", "Buffer", "MerkleTree", "FriState", "lib", "lib_path",
           "build_library", "prim_nth_root", "fri_num_rounds", "fri_proof_size", "fri_sample_indices",
           "fiat_shamir_challenge", "hash_from_u64", "Group", "mgpu_unique_id", "mgpu_columns_of_rank"]

P = 998244353
_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.path.join(_HERE, "libstark_b200.so")

U64, U32, SZ, U8P = C.c_uint64, C.c_uint32, C.c_size_t, C.POINTER(C.c_uint8)
U64P = C.POINTER(C.c_uint64)


class StarkError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("stark_b200 status %d: %s" % (status, message))
        self.status, self.message = status, message


class StarkPanic(StarkError):
    """STARK_ERR_ARG: a precondition that is an assert!/panic! in the reference (same message text)."""


def build_library(force=False):
    """Compile libstark_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    import subprocess
    args = ["make", "-C", _HERE, "-j8", "-s"] + (["-B"] if force else [])
    subprocess.check_call(args)
    return lib_path


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(lib_path):
            raise StarkError(2, "libstark_b200.so not built (run `make -C stark-rs_b200` or __graft_entry__.build()); "
                                "there is no CPU fallback")
        _lib = C.CDLL(lib_path)
        _lib.stark_last_error.restype = C.c_char_p
        _lib.stark_version.restype = C.c_char_p
        _lib.stark_fri_verify_reason.restype = C.c_char_p
        _lib.stark_ctx_launches.restype = C.c_uint64
        _lib.stark_ctx_stream.restype = C.c_void_p
        _lib.stark_buf_ptr.restype = C.c_void_p
        _lib.stark_buf_len.restype = C.c_size_t
        _lib.stark_merkle_num_leaves.restype = C.c_size_t
        _lib.stark_merkle_num_levels.restype = C.c_uint32
        _lib.stark_fri_rounds.restype = C.c_uint32
        _lib.stark_merkle_nodes_ptr.restype = C.c_void_p
    return _lib


def _chk(rc):
    if rc != 0:
        msg = lib().stark_last_error().decode()
        raise (StarkPanic if rc == 1 else StarkError)(rc, msg)


def _u64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint64).reshape(-1))


def _p64(a):
    return a.ctypes.data_as(U64P)


def _p8(a):
    return a.ctypes.data_as(U8P)


def _b(x):
    if x is None:
        return np.zeros(0, dtype=np.uint8)
    if isinstance(x, np.ndarray):
        return np.ascontiguousarray(x, dtype=np.uint8).reshape(-1)
    return np.frombuffer(bytes(x), dtype=np.uint8).copy()


def prim_nth_root(n):
    """FiniteField::prim_nth_root (ff.rs:215-223)."""
    out = U64()
    _chk(lib().stark_ff_prim_nth_root(U64(n), C.byref(out)))
    return out.value


def fri_num_rounds(domain_length, expansion_factor, num_colinearity_tests):
    out = U32()
    _chk(lib().stark_fri_num_rounds(SZ(domain_length), U32(expansion_factor), U32(num_colinearity_tests), C.byref(out)))
    return out.value


def fri_proof_size(domain_length, expansion_factor, num_colinearity_tests):
    out = SZ()
    _chk(lib().stark_fri_proof_size(SZ(domain_length), U32(expansion_factor), U32(num_colinearity_tests), C.byref(out)))
    return out.value


def fri_sample_indices(seed, size, reduced_size, number):
    seed = _b(seed)
    out = np.zeros(max(number, 1), dtype=np.uint64)
    _chk(lib().stark_fri_sample_indices(_p8(seed), SZ(len(seed)), SZ(size), SZ(reduced_size), SZ(number), _p64(out)))
    return out[:number]


def fiat_shamir_challenge(transcript):
    """FiatShamir::challenge (fiat_shamir.rs:19-25): raw, unreduced u64 of a host-held transcript."""
    t = _b(transcript)
    out = U64()
    _chk(lib().stark_fiat_shamir_challenge(_p8(t), SZ(len(t)), C.byref(out)))
    return out.value


def hash_from_u64(value):
    """Hash::from_u64 (hash.rs:37-39), host side (the index seed of fri.rs:272)."""
    out = np.empty(32, dtype=np.uint8)
    _chk(lib().stark_hash_from_u64(U64(value), _p8(out)))
    return out.tobytes()


class Buffer:
    """stark_buf: a device array of uint32 field elements."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle

    def __len__(self):
        return lib().stark_buf_len(self.h)

    @property
    def ptr(self):
        return lib().stark_buf_ptr(self.h)

    def download(self, off=0, n=None):
        n = len(self) - off if n is None else n
        out = np.empty(n, dtype=np.uint64)
        _chk(lib().stark_buf_download(self.ctx.h, self.h, SZ(off), SZ(n), _p64(out)))
        return out

    def free(self):
        if self.h and self.ctx.h:      # a handle must not outlive its context (the library frees through it)
            lib().stark_buf_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class MerkleTree:
    """MerkleTree (merkle.rs:4-97) with every level resident on the device."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle

    @property
    def num_leaves(self):
        return lib().stark_merkle_num_leaves(self.h)

    @property
    def num_levels(self):
        return lib().stark_merkle_num_levels(self.h)

    def get_root(self):
        out = np.empty(32, dtype=np.uint8)
        _chk(lib().stark_merkle_root(self.h, _p8(out)))
        return out.tobytes()

    def level(self, l):
        out = np.empty((self.num_leaves >> l, 32), dtype=np.uint8)
        _chk(lib().stark_merkle_level(self.h, U32(l), _p8(out)))
        return out

    def nodes(self):
        return np.concatenate([self.level(l) for l in range(self.num_levels)])

    def open(self, index):
        out = np.empty((max(self.num_levels - 1, 1), 32), dtype=np.uint8)
        n = SZ()
        _chk(lib().stark_merkle_open(self.h, SZ(index), _p8(out), C.byref(n)))
        return out[: n.value].copy()

    def open_batch(self, indices):
        """MerkleTree::open for many leaves: (len(indices), log2 n, 32) uint8"""
        idx = _u64(indices)
        depth = self.num_levels - 1
        out = np.zeros((len(idx), depth, 32), dtype=np.uint8)
        _chk(lib().stark_merkle_open_batch(self.h, _p64(idx), SZ(len(idx)), _p8(out)))
        return out

    @property
    def nodes_ptr(self):
        """device address of the flattened node array (level l at hash offset 2n - 2(n >> l))"""
        return lib().stark_merkle_nodes_ptr(self.h)

    @property
    def root_ptr(self):
        return self.nodes_ptr + 32 * (2 * self.num_leaves - 2)

    def free(self):
        if self.h and self.ctx.h:      # a handle must not outlive its context (the library frees through it)
            lib().stark_merkle_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class FriState:
    """What Fri::commit (fri.rs:105-156) returns and leaves behind: codewords, trees, roots, challenges."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle

    @property
    def rounds(self):
        return lib().stark_fri_rounds(self.h)

    def roots(self):
        out = np.empty((self.rounds, 32), dtype=np.uint8)
        _chk(lib().stark_fri_roots(self.h, _p8(out)))
        return out

    def alphas(self):
        out = np.zeros(max(self.rounds, 1), dtype=np.uint64)
        _chk(lib().stark_fri_alphas(self.h, _p64(out)))
        return [int(x) for x in out[: max(self.rounds - 1, 0)]]

    def codeword(self, r):
        n = SZ()
        _chk(lib().stark_fri_codeword_len(self.h, U32(r), C.byref(n)))
        out = np.empty(n.value, dtype=np.uint64)
        _chk(lib().stark_fri_codeword(self.h, U32(r), _p64(out)))
        return out

    def open(self, r, index):
        out = np.empty((64, 32), dtype=np.uint8)
        n = SZ()
        _chk(lib().stark_fri_open(self.h, U32(r), SZ(index), _p8(out), C.byref(n)))
        return out[: n.value].copy()

    def free(self):
        if self.h and self.ctx.h:      # a handle must not outlive its context (the library frees through it)
            lib().stark_fri_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """stark_ctx: one CUDA device + one stream.  `stream` may be a raw cudaStream_t (int) to borrow."""

    def __init__(self, device=0, stream=None):
        self.h = C.c_void_p()
        if stream is None:
            _chk(lib().stark_ctx_create(C.c_int(device), C.byref(self.h)))
        else:
            _chk(lib().stark_ctx_create_on_stream(C.c_int(device), C.c_void_p(stream), C.byref(self.h)))
        self.device = device

    def close(self):
        if self.h:
            lib().stark_ctx_destroy(self.h)
            self.h = None

    def sync(self):
        _chk(lib().stark_ctx_sync(self.h))

    def make_current(self):
        """cudaSetDevice(this context's device) -- for a thread that uses contexts on several devices in turn"""
        _chk(lib().stark_ctx_make_current(self.h))
        return self

    @property
    def launches(self):
        return lib().stark_ctx_launches(self.h)

    @property
    def stream(self):
        return lib().stark_ctx_stream(self.h)

    # ---- measurement hooks
    def profile_begin(self):
        _chk(lib().stark_ctx_profile_begin(self.h))

    def profile_end(self):
        """per-kernel CUDA-event timings since profile_begin: [{kernel, launches, ms, bytes}]"""
        import json
        buf = C.create_string_buffer(1 << 16)
        _chk(lib().stark_ctx_profile_end(self.h, buf, SZ(len(buf))))
        return json.loads(buf.value.decode())

    def int_peak(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        _chk(lib().stark_bench_int_peak(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"imad_per_s": a.value, "alu_per_s": b.value, "mixed_per_s": c.value}

    def mul_peak(self):
        """rates of the pieces of a Montgomery product (thread-operations/s), stark_bench_mul_peak"""
        out = (C.c_double * 4)()
        _chk(lib().stark_bench_mul_peak(self.h, out))
        return {"imad_wide_per_s": out[0], "imad_hi_per_s": out[1], "mont_mul_per_s": out[2], "butterfly_step_per_s": out[3]}

    def hash_latency(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        _chk(lib().stark_bench_hash_latency(self.h, C.byref(a), C.byref(b), C.byref(c)))
        d = C.c_double()
        _chk(lib().stark_bench_hash_latency_hso(self.h, C.byref(d)))
        return {"hs_cycles": a.value, "hs2_cycles": b.value, "hsq_cycles": c.value, "hso_cycles": d.value}

    # ---- buffers
    def alloc(self, n):
        h = C.c_void_p()
        _chk(lib().stark_buf_alloc(self.h, SZ(n), C.byref(h)))
        return Buffer(self, h)

    def upload(self, values):
        v = _u64(values)
        h = C.c_void_p()
        _chk(lib().stark_buf_upload(self.h, _p64(v), SZ(len(v)), C.byref(h)))
        return Buffer(self, h)

    def upload_ptr(self, host_ptr, n):
        """upload n uint64 values from a raw (e.g. pinned) host pointer"""
        h = C.c_void_p()
        _chk(lib().stark_buf_upload(self.h, C.cast(host_ptr, U64P), SZ(n), C.byref(h)))
        return Buffer(self, h)

    def wrap(self, device_ptr, n):
        h = C.c_void_p()
        _chk(lib().stark_buf_wrap(self.h, C.c_void_p(device_ptr), SZ(n), C.byref(h)))
        return Buffer(self, h)

    # ---- ff.rs batch ops
    def _vec2(self, name, a, b):
        a, b = _u64(a), _u64(b)
        assert len(a) == len(b)
        out = np.empty_like(a)
        _chk(getattr(lib(), name)(self.h, _p64(a), _p64(b), _p64(out), SZ(len(a))))
        return out

    def ff_vec_add(self, a, b):
        return self._vec2("stark_ff_vec_add", a, b)

    def ff_vec_sub(self, a, b):
        return self._vec2("stark_ff_vec_sub", a, b)

    def ff_vec_mul(self, a, b):
        return self._vec2("stark_ff_vec_mul", a, b)

    def ff_vec_neg(self, a):
        a = _u64(a)
        out = np.empty_like(a)
        _chk(lib().stark_ff_vec_neg(self.h, _p64(a), _p64(out), SZ(len(a))))
        return out

    def ff_vec_inv(self, a):
        a = _u64(a)
        out = np.empty_like(a)
        _chk(lib().stark_ff_vec_inv(self.h, _p64(a), _p64(out), SZ(len(a))))
        return out

    def ff_vec_pow(self, a, e):
        a = _u64(a)
        out = np.empty_like(a)
        _chk(lib().stark_ff_vec_pow(self.h, _p64(a), U64(e), _p64(out), SZ(len(a))))
        return out

    # ---- univariate
    def poly_mul(self, a, b):
        """Polynomial::mul (mul.rs:6-29)"""
        a, b = _u64(a), _u64(b)
        out = np.zeros(max(len(a) + len(b), 1), dtype=np.uint64)
        n = SZ()
        _chk(lib().stark_poly_mul(self.h, _p64(a), SZ(len(a)), _p64(b), SZ(len(b)), _p64(out), C.byref(n)))
        return out[: n.value].copy()

    def poly_div(self, a, b):
        """Polynomial::div (div.rs:6-53) -> (quotient, remainder)"""
        a, b = _u64(a), _u64(b)
        q = np.zeros(max(len(a), 1), dtype=np.uint64)
        r = np.zeros(max(len(a) + len(b), 1), dtype=np.uint64)
        nq, nr = SZ(), SZ()
        _chk(lib().stark_poly_div(self.h, _p64(a), SZ(len(a)), _p64(b), SZ(len(b)), _p64(q), C.byref(nq), _p64(r),
                                  C.byref(nr)))
        return q[: nq.value].copy(), r[: nr.value].copy()

    def poly_eval_coset(self, coeffs, offset, log_n):
        c = _u64(coeffs)
        out = np.empty(1 << log_n, dtype=np.uint64)
        _chk(lib().stark_poly_eval_coset(self.h, _p64(c), SZ(len(c)), U64(offset), U32(log_n), _p64(out)))
        return out

    def poly_interpolate_coset(self, vals, offset, log_n):
        v = _u64(vals)
        assert len(v) == 1 << log_n
        out = np.empty(1 << log_n, dtype=np.uint64)
        n = SZ()
        _chk(lib().stark_poly_interpolate_coset(self.h, _p64(v), U64(offset), U32(log_n), _p64(out), C.byref(n)))
        return out[: n.value].copy()

    def poly_eval_domain(self, coeffs, domain):
        c, d = _u64(coeffs), _u64(domain)
        out = np.empty(len(d), dtype=np.uint64)
        _chk(lib().stark_poly_eval_domain(self.h, _p64(c), SZ(len(c)), _p64(d), SZ(len(d)), _p64(out)))
        return out

    def poly_interpolate_domain(self, domain, vals):
        d, v = _u64(domain), _u64(vals)
        assert len(d) == len(v)
        out = np.empty(max(len(d), 1), dtype=np.uint64)
        n = SZ()
        _chk(lib().stark_poly_interpolate_domain(self.h, _p64(d), _p64(v), SZ(len(d)), _p64(out), C.byref(n)))
        return out[: n.value].copy()

    def poly_scale(self, coeffs, factor):
        c = _u64(coeffs)
        out = np.empty_like(c)
        _chk(lib().stark_poly_scale(self.h, _p64(c), SZ(len(c)), U64(factor), _p64(out)))
        return out

    def poly_zerofier_coset(self, offset, log_n):
        out = np.empty((1 << log_n) + 1, dtype=np.uint64)
        _chk(lib().stark_poly_zerofier_coset(self.h, U64(offset), U32(log_n), _p64(out)))
        return out

    def poly_zerofier_domain(self, domain):
        d = _u64(domain)
        out = np.empty(len(d) + 1, dtype=np.uint64)
        _chk(lib().stark_poly_zerofier_domain(self.h, _p64(d), SZ(len(d)), _p64(out)))
        return out

    # ---- LDE
    def lde(self, cols, log_blowup, offset=3):
        """cols: (n_cols, n) array, one trace column per row (column-major storage)."""
        cols = np.ascontiguousarray(np.asarray(cols, dtype=np.uint64))
        if cols.ndim == 1:
            cols = cols[None, :]
        n_cols, n = cols.shape
        log_n = n.bit_length() - 1
        assert 1 << log_n == n
        out = np.empty((n_cols, n << log_blowup), dtype=np.uint64)
        _chk(lib().stark_lde(self.h, _p64(cols), U32(n_cols), U32(log_n), U32(log_blowup), U64(offset), _p64(out)))
        return out

    def lde_dev(self, cols_buf, n_cols, log_n, log_blowup, offset=3, out=None):
        out = out or self.alloc(n_cols << (log_n + log_blowup))
        _chk(lib().stark_lde_dev(self.h, cols_buf.h, U32(n_cols), U32(log_n), U32(log_blowup), U64(offset), out.h))
        return out

    def ntt_dev(self, in_buf, out_buf, log_n, batch=1, inverse=False):
        _chk(lib().stark_ntt_dev(self.h, in_buf.h, out_buf.h, U32(log_n), U32(batch), C.c_int(1 if inverse else 0)))
        return out_buf

    # ---- hash / merkle
    def hash_bytes(self, msgs, msg_len):
        m = _b(msgs)
        n = len(m) // msg_len if msg_len else 0
        out = np.empty((n, 32), dtype=np.uint8)
        _chk(lib().stark_hash_bytes(self.h, _p8(m), SZ(n), SZ(msg_len), _p8(out)))
        return out

    def hash_empty(self, n):
        out = np.empty((n, 32), dtype=np.uint8)
        m = np.zeros(1, dtype=np.uint8)
        _chk(lib().stark_hash_bytes(self.h, _p8(m), SZ(n), SZ(0), _p8(out)))
        return out

    def hash_leaves(self, vals, width=1):
        v = _u64(vals)
        n = len(v) // width
        out = np.empty((n, 32), dtype=np.uint8)
        _chk(lib().stark_hash_leaves(self.h, _p64(v), SZ(n), U32(width), _p8(out)))
        return out

    def merkle_build(self, leaves):
        l = np.ascontiguousarray(leaves, dtype=np.uint8).reshape(-1, 32)
        h = C.c_void_p()
        _chk(lib().stark_merkle_build(self.h, _p8(l), SZ(len(l)), C.byref(h)))
        return MerkleTree(self, h)

    def merkle_build_from_values(self, vals, width=1):
        v = _u64(vals)
        h = C.c_void_p()
        _chk(lib().stark_merkle_build_from_values(self.h, _p64(v), SZ(len(v) // width), U32(width), C.byref(h)))
        return MerkleTree(self, h)

    def merkle_build_from_buf(self, buf, n_leaves, width=1):
        h = C.c_void_p()
        _chk(lib().stark_merkle_build_from_buf(self.h, buf.h, SZ(n_leaves), U32(width), C.byref(h)))
        return MerkleTree(self, h)

    def merkle_build_dev(self, leaves_dev_ptr, n):
        """MerkleTree::new over n leaf hashes already on the device"""
        h = C.c_void_p()
        _chk(lib().stark_merkle_build_dev(self.h, C.c_void_p(leaves_dev_ptr), SZ(n), C.byref(h)))
        return MerkleTree(self, h)

    def merkle_commit(self, leaves):
        l = np.ascontiguousarray(leaves, dtype=np.uint8).reshape(-1, 32)
        out = np.empty(32, dtype=np.uint8)
        _chk(lib().stark_merkle_commit(self.h, _p8(l), SZ(len(l)), _p8(out)))
        return out.tobytes()

    # ---- fri
    def fri_fold(self, codeword, alpha_raw, offset, omega):
        cw = _u64(codeword)
        out = np.empty(len(cw) // 2, dtype=np.uint64)
        _chk(lib().stark_fri_fold(self.h, _p64(cw), SZ(len(cw)), U64(alpha_raw), U64(offset), U64(omega), _p64(out)))
        return out

    def fri_fold_dev(self, cw_buf, n, alpha_raw, offset, omega, out=None):
        out = out or self.alloc(n // 2)
        _chk(lib().stark_fri_fold_dev(self.h, cw_buf.h, SZ(n), U64(alpha_raw), U64(offset), U64(omega), out.h))
        return out

    def fri_fold_range_dev(self, cw_buf, n, alpha_raw, offset, omega, i0, count, out_buf, out_off=0):
        """outputs [i0, i0+count) of fold_codeword into out_buf[out_off:]"""
        _chk(lib().stark_fri_fold_range_dev(self.h, cw_buf.h, SZ(n), U64(alpha_raw), U64(offset), U64(omega), SZ(i0),
                                            SZ(count), out_buf.h, SZ(out_off)))
        return out_buf

    def fri_fold_bcast_dev(self, cw_buf, n, alpha_raw, offset, omega, i0, count, peer_ptrs, multicast_ptr=0):
        """outputs [i0, i0+count) of fold_codeword stored straight into every rank's replica (peer device addresses)"""
        arr = (C.c_void_p * max(len(peer_ptrs), 1))(*[C.c_void_p(int(p)) for p in peer_ptrs])
        _chk(lib().stark_fri_fold_bcast_dev(self.h, cw_buf.h, SZ(n), U64(alpha_raw), U64(offset), U64(omega), SZ(i0),
                                            SZ(count), arr, C.c_int(len(peer_ptrs)),
                                            C.c_void_p(int(multicast_ptr)) if multicast_ptr else None))

    def fri_commit(self, codeword, offset, omega, expansion_factor, num_colinearity_tests, transcript=b""):
        cw, t = _u64(codeword), _b(transcript)
        h = C.c_void_p()
        _chk(lib().stark_fri_commit(self.h, _p64(cw), SZ(len(cw)), U64(offset), U64(omega), U32(expansion_factor),
                                    U32(num_colinearity_tests), _p8(t), SZ(len(t)), C.byref(h)))
        return FriState(self, h)

    def fri_commit_dev(self, cw_buf, n, offset, omega, expansion_factor, num_colinearity_tests, transcript=b""):
        t = _b(transcript)
        h = C.c_void_p()
        _chk(lib().stark_fri_commit_dev(self.h, cw_buf.h, SZ(n), U64(offset), U64(omega), U32(expansion_factor),
                                        U32(num_colinearity_tests), _p8(t), SZ(len(t)), C.byref(h)))
        return FriState(self, h)

    def fri_prove(self, codeword, offset, omega, expansion_factor, num_colinearity_tests, transcript=b"",
                  domain_length=None):
        """Fri::prove + ProofStream::serialize.  Returns (proof_bytes, top_level_indices)."""
        cw, t = _u64(codeword), _b(transcript)
        dl = len(cw) if domain_length is None else domain_length
        ln = SZ()
        cap = 0
        try:
            cap = fri_proof_size(dl, expansion_factor, num_colinearity_tests)
        except StarkPanic:
            pass
        proof = np.empty(max(cap, 1), dtype=np.uint8)
        top = np.zeros(max(num_colinearity_tests, 1), dtype=np.uint64)
        _chk(lib().stark_fri_prove(self.h, _p64(cw), SZ(len(cw)), SZ(dl), U64(offset), U64(omega),
                                   U32(expansion_factor), U32(num_colinearity_tests), _p8(t), SZ(len(t)), _p8(proof),
                                   SZ(cap), C.byref(ln), _p64(top)))
        return proof[: ln.value].tobytes(), [int(x) for x in top[:num_colinearity_tests]]

    def fri_prove_dev(self, cw_buf, n, offset, omega, expansion_factor, num_colinearity_tests, proof_out=None,
                      transcript=b""):
        t = _b(transcript)
        cap = fri_proof_size(n, expansion_factor, num_colinearity_tests)
        proof = proof_out if proof_out is not None else np.empty(cap, dtype=np.uint8)
        ln = SZ()
        _chk(lib().stark_fri_prove_dev(self.h, cw_buf.h, SZ(n), SZ(n), U64(offset), U64(omega), U32(expansion_factor),
                                       U32(num_colinearity_tests), _p8(t), SZ(len(t)), _p8(proof), SZ(len(proof)),
                                       C.byref(ln), None))
        return proof[: ln.value]

    def fri_verify(self, proof, omega, offset, domain_length, expansion_factor, num_colinearity_tests, transcript=b"",
                   details=False):
        """Fri::verify (fri.rs:313-505) on the device.  Returns (ok, reason_text); with details=True also a dict with
        the roots popped from the stream, the sampled top-level indices and polynomial_values (fri.rs:437-441)."""
        b = np.frombuffer(bytes(proof), dtype=np.uint8) if len(proof) else np.zeros(0, dtype=np.uint8)
        t = _b(transcript)
        ok, why = C.c_int(0), U32(0)
        nq = num_colinearity_tests
        roots = np.zeros((64, 32), dtype=np.uint8)
        top = np.zeros(max(nq, 1), dtype=np.uint64)
        pidx = np.zeros(max(2 * nq, 1), dtype=np.uint64)
        pval = np.zeros(max(2 * nq, 1), dtype=np.uint64)
        _chk(lib().stark_fri_verify(self.h, _p8(b), SZ(len(b)), SZ(domain_length), U64(offset), U64(omega),
                                    U32(expansion_factor), U32(nq), _p8(t), SZ(len(t)), C.byref(ok), C.byref(why),
                                    _p8(roots), _p64(top), _p64(pidx), _p64(pval)))
        text = lib().stark_fri_verify_reason(why).decode()
        if not details:
            return bool(ok.value), text
        return bool(ok.value), text, {"roots": roots, "top": [int(x) for x in top[:nq]],
                                      "polynomial_values": [(int(i), int(v)) for i, v in zip(pidx[:2 * nq], pval[:2 * nq])]}

    # ---- pipeline (BASELINE config 3)
    def prove_trace(self, cols, log_blowup, offset=3, num_colinearity_tests=32):
        """LDE + per-column Merkle commit + Fri::prove on column 0.  Returns (column_roots, proof_bytes)."""
        cols = np.ascontiguousarray(np.asarray(cols, dtype=np.uint64))
        if cols.ndim == 1:
            cols = cols[None, :]
        n_cols, n = cols.shape
        log_n = n.bit_length() - 1
        cap = fri_proof_size(n << log_blowup, 1 << log_blowup, num_colinearity_tests)
        proof = np.empty(cap, dtype=np.uint8)
        roots = np.empty((n_cols, 32), dtype=np.uint8)
        ln = SZ()
        _chk(lib().stark_prove_trace(self.h, _p64(cols), U32(n_cols), U32(log_n), U32(log_blowup), U64(offset),
                                     U32(num_colinearity_tests), _p8(roots), _p8(proof), SZ(cap), C.byref(ln)))
        return roots, proof[: ln.value].tobytes()

    @staticmethod
    def _i128_rows(rows):
        """rows: sequence of rows of Python ints (i128 range) -> contiguous little-endian 16-byte values"""
        n_rows, n_cols = len(rows), len(rows[0])
        raw = bytearray(16 * n_rows * n_cols)
        k = 0
        for r in rows:
            assert len(r) == n_cols
            for v in r:
                raw[k:k + 16] = (int(v) & ((1 << 128) - 1)).to_bytes(16, "little")
                k += 16
        return np.frombuffer(bytes(raw), dtype=np.uint8), n_rows, n_cols

    def trace_to_columns(self, rows):
        """Trace::to_field_elements + get_col for every column (trace.rs:21-34) -> device Buffer, column-major"""
        raw, n_rows, n_cols = self._i128_rows(rows)
        h = C.c_void_p()
        _chk(lib().stark_trace_to_columns(self.h, raw.ctypes.data_as(C.c_void_p), SZ(n_rows), U32(n_cols), C.byref(h)))
        return Buffer(self, h)

    def prove_trace_rows(self, rows, log_blowup, offset=3, num_colinearity_tests=32):
        raw, n_rows, n_cols = self._i128_rows(rows)
        log_n = n_rows.bit_length() - 1
        assert 1 << log_n == n_rows
        cap = fri_proof_size(n_rows << log_blowup, 1 << log_blowup, num_colinearity_tests)
        proof = np.empty(cap, dtype=np.uint8)
        roots = np.empty((n_cols, 32), dtype=np.uint8)
        ln = SZ()
        _chk(lib().stark_prove_trace_rows(self.h, raw.ctypes.data_as(C.c_void_p), U32(n_cols), U32(log_n), U32(log_blowup),
                                          U64(offset), U32(num_colinearity_tests), _p8(roots), _p8(proof), SZ(cap),
                                          C.byref(ln)))
        return roots, proof[: ln.value].tobytes()

    def prove_trace_ptr(self, host_ptr, n_cols, log_n, log_blowup, offset, nq, roots, proof):
        """same, from a raw (pinned) host pointer into preallocated numpy outputs; returns proof length"""
        ln = SZ()
        _chk(lib().stark_prove_trace(self.h, C.cast(host_ptr, U64P), U32(n_cols), U32(log_n), U32(log_blowup),
                                     U64(offset), U32(nq), _p8(roots), _p8(proof), SZ(len(proof)), C.byref(ln)))
        return ln.value

    def prove_trace_dev(self, cols_buf, n_cols, log_n, log_blowup, offset, nq, roots, proof):
        ln = SZ()
        _chk(lib().stark_prove_trace_dev(self.h, cols_buf.h, U32(n_cols), U32(log_n), U32(log_blowup), U64(offset),
                                         U32(nq), _p8(roots), _p8(proof), SZ(len(proof)), C.byref(ln)))
        return ln.value


# ---------------------------------------------------------------------------------------------- groups of GPUs

def mgpu_columns_of_rank(rank, world, n_cols):
    """which trace columns (> 0) rank `rank` of a group of `world` commits in stark_mgpu_prove_trace (host logic only)"""
    out = (C.c_uint32 * max(n_cols, 1))()
    k = lib().stark_mgpu_columns_of_rank(C.c_int(rank), C.c_int(world), U32(n_cols), out)
    return [int(out[i]) for i in range(k)]


def mgpu_unique_id():
    """stark_mgpu_unique_id: 128 bytes rank 0 hands to the other ranks (any channel) before Group.init"""
    out = np.zeros(128, dtype=np.uint8)
    _chk(lib().stark_mgpu_unique_id(_p8(out)))
    return out.tobytes()


class Group:
    """The ranks of a GPU group THIS PROCESS drives (include/stark_b200.h, "groups of GPUs"): one handle after
    Group.init (one process per GPU, NCCL + CUDA IPC inside the library), all of them after Group.local (one host thread,
    several contexts).  Every operation is collective: all ranks of the group call it in the same order.  Per-rank
    results come back as lists with one entry per driven rank."""

    def __init__(self, handles, ctxs, local):
        self.h, self.ctxs, self.is_local = handles, ctxs, local
        self.n_here = len(handles)
        self._arr = (C.c_void_p * self.n_here)(*[x.value for x in handles])
        lib().stark_mgpu_bytes_sent.restype = C.c_uint64

    @classmethod
    def init(cls, ctx, unique_id, rank, world, max_codeword):
        h = C.c_void_p()
        uid = _b(unique_id)
        assert len(uid) == 128
        _chk(lib().stark_mgpu_init(ctx.h, _p8(uid), C.c_int(rank), C.c_int(world), SZ(max_codeword), C.byref(h)))
        return cls([h], [ctx], False)

    @classmethod
    def local(cls, ctxs, max_codeword):
        n = len(ctxs)
        arr = (C.c_void_p * n)(*[c.h.value for c in ctxs])
        out = (C.c_void_p * n)()
        _chk(lib().stark_mgpu_create_local(arr, C.c_int(n), SZ(max_codeword), out))
        return cls([C.c_void_p(out[i]) for i in range(n)], list(ctxs), True)

    def close(self):
        if self.h:
            if self.is_local:
                lib().stark_mgpu_destroy(self.h[0])
            else:
                for x in self.h:
                    lib().stark_mgpu_destroy(x)
            self.h = []

    @property
    def world(self):
        return lib().stark_mgpu_world(self.h[0])

    @property
    def ranks(self):
        return [lib().stark_mgpu_rank(x) for x in self.h]

    @property
    def bytes_sent(self):
        return [int(lib().stark_mgpu_bytes_sent(x)) for x in self.h]

    def set_shard_log(self, log_n):
        _chk(lib().stark_mgpu_set_shard_log(self.h[0], U32(log_n)))

    def barrier(self):
        if self.is_local:
            _chk(lib().stark_mgpu_barrier(self.h[0]))
        else:
            for x in self.h:
                _chk(lib().stark_mgpu_barrier(x))

    def owned_columns(self, n_cols):
        """per driven rank: the trace columns (> 0) it commits in prove_trace"""
        res = []
        for x in self.h:
            out = (C.c_uint32 * max(n_cols, 1))()
            k = lib().stark_mgpu_owned_columns(x, U32(n_cols), out)
            res.append([int(out[i]) for i in range(k)])
        return res

    # per-rank output arrays -> C arrays of pointers
    def _outs(self, arrays):
        return (C.c_void_p * self.n_here)(*[a.ctypes.data for a in arrays])

    def _bufs(self, bufs):
        return (C.c_void_p * len(bufs))(*[b.h.value for b in bufs])

    def prove_trace(self, cols, log_blowup, offset=3, num_colinearity_tests=32):
        """stark_mgpu_prove_trace: cols = the WHOLE trace (n_cols, n).  -> [(column_roots, proof bytes)] per driven rank"""
        cols = np.ascontiguousarray(np.asarray(cols, dtype=np.uint64))
        if cols.ndim == 1:
            cols = cols[None, :]
        n_cols, n = cols.shape
        log_n = n.bit_length() - 1
        cap = fri_proof_size(n << log_blowup, 1 << log_blowup, num_colinearity_tests)
        proofs = [np.zeros(cap, dtype=np.uint8) for _ in self.h]
        roots = [np.zeros((n_cols, 32), dtype=np.uint8) for _ in self.h]
        ln = SZ()
        _chk(lib().stark_mgpu_prove_trace(self._arr, C.c_int(self.n_here), _p64(cols), U32(n_cols), U32(log_n), U32(log_blowup),
                                          U64(offset), U32(num_colinearity_tests), self._outs(roots), self._outs(proofs),
                                          SZ(cap), C.byref(ln)))
        return [(r, p[: ln.value].tobytes()) for r, p in zip(roots, proofs)]

    def prove_trace_ptr(self, host_ptr, n_cols, log_n, log_blowup, offset, nq, roots, proofs):
        """same from a raw (pinned) host pointer to the whole trace into preallocated per-rank numpy outputs"""
        ln = SZ()
        _chk(lib().stark_mgpu_prove_trace(self._arr, C.c_int(self.n_here), C.cast(host_ptr, U64P), U32(n_cols), U32(log_n),
                                          U32(log_blowup), U64(offset), U32(nq), self._outs(roots), self._outs(proofs),
                                          SZ(len(proofs[0])), C.byref(ln)))
        return ln.value

    def prove_trace_dev(self, my_cols, n_cols, log_n, log_blowup, offset, nq, roots, proofs):
        """my_cols[k]: Buffer with column 0 followed by rank k's owned columns"""
        ln = SZ()
        _chk(lib().stark_mgpu_prove_trace_dev(self._arr, C.c_int(self.n_here), self._bufs(my_cols), U32(n_cols), U32(log_n),
                                              U32(log_blowup), U64(offset), U32(nq), self._outs(roots), self._outs(proofs),
                                              SZ(len(proofs[0])), C.byref(ln)))
        return ln.value

    def fri_prove_dev(self, codewords, n, offset, omega, expansion_factor, num_colinearity_tests, transcript=b"",
                      domain_length=None):
        """-> [(proof bytes, top indices)] per driven rank"""
        t = _b(transcript)
        cap = fri_proof_size(n, expansion_factor, num_colinearity_tests)
        proofs = [np.zeros(cap, dtype=np.uint8) for _ in self.h]
        tops = [np.zeros(max(num_colinearity_tests, 1), dtype=np.uint64) for _ in self.h]
        ln = SZ()
        _chk(lib().stark_mgpu_fri_prove_dev(self._arr, C.c_int(self.n_here), self._bufs(codewords), SZ(n),
                                            SZ(n if domain_length is None else domain_length), U64(offset), U64(omega),
                                            U32(expansion_factor), U32(num_colinearity_tests), _p8(t), SZ(len(t)),
                                            self._outs(proofs), SZ(cap), C.byref(ln), self._outs(tops)))
        return [(p[: ln.value].tobytes(), [int(x) for x in tp[:num_colinearity_tests]]) for p, tp in zip(proofs, tops)]

    def fold_commit_round(self, codewords, n, offset, omega):
        """BASELINE config 5 round -> [(root bytes, alpha_raw, folded Buffer view)] per driven rank"""
        roots = [np.zeros(32, dtype=np.uint8) for _ in self.h]
        alphas = (C.c_uint64 * self.n_here)()
        folded = (C.c_void_p * self.n_here)()
        _chk(lib().stark_mgpu_fold_commit_round(self._arr, C.c_int(self.n_here), self._bufs(codewords), SZ(n), U64(offset),
                                                U64(omega), self._outs(roots), alphas, folded))
        return [(roots[k].tobytes(), int(alphas[k]), Buffer(self.ctxs[k], C.c_void_p(folded[k]))) for k in range(self.n_here)]

    def lde_commit(self, cols, n_groups, group_width, log_n, log_blowup, offset=3):
        """BASELINE config 4 from the whole host trace (n_groups * group_width columns of 2^log_n rows, column-major)
        -> [(group roots (n_groups, 32), commitment bytes)] per driven rank"""
        cols = _u64(cols)
        assert len(cols) == (n_groups * group_width) << log_n
        roots = [np.zeros((n_groups, 32), dtype=np.uint8) for _ in self.h]
        com = [np.zeros(32, dtype=np.uint8) for _ in self.h]
        _chk(lib().stark_mgpu_lde_commit(self._arr, C.c_int(self.n_here), _p64(cols), U32(n_groups), U32(group_width),
                                         U32(log_n), U32(log_blowup), U64(offset), self._outs(roots), self._outs(com)))
        return [(r, c.tobytes()) for r, c in zip(roots, com)]

    def lde_commit_dev(self, owned_groups, n_groups, group_width, log_n, log_blowup, offset=3):
        """owned_groups: rank-major flat list of Buffers (n_groups / world per driven rank)"""
        roots = [np.zeros((n_groups, 32), dtype=np.uint8) for _ in self.h]
        com = [np.zeros(32, dtype=np.uint8) for _ in self.h]
        _chk(lib().stark_mgpu_lde_commit_dev(self._arr, C.c_int(self.n_here), self._bufs(owned_groups), U32(n_groups),
                                             U32(group_width), U32(log_n), U32(log_blowup), U64(offset), self._outs(roots),
                                             self._outs(com)))
        return [(r, c.tobytes()) for r, c in zip(roots, com)]
