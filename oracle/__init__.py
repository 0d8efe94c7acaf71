"""ctypes front end of the CPU oracle (oracle/stark_oracle.c, oracle/fast_cpu.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  The product package (stark-rs_b200/) never imports it.

A reference panic surfaces as OraclePanic(message) with the reference's own message text
(ff.rs:171 "no inverse", merkle.rs:12 "Cannot create tree from empty leaves", ...).
"""
import ctypes as C
import os
import subprocess

import numpy as np

P = 998244353
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")


class OraclePanic(Exception):
    pass


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("stark_oracle.c", "fast_cpu.c", "bench_cpu.c")]
    if (not force and os.path.exists(_LIB)
            and all(not os.path.exists(s) or os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in srcs)):
        return _LIB
    subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "CC=gcc"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.oracle_last_error.restype = C.c_char_p
    return _lib


def _chk(rc):
    if rc != 0:
        raise OraclePanic(lib().oracle_last_error().decode())


def _u64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint64).reshape(-1))


def _p64(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _p8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _bytes(b):
    return np.frombuffer(bytes(b), dtype=np.uint8).copy() if not isinstance(b, np.ndarray) else np.ascontiguousarray(b, dtype=np.uint8)


U64, SZ = C.c_uint64, C.c_size_t

# ------------------------------------------------------------------ ff.rs


def _scalar2(name):
    def f(a, b, p=P):
        out = U64()
        _chk(getattr(lib(), name)(U64(p), U64(a), U64(b), C.byref(out)))
        return out.value
    return f


ff_mul = _scalar2("oracle_ff_mul")
ff_add = _scalar2("oracle_ff_add")
ff_sub = _scalar2("oracle_ff_sub")
ff_div = _scalar2("oracle_ff_div")
ff_exp = _scalar2("oracle_ff_exp")


def ff_neg(a, p=P):
    out = U64()
    _chk(lib().oracle_ff_neg(U64(p), U64(a), C.byref(out)))
    return out.value


def ff_inv(a, p=P):
    out = U64()
    _chk(lib().oracle_ff_inv(U64(p), U64(a), C.byref(out)))
    return out.value


def ff_g(p=P):
    out = U64()
    _chk(lib().oracle_ff_g(U64(p), C.byref(out)))
    return out.value


def ff_prim_nth_root(n, p=P):
    out = U64()
    _chk(lib().oracle_ff_prim_nth_root(U64(p), U64(n), C.byref(out)))
    return out.value


def ff_sample(salt, p=P):
    b = _bytes(salt)
    out = U64()
    _chk(lib().oracle_ff_sample(U64(p), _p8(b), SZ(len(b)), C.byref(out)))
    return out.value


_OPS = {"add": 0, "sub": 1, "mul": 2, "neg": 3, "inv": 4, "pow": 5, "div": 6}


def ff_vec(op, a, b=None, e=0, p=P):
    a = _u64(a)
    b = _u64(b) if b is not None else a
    out = np.empty_like(a)
    _chk(lib().oracle_ff_vec(U64(p), C.c_int(_OPS[op]), _p64(a), _p64(b), U64(e), _p64(out), SZ(len(a))))
    return out

# ---------------------------------------------------------- univariate/*


def _poly2(name):
    def f(a, b, p=P):
        a, b = _u64(a), _u64(b)
        cap = len(a) + len(b) + 2
        out = np.zeros(cap, dtype=np.uint64)
        n = SZ()
        _chk(getattr(lib(), name)(U64(p), _p64(a), SZ(len(a)), _p64(b), SZ(len(b)), _p64(out), SZ(cap), C.byref(n)))
        return out[: n.value].copy()
    return f


poly_add = _poly2("oracle_poly_add")
poly_sub = _poly2("oracle_poly_sub")
poly_mul = _poly2("oracle_poly_mul")


def poly_deg(a):
    a = _u64(a)
    out = C.c_int64()
    _chk(lib().oracle_poly_deg(_p64(a), SZ(len(a)), C.byref(out)))
    return out.value


def poly_div(a, b, p=P):
    a, b = _u64(a), _u64(b)
    cap = len(a) + len(b) + 2
    q, r = np.zeros(cap, dtype=np.uint64), np.zeros(cap, dtype=np.uint64)
    nq, nr = SZ(), SZ()
    _chk(lib().oracle_poly_div(U64(p), _p64(a), SZ(len(a)), _p64(b), SZ(len(b)), _p64(q), SZ(cap), C.byref(nq),
                               _p64(r), SZ(cap), C.byref(nr)))
    return q[: nq.value].copy(), r[: nr.value].copy()


def poly_exp(a, e, p=P):
    a = _u64(a)
    cap = max(1, len(a)) * max(1, e) + 2
    out = np.zeros(cap, dtype=np.uint64)
    n = SZ()
    _chk(lib().oracle_poly_exp(U64(p), _p64(a), SZ(len(a)), U64(e), _p64(out), SZ(cap), C.byref(n)))
    return out[: n.value].copy()


def poly_eval(a, x, p=P):
    a = _u64(a)
    out = U64()
    _chk(lib().oracle_poly_eval(U64(p), _p64(a), SZ(len(a)), U64(x), C.byref(out)))
    return out.value


def poly_eval_domain(a, dom, p=P):
    a, dom = _u64(a), _u64(dom)
    out = np.empty(len(dom), dtype=np.uint64)
    _chk(lib().oracle_poly_eval_domain(U64(p), _p64(a), SZ(len(a)), _p64(dom), SZ(len(dom)), _p64(out)))
    return out


def poly_interpolate_domain(dom, vals, p=P):
    dom, vals = _u64(dom), _u64(vals)
    assert len(dom) == len(vals)
    cap = len(dom) + 2
    out = np.zeros(cap, dtype=np.uint64)
    n = SZ()
    _chk(lib().oracle_poly_interpolate_domain(U64(p), _p64(dom), _p64(vals), SZ(len(dom)), _p64(out), SZ(cap),
                                              C.byref(n)))
    return out[: n.value].copy()


def poly_zerofier(dom, p=P):
    dom = _u64(dom)
    cap = len(dom) + 2
    out = np.zeros(cap, dtype=np.uint64)
    n = SZ()
    _chk(lib().oracle_poly_zerofier(U64(p), _p64(dom), SZ(len(dom)), _p64(out), SZ(cap), C.byref(n)))
    return out[: n.value].copy()


def poly_scale(a, factor, p=P):
    a = _u64(a)
    out = np.zeros(len(a), dtype=np.uint64)
    _chk(lib().oracle_poly_scale(U64(p), _p64(a), SZ(len(a)), U64(factor), _p64(out)))
    return out


def poly_test_colinearity(xs, ys, p=P):
    xs, ys = _u64(xs), _u64(ys)
    out = C.c_int()
    _chk(lib().oracle_poly_test_colinearity(U64(p), _p64(xs), _p64(ys), SZ(len(xs)), C.byref(out)))
    return bool(out.value)


def lde(col, blowup, offset=3, p=P):
    """Reference-algorithm LDE (O(n^3)); small n only."""
    col = _u64(col)
    out = np.empty(len(col) * blowup, dtype=np.uint64)
    _chk(lib().oracle_lde(U64(p), _p64(col), SZ(len(col)), SZ(blowup), U64(offset), _p64(out)))
    return out

# ------------------------------------------------------------ hash / merkle


def hash_from_bytes(b):
    b = _bytes(b)
    out = np.empty(32, dtype=np.uint8)
    _chk(lib().oracle_hash_from_bytes(_p8(b), SZ(len(b)), _p8(out)))
    return out.tobytes()


def hash_from_field_elements(e):
    e = _u64(e)
    out = np.empty(32, dtype=np.uint8)
    _chk(lib().oracle_hash_from_field_elements(_p64(e), SZ(len(e)), _p8(out)))
    return out.tobytes()


def hash_from_u64(v):
    return hash_from_field_elements([v])


def hash_combine(l, r):
    l, r = _bytes(l), _bytes(r)
    out = np.empty(32, dtype=np.uint8)
    _chk(lib().oracle_hash_combine(_p8(l), _p8(r), _p8(out)))
    return out.tobytes()


def sbox(b):
    out = C.c_uint8()
    _chk(lib().oracle_sbox(C.c_uint8(b), C.byref(out)))
    return out.value


def hash_leaves(vals, width=1):
    vals = _u64(vals)
    n = len(vals) // width
    out = np.empty((n, 32), dtype=np.uint8)
    _chk(lib().oracle_hash_leaves(_p64(vals), SZ(n), SZ(width), _p8(out)))
    return out


def merkle_build(leaves):
    """All levels concatenated: (2n-1, 32) uint8 -- leaves first, root last (merkle.rs:18-29)."""
    leaves = np.ascontiguousarray(leaves, dtype=np.uint8).reshape(-1, 32)
    n = len(leaves)
    out = np.empty((max(2 * n - 1, 0), 32), dtype=np.uint8)
    _chk(lib().oracle_merkle_build(_p8(leaves), SZ(n), _p8(out)))
    return out


def merkle_commit(leaves):
    leaves = np.ascontiguousarray(leaves, dtype=np.uint8).reshape(-1, 32)
    out = np.empty(32, dtype=np.uint8)
    _chk(lib().oracle_merkle_commit(_p8(leaves), SZ(len(leaves)), _p8(out)))
    return out.tobytes()


def merkle_open(leaves, index):
    leaves = np.ascontiguousarray(leaves, dtype=np.uint8).reshape(-1, 32)
    out = np.empty((64, 32), dtype=np.uint8)
    n = SZ()
    _chk(lib().oracle_merkle_open(_p8(leaves), SZ(len(leaves)), SZ(index), _p8(out), C.byref(n)))
    return out[: n.value].copy()


def merkle_verify(leaf, index, proof, root):
    leaf, root = _bytes(leaf), _bytes(root)
    proof = np.ascontiguousarray(proof, dtype=np.uint8).reshape(-1, 32)
    ok = C.c_int()
    _chk(lib().oracle_merkle_verify(_p8(leaf), SZ(index), _p8(proof), SZ(len(proof)), _p8(root), C.byref(ok)))
    return bool(ok.value)

# ------------------------------------------------------- fiat-shamir / fri


def fs_challenge(transcript):
    t = _bytes(transcript)
    out = U64()
    _chk(lib().oracle_fs_challenge(_p8(t), SZ(len(t)), C.byref(out)))
    return out.value


def fri_num_rounds(n, ef, nq, omega=1, offset=1, p=P):
    out = U64()
    _chk(lib().oracle_fri_num_rounds(U64(p), U64(omega), U64(offset), SZ(n), SZ(ef), SZ(nq), C.byref(out)))
    return out.value


def fri_fold(cw, alpha, offset, omega, p=P):
    cw = _u64(cw)
    out = np.empty(len(cw) // 2, dtype=np.uint64)
    _chk(lib().oracle_fri_fold(U64(p), _p64(cw), SZ(len(cw)), U64(alpha), U64(offset), U64(omega), _p64(out)))
    return out


def fri_sample_indices(seed, size, reduced, number):
    seed = _bytes(seed)
    out = np.empty(number, dtype=np.uint64)
    _chk(lib().oracle_fri_sample_indices(_p8(seed), SZ(len(seed)), SZ(size), SZ(reduced), SZ(number), _p64(out)))
    return out


def fri_prove(cw, omega, offset, ef, nq, p=P, domain_length=None):
    """Fri::prove + ProofStream::serialize.  Returns dict(proof, top_indices, alphas, seed_challenge)."""
    cw = _u64(cw)
    n = len(cw)
    dl = n if domain_length is None else domain_length
    rounds = 0
    try:
        rounds = fri_num_rounds(dl, ef, nq)
    except OraclePanic:
        pass
    top = np.zeros(max(nq, 1), dtype=np.uint64)
    alphas = np.zeros(max(rounds, 1), dtype=np.uint64)
    seed = U64()
    ln = SZ()
    cap = 1 << 16
    while True:
        out = np.empty(cap, dtype=np.uint8)
        _chk(lib().oracle_fri_prove(U64(p), _p64(cw), SZ(n), SZ(dl), U64(omega), U64(offset), SZ(ef), SZ(nq), _p8(out), SZ(cap),
                                    C.byref(ln), _p64(top), _p64(alphas), C.byref(seed)))
        if ln.value <= cap:
            break
        cap = ln.value
    return dict(proof=out[: ln.value].tobytes(), top_indices=[int(x) for x in top[:nq]],
                alphas=[int(x) for x in alphas[: max(rounds - 1, 0)]], seed_challenge=seed.value, rounds=rounds)


def fri_verify(proof, omega, offset, n, ef, nq, p=P):
    b = _bytes(proof)
    ok = C.c_int()
    why = C.create_string_buffer(128)
    _chk(lib().oracle_fri_verify(U64(p), _p8(b), SZ(len(b)), U64(omega), U64(offset), SZ(n), SZ(ef), SZ(nq),
                                 C.byref(ok), why, SZ(128)))
    return bool(ok.value), why.value.decode()


def stream_count(proof):
    b = _bytes(proof)
    n = SZ()
    _chk(lib().oracle_stream_count(_p8(b), SZ(len(b)), C.byref(n)))
    return n.value


def trace_fibonacci(length):
    out = np.empty(length, dtype=np.uint64)
    _chk(lib().oracle_trace_fibonacci(SZ(length), _p64(out)))
    return out


def trace_columns(rows):
    """Trace::to_field_elements + get_col (trace.rs:21-34): rows of Python ints (i128) -> (n_cols, n_rows) raw u64 casts"""
    n_rows, n_cols = len(rows), len(rows[0])
    raw = b"".join((int(v) & ((1 << 128) - 1)).to_bytes(16, "little") for r in rows for v in r)
    buf = np.frombuffer(raw, dtype=np.uint8)
    out = np.empty((n_cols, n_rows), dtype=np.uint64)
    _chk(lib().oracle_trace_columns(_p8(buf), SZ(n_rows), SZ(n_cols), _p64(out)))
    return out

# --------------------------------------------- fast_cpu.c (algorithm-matched, not the reference)


def fast_eval_coset(coeffs, offset, log_n):
    c = _u64(coeffs)
    out = np.empty(1 << log_n, dtype=np.uint64)
    assert lib().fast_eval_coset(_p64(c), SZ(len(c)), U64(offset), C.c_uint32(log_n), _p64(out)) == 0
    return out


def fast_interpolate_coset(vals, offset, log_n):
    v = _u64(vals)
    assert len(v) == 1 << log_n
    out = np.empty(1 << log_n, dtype=np.uint64)
    assert lib().fast_interpolate_coset(_p64(v), U64(offset), C.c_uint32(log_n), _p64(out)) == 0
    return out


def fast_lde(col, log_n, log_blowup, offset=3):
    v = _u64(col)
    assert len(v) == 1 << log_n
    out = np.empty(1 << (log_n + log_blowup), dtype=np.uint64)
    assert lib().fast_lde(_p64(v), C.c_uint32(log_n), C.c_uint32(log_blowup), U64(offset), _p64(out)) == 0
    return out


def fast_poly_mul(a, b):
    a, b = _u64(a), _u64(b)
    out = np.empty(len(a) + len(b) - 1, dtype=np.uint64)
    assert lib().fast_poly_mul(_p64(a), SZ(len(a)), _p64(b), SZ(len(b)), _p64(out)) == 0
    return out


def fast_fri_fold(cw, alpha, offset, omega):
    cw = _u64(cw)
    out = np.empty(len(cw) // 2, dtype=np.uint64)
    assert lib().fast_fri_fold(_p64(cw), SZ(len(cw)), U64(alpha), U64(offset), U64(omega), _p64(out)) == 0
    return out


def set_threads(n):
    """threads used inside leaf hashing, Merkle levels and the fold (oracle_set_threads); returns the previous value"""
    return lib().oracle_set_threads(C.c_int(int(n)))


def splitmix64(seed, n, p=P):
    """SURVEY 8(d) deterministic input generator: element = next() mod p."""
    out = np.empty(n, dtype=np.uint64)
    state = np.uint64(seed)
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = state + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    out[:] = z % np.uint64(p)
    return out


def bench_pipeline(mode, log_n, log_blowup, nq, seed, threads):
    """bench_cpu.c: `threads` independent config-3 pipelines; returns (seconds, sha-less digest of thread 0's proof).
    mode 0 = reference algorithms end to end ("port"), mode 1 = O(n log n) LDE + reference hash/Merkle/FRI."""
    sec = C.c_double()
    dig = np.zeros(32, dtype=np.uint8)
    rc = lib().oracle_bench_pipeline(C.c_int(mode), C.c_uint32(log_n), C.c_uint32(log_blowup), C.c_uint32(nq), U64(seed),
                                     C.c_int(threads), C.byref(sec), _p8(dig))
    if rc != 0:
        raise OraclePanic("bench pipeline failed: " + lib().oracle_last_error().decode())
    return sec.value, dig.tobytes()
