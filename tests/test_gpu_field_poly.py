"""GPU parity: ff.rs batch ops, NTT-based univariate ops and the LDE, through the C ABI, vs the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P = 998244353


def rf(seed, n):
    return np.random.default_rng(seed).integers(0, P, n, dtype=np.uint64)


EDGE = np.array([0, 1, 2, 3, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, 1 << 23, (1 << 30) - 1 - 75497470],
                dtype=np.uint64) % P


def test_ff_vec_ops(ctx, oracle):
    a = np.concatenate([EDGE, rf(1, 5000)])
    b = np.concatenate([EDGE[::-1], rf(2, 5000)])
    assert np.array_equal(ctx.ff_vec_add(a, b), oracle.ff_vec("add", a, b))      # ff.rs:146-152
    assert np.array_equal(ctx.ff_vec_sub(a, b), oracle.ff_vec("sub", a, b))      # ff.rs:154-160
    assert np.array_equal(ctx.ff_vec_mul(a, b), oracle.ff_vec("mul", a, b))      # ff.rs:138-144
    assert np.array_equal(ctx.ff_vec_neg(a), oracle.ff_vec("neg", a))            # ff.rs:162-167
    for e in (0, 1, 2, 10, P - 2, P - 1, (1 << 64) - 1):
        assert np.array_equal(ctx.ff_vec_pow(a, e), oracle.ff_vec("pow", a, e=e))  # ff.rs:200-213
    nz = a[a != 0]
    assert np.array_equal(ctx.ff_vec_inv(nz), oracle.ff_vec("inv", nz))          # ff.rs:169-178
    for n in (1, 7, 8, 9, 1025):
        x = rf(n, n) + (rf(n, n) == 0)
        assert np.array_equal(ctx.ff_vec_inv(x), oracle.ff_vec("inv", x))


def test_ff_reference_kats(ctx):
    # ff.rs:344-505 known answers, batched
    a = np.array([100, P - 1, 200, 5, 0, 123, 1000000], dtype=np.uint64)
    b = np.array([200, 5, 100, 10, 123, 456, 2000000], dtype=np.uint64)
    assert list(ctx.ff_vec_add(a[:2], b[:2])) == [300, 4]
    assert list(ctx.ff_vec_sub(a[2:5], b[2:5])) == [100, P - 5, P - 123]
    assert list(ctx.ff_vec_mul(a[5:], b[5:])) == [123 * 456 % P, 2000000000000 % P]
    assert list(ctx.ff_vec_neg(np.array([100, 0], dtype=np.uint64))) == [P - 100, 0]
    assert list(ctx.ff_vec_pow(np.array([3, 12345, 2], dtype=np.uint64), 2))[0] == 9


def test_ff_errors(ctx, S):
    with pytest.raises(S.StarkPanic, match="no inverse"):                          # ff.rs:171, test ff.rs:552
        ctx.ff_vec_inv(np.array([5, 0, 7], dtype=np.uint64))
    with pytest.raises(S.StarkPanic, match="non-canonical"):
        ctx.ff_vec_add(np.array([P], dtype=np.uint64), np.array([1], dtype=np.uint64))
    assert len(ctx.ff_vec_add([], [])) == 0


@pytest.mark.parametrize("log_n", list(range(0, 9)))
def test_eval_interpolate_coset_vs_reference_algorithm(ctx, oracle, log_n):
    """eval.rs:16-21 / interpolate.rs:6-44 (Horner / O(n^3) Lagrange) on the coset of fri.rs:575-578"""
    n = 1 << log_n
    w = oracle.ff_prim_nth_root(n)
    for offset in (1, 3, 17):
        dom = [oracle.ff_mul(offset, oracle.ff_exp(w, i)) for i in range(n)]
        coeffs = rf(log_n * 7 + offset, n)
        ev = oracle.poly_eval_domain(coeffs, dom)
        assert np.array_equal(ctx.poly_eval_coset(coeffs, offset, log_n), ev)
        nc = max(1, n // 2 - 1)
        assert np.array_equal(ctx.poly_eval_coset(coeffs[:nc], offset, log_n), oracle.poly_eval_domain(coeffs[:nc], dom))
        if log_n <= 7:
            assert np.array_equal(ctx.poly_interpolate_coset(ev, offset, log_n), oracle.poly_interpolate_domain(dom, ev))


@pytest.mark.parametrize("log_n", [9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20])
def test_eval_interpolate_coset_vs_fast_cpu(ctx, oracle, log_n):
    n = 1 << log_n
    coeffs = rf(log_n, n)
    ev = ctx.poly_eval_coset(coeffs, 3, log_n)
    assert np.array_equal(ev, oracle.fast_eval_coset(coeffs, 3, log_n))
    assert np.array_equal(ctx.poly_interpolate_coset(ev, 3, log_n), coeffs)
    nc = n // 4 + 5
    assert np.array_equal(ctx.poly_eval_coset(coeffs[:nc], 7, log_n), oracle.fast_eval_coset(coeffs[:nc], 7, log_n))
    vals = rf(log_n + 100, n)
    assert np.array_equal(ctx.poly_interpolate_coset(vals, 1, log_n), oracle.fast_interpolate_coset(vals, 1, log_n))


@pytest.mark.parametrize("log_n", [21, 22, 23])
def test_ntt_large_roundtrip_and_spot(ctx, oracle, log_n):
    """maximum sizes (ff.rs:218): round trip + point evaluations checked with the reference's Horner eval"""
    n = 1 << log_n
    coeffs = rf(log_n, n)
    ev = ctx.poly_eval_coset(coeffs, 3, log_n)
    w = oracle.ff_prim_nth_root(n)
    for i in (0, 1, n // 2, n - 1, 12345):
        x = oracle.ff_mul(3, oracle.ff_exp(w, i))
        assert int(ev[i]) == oracle.poly_eval(coeffs, x)
    assert np.array_equal(ctx.poly_interpolate_coset(ev, 3, log_n), coeffs)


def test_coset_shape_rules_and_errors(ctx, S):
    assert len(ctx.poly_interpolate_coset(np.zeros(8, dtype=np.uint64), 3, 3)) == 0
    assert list(ctx.poly_interpolate_coset(np.zeros(1, dtype=np.uint64), 3, 0)) == [0]
    assert len(ctx.poly_interpolate_coset(np.array([0, 0, 5, 0], dtype=np.uint64), 3, 2)) == 4
    with pytest.raises(S.StarkPanic, match="n > 2\\^23 not supported"):                  # ff.rs:218
        ctx.poly_eval_coset([1, 2], 3, 24)
    with pytest.raises(S.StarkPanic, match="no inverse"):
        ctx.poly_interpolate_coset(np.ones(4, dtype=np.uint64), 0, 2)
    assert list(ctx.poly_eval_coset([], 3, 2)) == [0, 0, 0, 0]


def test_poly_mul(ctx, oracle):
    assert list(ctx.poly_mul([1, 1], [1, 1])) == [1, 2, 1]                               # mul.rs:89-101
    assert list(ctx.poly_mul([1, 0, 2], [3, 0, 4])) == [3, 0, 10, 0, 8]                  # mul.rs:104-119
    assert list(ctx.poly_mul([P - 1], [2])) == [P - 2]                                   # mul.rs:182-195
    assert len(ctx.poly_mul([], [1, 2])) == 0 and len(ctx.poly_mul([0, 0], [1, 2])) == 0  # mul.rs:7-12
    assert len(ctx.poly_mul([1, 2], [0])) == 0
    assert list(ctx.poly_mul([1, 0], [1, 0, 0])) == [1, 0, 0, 0]                         # mul.rs:14 length rule
    for na, nb in [(1, 1), (1, 9), (3, 5), (64, 64), (100, 29), (257, 255), (512, 513)]:
        a, b = rf(na, na), rf(nb + 1000, nb)
        assert np.array_equal(ctx.poly_mul(a, b), oracle.poly_mul(a, b)), (na, nb)       # schoolbook mul.rs:16-24
    a, b = rf(5, 1 << 16), rf(6, 1 << 16)                                                # BASELINE config 2 size
    assert np.array_equal(ctx.poly_mul(a, b), oracle.fast_poly_mul(a, b))
    a, b = rf(7, (1 << 22) + 1), rf(8, 1 << 22)
    assert len(ctx.poly_mul(a, b)) == (1 << 23)


def test_poly_eval_domain_arbitrary(ctx, oracle):
    assert list(ctx.poly_eval_domain([1, 2, 3, 4], [2])) == [49]                         # eval.rs:83-95
    assert list(ctx.poly_eval_domain([1, 1], [0, 1, 2, 3])) == [1, 2, 3, 4]              # eval.rs:98-118
    for nc, m in [(0, 5), (1, 1), (7, 300), (1500, 700), (3000, 5)]:
        c, d = rf(nc, nc), rf(m + 1, m)
        d[:2] = d[-1]                                                                    # duplicates are legal
        assert np.array_equal(ctx.poly_eval_domain(c, d), oracle.poly_eval_domain(c, d))


def test_poly_interpolate_domain_arbitrary(ctx, oracle, S):
    assert list(ctx.poly_interpolate_domain([1, 2, 3], [1, 4, 9])) == [0, 0, 1]          # interpolate.rs:57-77
    assert list(ctx.poly_interpolate_domain([1, 3], [5, 9])) == [3, 2]                   # interpolate.rs:80-90
    assert list(ctx.poly_interpolate_domain([1, 2, 3], [2, 5, 10])) == [1, 0, 1]         # interpolate.rs:93-113
    assert list(ctx.poly_interpolate_domain([0, 1, P - 5], [P - 2, 6, 48])) == [P - 2, 5, 3]  # interpolate.rs:139-163
    with pytest.raises(S.StarkPanic, match="no inverse"):                                # mod.rs:613-625
        ctx.poly_interpolate_domain([1, 1], [2, 3])
    assert len(ctx.poly_interpolate_domain([1, 2, 3], [0, 0, 0])) == 0
    assert list(ctx.poly_interpolate_domain([4], [0])) == [0]
    for n in (1, 2, 5, 64, 100, 300, 700):
        d = np.unique(rf(n, n + 50))[:n]
        np.random.default_rng(n).shuffle(d)
        v = rf(n + 9, len(d))
        got = ctx.poly_interpolate_domain(d, v)
        if n <= 100:
            assert np.array_equal(got, oracle.poly_interpolate_domain(d, v)), n          # O(n^3) reference algorithm
        assert np.array_equal(oracle.poly_eval_domain(got, d), v), n


def test_scale_zerofier(ctx, oracle):
    assert list(ctx.poly_scale([2, 3], 5)) == [2, 15] and list(ctx.poly_scale([1, 2, 3], 2)) == [1, 4, 12]  # mod.rs:427-459
    c = rf(3, 5000)
    assert np.array_equal(ctx.poly_scale(c, 12345), oracle.poly_scale(c, 12345))         # mod.rs:99-113
    assert np.array_equal(ctx.poly_scale(c, 0), oracle.poly_scale(c, 0))
    assert list(ctx.poly_zerofier_domain([5])) == [P - 5, 1]                             # mod.rs:320-333
    assert list(ctx.poly_zerofier_domain([2, 3])) == [6, P - 5, 1]
    assert list(ctx.poly_zerofier_domain([1, 2, 3])) == [P - 6, 11, P - 6, 1]
    d = rf(4, 300)
    assert np.array_equal(ctx.poly_zerofier_domain(d), oracle.poly_zerofier(d))          # mod.rs:77-96
    for log_n, off in [(0, 3), (3, 3), (6, 7)]:
        w = oracle.ff_prim_nth_root(1 << log_n)
        dom = [oracle.ff_mul(off, oracle.ff_exp(w, i)) for i in range(1 << log_n)]
        assert np.array_equal(ctx.poly_zerofier_coset(off, log_n), oracle.poly_zerofier(dom))


@pytest.mark.parametrize("log_n,log_b", [(0, 2), (1, 2), (3, 2), (5, 1), (6, 2), (4, 3)])
def test_lde_vs_reference_algorithm(ctx, oracle, log_n, log_b):
    """eval_domain(interpolate_domain(..)) with the reference's own algorithms"""
    cols = rf(log_n + 10 * log_b, 3 << log_n).reshape(3, -1)
    got = ctx.lde(cols, log_b, 3)
    for c in range(3):
        assert np.array_equal(got[c], oracle.lde(cols[c], 1 << log_b, 3))


@pytest.mark.parametrize("log_n,log_b,n_cols", [(10, 2, 5), (13, 2, 3), (16, 2, 2), (12, 4, 2), (20, 2, 1), (21, 2, 1), (22, 1, 1)])
def test_lde_vs_fast_cpu(ctx, oracle, log_n, log_b, n_cols):
    cols = rf(log_n, n_cols << log_n).reshape(n_cols, -1)
    got = ctx.lde(cols, log_b, 3)
    for c in {0, n_cols - 1}:
        assert np.array_equal(got[c], oracle.fast_lde(cols[c], log_n, log_b, 3))


def test_lde_fibonacci_trace(ctx, oracle):
    col = oracle.trace_fibonacci(64) % np.uint64(P)    # trace.rs:36-49 (values < 2^46; the shim reduces for the field)
    assert np.array_equal(ctx.lde(col, 2, 3)[0], oracle.lde(col, 4, 3))


def test_lde_dev_and_ntt_dev(ctx, oracle):
    log_n, log_b, n_cols = 14, 2, 4
    cols = rf(77, n_cols << log_n)
    buf = ctx.upload(cols)
    out = ctx.lde_dev(buf, n_cols, log_n, log_b, 3).download().reshape(n_cols, -1)
    for c in range(n_cols):
        assert np.array_equal(out[c], oracle.fast_lde(cols.reshape(n_cols, -1)[c], log_n, log_b, 3))
    tmp = ctx.alloc(len(cols))
    ctx.ntt_dev(buf, tmp, log_n, batch=n_cols)
    assert np.array_equal(tmp.download()[: 1 << log_n], oracle.fast_eval_coset(cols[: 1 << log_n], 1, log_n))
    ctx.ntt_dev(tmp, tmp, log_n, batch=n_cols, inverse=True)       # in place
    assert np.array_equal(tmp.download(), cols)


def test_poly_div_vs_reference_algorithm(ctx, oracle, S):
    """Polynomial::div / intdiv / modulo (div.rs:6-53): quotient, remainder AND their vector lengths"""
    def chk(a, b):
        q, r = ctx.poly_div(a, b)
        qo, ro = oracle.poly_div(a, b)
        assert np.array_equal(q, qo) and np.array_equal(r, ro), (len(a), len(b), len(q), len(qo), len(r), len(ro))
    # the reference's own cases (div.rs:83-123): (2+3x+x^2)/(1+x) = 2+x r 0 ; (1+x^2)/(1+x) r 2
    q, r = ctx.poly_div([2, 3, 1], [1, 1])
    assert list(q) == [2, 1] and not r.any()
    q, r = ctx.poly_div([1, 0, 1], [1, 1])
    assert list(q) == [P - 1, 1] and list(r[:1]) == [2] and not r[1:].any()
    rng = np.random.default_rng(5)
    for na, nb in [(1, 1), (2, 1), (5, 3), (8, 8), (9, 8), (64, 1), (100, 37), (257, 129), (1000, 999), (4096, 17),
                   (5000, 2500), (1 << 13, 1 << 12)]:
        a = rng.integers(0, P, na, dtype=np.uint64)
        b = rng.integers(0, P, nb, dtype=np.uint64)
        b[-1] = max(int(b[-1]), 1)
        chk(a, b)
    chk([5, 4, 3], [1, 2, 3, 4])                         # deg a < deg b: ([], a)
    chk([1, 2, 3, 0, 0], [4, 5, 0, 0, 0])                # trailing zeros on both sides (length rules)
    chk([0, 0, 0], [7])                                  # zero numerator
    chk([3, 1, 4, 1, 5, 9, 2, 6], [2])                   # constant divisor
    a = rng.integers(0, P, 300, dtype=np.uint64)
    b = rng.integers(1, P, 120, dtype=np.uint64)
    prod = oracle.poly_mul(a, b)
    q, r = ctx.poly_div(prod, b)                         # intdiv: exact division, zero remainder
    assert np.array_equal(q, a) and not r.any()
    with pytest.raises(S.StarkPanic, match="No division by zero"):
        ctx.poly_div([1, 2, 3], [0, 0])
    with pytest.raises(S.StarkPanic, match="No division by zero"):
        ctx.poly_div([1, 2, 3], [])


def test_poly_div_large(ctx, oracle):
    """2^18 / 2^17: checked through a = q b + r with the NTT multiply (the reference's long division is O(n m))"""
    rng = np.random.default_rng(6)
    a = rng.integers(0, P, 1 << 18, dtype=np.uint64)
    b = rng.integers(1, P, 1 << 17, dtype=np.uint64)
    q, r = ctx.poly_div(a, b)
    assert len(q) == (1 << 18) - (1 << 17) + 1 and len(r) == 1 << 18 and not r[(1 << 17) - 1:].any()
    back = oracle.fast_poly_mul(q, b)
    back = (back[: 1 << 18] + r) % P
    assert np.array_equal(back, a)


def test_batched_transforms_larger_than_l2(ctx, oracle):
    """batches of >= 48 MB take the grouped path of ntt_transform (column groups on side streams, ntt.cu): every column
    must still be the transform of that column -- three-pass and two-pass sizes, out of place and in place, and the LDE"""
    for log_n, batch in ((20, 16), (14, 1024)):
        n = 1 << log_n
        cols = rf(log_n + batch, n * batch)
        src, dst = ctx.upload(cols), ctx.alloc(n * batch)
        ctx.ntt_dev(src, dst, log_n, batch=batch)
        got = dst.download().reshape(batch, n)
        for c in (0, 1, batch // 2, batch - 1):
            assert np.array_equal(got[c], oracle.fast_eval_coset(cols.reshape(batch, n)[c], 1, log_n)), (log_n, c)
        ctx.ntt_dev(dst, dst, log_n, batch=batch, inverse=True)        # in place, back to the input
        assert np.array_equal(dst.download(), cols)
    log_n, n_cols = 20, 16
    cols = rf(99, n_cols << log_n)
    out = ctx.lde_dev(ctx.upload(cols), n_cols, log_n, 2, 3).download().reshape(n_cols, -1)
    for c in (0, 7, n_cols - 1):
        assert np.array_equal(out[c], oracle.fast_lde(cols.reshape(n_cols, -1)[c], log_n, 2, 3))


def test_two_pass_big_tile_plans(oracle, S, monkeypatch):
    """STARK_NTT_BIG=1 (read when a context is created): 2^20 .. 2^22 run as two passes on 16384-element tiles -- an
    experiment that is off by default (slower) but must stay bit-exact"""
    monkeypatch.setenv("STARK_NTT_BIG", "1")
    c = S.Context(0)
    try:
        for log_n in (20, 21, 22):
            col = rf(log_n, 1 << log_n)
            ev = c.poly_eval_coset(col, 3, log_n)
            assert np.array_equal(ev, oracle.fast_eval_coset(col, 3, log_n))
            assert np.array_equal(c.poly_interpolate_coset(ev, 3, log_n), col)
        col = rf(5, 1 << 20)
        assert np.array_equal(c.lde(col, 2, 3)[0], oracle.fast_lde(col, 20, 2, 3))
    finally:
        c.close()



def test_batched_small_transforms_2_12(ctx, oracle, S, monkeypatch):
    """batches of >= 8 transforms of length 2^12 take the fused two-pass kernel (k_ntt_small12, ntt.cu): forward, inverse
    (constant post-scale), in place, and -- through the LDE 2^10 -> 2^12 -- the zero-padded input with the geometric
    post-scale before it; every transform checked, and the one-CTA-per-transform kernel must give the same bytes"""
    log_n, batch = 12, 24
    n = 1 << log_n
    cols = rf(4242, n * batch)
    src, dst = ctx.upload(cols), ctx.alloc(n * batch)
    ctx.ntt_dev(src, dst, log_n, batch=batch)
    got = dst.download().reshape(batch, n)
    for c in range(batch):
        assert np.array_equal(got[c], oracle.fast_eval_coset(cols.reshape(batch, n)[c], 1, log_n)), c
    ctx.ntt_dev(dst, dst, log_n, batch=batch, inverse=True)            # in place, back to the input
    assert np.array_equal(dst.download(), cols)
    small = rf(77, 16 << 10).reshape(16, -1)
    lde = ctx.lde(small, 2, 3)                                          # iNTT 16 x 2^10, then NTT 16 x 2^12 with padding
    for c in range(16):
        assert np.array_equal(lde[c], oracle.fast_lde(small[c], 10, 2, 3)), c
    monkeypatch.setenv("STARK_NTT_SMALL_OFF", "1")
    c2 = S.Context(0)
    try:
        d2 = c2.alloc(n * batch)
        c2.ntt_dev(c2.upload(cols), d2, log_n, batch=batch)
        assert np.array_equal(d2.download().reshape(batch, n), got)
    finally:
        c2.close()
