"""GPU parity: FRI fold / commit / prove (fri.rs) + Fiat-Shamir + proof bytes through the C ABI vs the oracle."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P = 998244353
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "survey_vectors.json")))


def rf(seed, n):
    return np.random.default_rng(seed).integers(0, P, n, dtype=np.uint64)


@pytest.mark.parametrize("n", [2, 4, 8, 16, 64, 512, 2048])
def test_fold_vs_reference_algorithm(ctx, oracle, n):
    """fri.rs:57-91 with per-element exp + 2 xgcd divisions"""
    w = oracle.ff_prim_nth_root(n)
    cw = rf(n, n)
    for alpha in (0, 1, P - 1, P, 15764728482632548394, (1 << 64) - 1):     # unreduced challenges, fiat_shamir.rs:21-24
        for offset in (3, 7):
            assert np.array_equal(ctx.fri_fold(cw, alpha, offset, w), oracle.fri_fold(cw, alpha, offset, w)), (n, alpha)


@pytest.mark.parametrize("log_n", [13, 16, 20, 22, 23])
def test_fold_large_vs_closed_form(ctx, oracle, log_n):
    n = 1 << log_n
    w = oracle.ff_prim_nth_root(n)
    cw = rf(log_n, n)
    alpha = 0xDEADBEEFCAFEBABE
    assert np.array_equal(ctx.fri_fold(cw, alpha, 3, w), oracle.fast_fri_fold(cw, alpha, 3, w))


def test_alignment_contract(ctx, oracle, S):
    """include/stark_b200.h, ALIGNMENT: a wrapped device pointer must be 16-byte aligned (the kernels use 128-bit accesses)
    and is rejected as an argument error, not left to raise a sticky misaligned-address fault; an output VIEW at an odd
    element offset is legal and takes the scalar path of the fold"""
    n = 1 << 12
    w = oracle.ff_prim_nth_root(n)
    cw = rf(77, n)
    src = ctx.upload(cw)
    with pytest.raises(S.StarkPanic, match="16-byte aligned"):
        ctx.wrap(src.ptr + 4, n - 1)
    view = ctx.wrap(src.ptr + 16, n - 4)                       # an aligned view is fine
    assert np.array_equal(view.download(), cw[4:])
    want = oracle.fri_fold(cw, 12345, 3, w)
    for out_off, i0, cnt in ((1, 0, n // 2), (3, 4, 100), (2, 1, 7), (0, 0, n // 2)):
        out = ctx.alloc(n // 2 + 8)
        ctx.fri_fold_range_dev(src, n, 12345, 3, w, i0, cnt, out, out_off)
        assert np.array_equal(out.download(out_off, cnt), want[i0:i0 + cnt]), (out_off, i0, cnt)
        out.free()
    assert np.array_equal(ctx.fri_fold(cw, 12345, 3, w), want)    # the context is still healthy
    src.free()


def test_fold_degenerate_domain(ctx, oracle, S):
    """Fri::new never checks omega's order (fri.rs:30-55): the fold is a pure function of its inputs"""
    cw = rf(1, 256)
    for omega in (1, 5, oracle.ff_prim_nth_root(8)):
        assert np.array_equal(ctx.fri_fold(cw, 99, 3, omega), oracle.fri_fold(cw, 99, 3, omega))
    with pytest.raises(S.StarkPanic, match="no division by zero"):          # ff.rs:182 via fri.rs:75
        ctx.fri_fold(cw, 99, 0, 5)
    big = rf(2, 1 << 24)                                                    # > 2^23: throughput-only domain (SURVEY 8d cfg 5)
    w23 = oracle.ff_prim_nth_root(1 << 23)
    assert np.array_equal(ctx.fri_fold(big, 7, 3, w23), oracle.fast_fri_fold(big, 7, 3, w23))


def _statement(O, n, offset, coeffs):
    w = O.ff_prim_nth_root(n)
    dom = [O.ff_mul(offset, O.ff_exp(w, i)) for i in range(n)]
    return w, O.poly_eval_domain(coeffs, dom)


@pytest.mark.parametrize("case", G["fri_proofs"], ids=lambda c: "n%d" % c["n"])
def test_prove_reference_statements(ctx, oracle, case):
    """the four statements of fri.rs:532-693: identical proof bytes, and Fri::verify accepts them"""
    n, off, ef, nq = case["n"], case["offset"], case["ef"], case["nq"]
    w, cw = _statement(oracle, n, off, case["coeffs"])
    proof, top = ctx.fri_prove(cw, off, w, ef, nq)
    ref = oracle.fri_prove(cw, w, off, ef, nq)
    assert proof == ref["proof"]
    assert top == ref["top_indices"] == case["top"]
    assert hashlib.sha256(proof).hexdigest() == case["sha256"] and len(proof) == case["bytes"]
    ok, why = oracle.fri_verify(proof, w, off, n, ef, nq)
    assert ok, why


@pytest.mark.parametrize("log_n,ef,nq", [(5, 4, 2), (6, 4, 3), (8, 8, 5), (10, 4, 8), (12, 4, 16), (14, 4, 32), (16, 16, 20)])
def test_commit_and_prove_random_low_degree(ctx, oracle, log_n, ef, nq):
    n = 1 << log_n
    w = oracle.ff_prim_nth_root(n)
    coeffs = rf(log_n, n // ef)
    cw = oracle.fast_eval_coset(coeffs, 3, log_n)
    ref = oracle.fri_prove(cw, w, 3, ef, nq)
    st = ctx.fri_commit(cw, 3, w, ef, nq)                                   # fri.rs:105-156
    assert st.rounds == ref["rounds"]
    assert st.alphas() == ref["alphas"]                                     # raw u64 challenges
    roots = st.roots()
    for r in range(st.rounds):
        assert roots[r].tobytes() == ref["proof"][33 * r + 1: 33 * r + 33]
    a0 = ref["alphas"][0]
    assert np.array_equal(st.codeword(1), oracle.fast_fri_fold(cw, a0, 3, w))
    assert np.array_equal(st.open(0, 5), oracle.merkle_open(oracle.hash_leaves(cw), 5))
    proof, top = ctx.fri_prove(cw, 3, w, ef, nq)                            # fri.rs:250-311 + stream.rs:35-64
    assert proof == ref["proof"] and top == ref["top_indices"]
    ok, why = oracle.fri_verify(proof, w, 3, n, ef, nq)
    assert ok, why


def test_prove_non_low_degree_and_transcript_prefix(ctx, oracle):
    n, ef, nq = 256, 4, 8
    w = oracle.ff_prim_nth_root(n)
    cw = rf(9, n)                                                           # random codeword: still must match byte for byte
    proof, _ = ctx.fri_prove(cw, 3, w, ef, nq)
    assert proof == oracle.fri_prove(cw, w, 3, ef, nq)["proof"]
    assert not oracle.fri_verify(proof, w, 3, n, ef, nq)[0]
    # a transcript that already holds data (FiatShamir::absorb before prove): roots must differ from the empty one
    for prefix in (b"x", b"0123456789abcdef0123456789abcdef", bytes(range(45))):
        st = ctx.fri_commit(cw, 3, w, ef, nq, transcript=prefix)
        roots = st.roots()
        tr = prefix + roots[0].tobytes()
        assert st.alphas()[0] == oracle.fs_challenge(tr)                    # fiat_shamir.rs:19-25
        tr += roots[1].tobytes()
        assert st.alphas()[1] == oracle.fs_challenge(tr)


def test_prove_edge_parameters(ctx, oracle, S):
    w = oracle.ff_prim_nth_root(16)
    cw = rf(3, 16)
    for ef, nq in [(4, 1), (4, 2), (4, 3), (8, 2), (16, 2), (4, 4)]:       # includes num_rounds 0 and 1
        ref = oracle.fri_prove(cw, w, 3, ef, nq)
        proof, top = ctx.fri_prove(cw, 3, w, ef, nq)
        assert proof == ref["proof"] and top == ref["top_indices"], (ef, nq)
    with pytest.raises(S.StarkPanic, match="initial codeword length does not match domain length"):   # fri.rs:256-260
        ctx.fri_prove(cw, 3, w, 4, 2, domain_length=32)
    with pytest.raises(S.StarkPanic, match="Expansion factor must be at least 4"):
        ctx.fri_prove(cw, 3, w, 2, 2)
    with pytest.raises(S.StarkPanic, match="cannot sample more indices|not enough entropy"):          # fri.rs:183-192
        ctx.fri_prove(cw, 3, w, 4, 40)


def test_prove_full_size_2_22(ctx, oracle):
    """BASELINE config 3 size: N = 2^22, ef 4, 32 queries, 15 rounds.  The oracle's Fri::verify must accept the
    GPU proof (any wrong fold / index / transcript / path order fails it) and round 0 is cross-checked against
    values recomputed on the CPU."""
    log_n = 22
    n = 1 << log_n
    w = oracle.ff_prim_nth_root(n)
    assert w == 267099868
    coeffs = rf(22, n // 4)
    cw = oracle.fast_eval_coset(coeffs, 3, log_n)
    proof, top = ctx.fri_prove(cw, 3, w, 4, 32)
    assert len(proof) == ctx_size(n) and len(set(t % 256 for t in top)) == 32
    ok, why = oracle.fri_verify(proof, w, 3, n, 4, 32)
    assert ok, why
    alpha0 = oracle.fs_challenge(proof[1:33])
    fold1 = oracle.fast_fri_fold(cw, alpha0, 3, w)
    # first revealed triple of round 0: (cw[a], cw[a + n/2], fold1[a]) with a = top[0] % (n/2)
    base = 33 * 15 + 9 + 8 * 256
    a = top[0] % (n // 2)
    trip = np.frombuffer(proof[base + 9: base + 33], dtype="<u8")
    assert list(trip) == [int(cw[a]), int(cw[a + n // 2]), int(fold1[a])]


def ctx_size(n):
    import stark_rs_b200 as S
    return S.fri_proof_size(n, 4, 32)


@pytest.mark.parametrize("log_n,n_cols", [(6, 1), (8, 3), (12, 2)])
def test_prove_trace_pipeline(ctx, oracle, log_n, n_cols):
    """BASELINE config 3 pipeline at oracle-checkable sizes: LDE (b=4, offset 3) + per-column Merkle + FRI"""
    cols = rf(log_n, n_cols << log_n).reshape(n_cols, -1)
    nq = 4 if log_n < 10 else 32
    roots, proof = ctx.prove_trace(cols, 2, 3, nq)
    N = 4 << log_n
    w = oracle.ff_prim_nth_root(N)
    for c in range(n_cols):
        lde = oracle.fast_lde(cols[c], log_n, 2, 3)
        assert roots[c].tobytes() == oracle.merkle_commit(oracle.hash_leaves(lde))
        if c == 0:
            assert proof == oracle.fri_prove(lde, w, 3, 4, nq)["proof"]
    assert oracle.fri_verify(proof, w, 3, N, 4, nq)[0]


def _i128_rows(seed, n_rows, n_cols):
    """random i128 rows: small values, negatives, values above 2^64 and the extremes (trace.rs:4-7 holds Vec<Vec<i128>>)"""
    rng = np.random.default_rng(seed)
    lo = rng.integers(0, 1 << 63, (n_rows, n_cols), dtype=np.uint64).astype(object) * 2 + rng.integers(0, 2, (n_rows, n_cols)).astype(object)
    hi = rng.integers(-(1 << 62), 1 << 62, (n_rows, n_cols)).astype(object) * 2
    kind = rng.integers(0, 4, (n_rows, n_cols))
    rows = [[int(lo[r, c]) % 1000 if kind[r, c] == 0 else int(lo[r, c]) if kind[r, c] == 1 else
             -int(lo[r, c]) if kind[r, c] == 2 else (int(hi[r, c]) << 64) + int(lo[r, c]) for c in range(n_cols)]
            for r in range(n_rows)]
    rows[0][0], rows[-1][-1] = (1 << 127) - 1, -(1 << 127)
    return rows


@pytest.mark.parametrize("n_rows,n_cols", [(1, 1), (5, 3), (33, 31), (64, 1), (100, 40), (1 << 12, 2)])
def test_trace_to_columns(ctx, oracle, n_rows, n_cols):
    """trace ingestion: row-major i128 -> column-major residues, ragged tile edges included"""
    rows = _i128_rows(n_rows * 131 + n_cols, n_rows, n_cols)
    got = ctx.trace_to_columns(rows).download().reshape(n_cols, n_rows)
    want = oracle.trace_columns(rows) % np.uint64(P)     # raw `as u64` casts (trace.rs:29-34), residue as ff.rs:138-152 sees them
    assert np.array_equal(got, want)


@pytest.mark.parametrize("log_n,n_cols", [(6, 1), (9, 3), (12, 2)])
def test_prove_trace_rows(ctx, oracle, log_n, n_cols):
    """the prove pipeline fed with the reference's own trace container layout gives the bytes of the column-major path"""
    rows = _i128_rows(log_n, 1 << log_n, n_cols)
    nq = 4 if log_n < 10 else 32
    cols = oracle.trace_columns(rows) % np.uint64(P)
    roots, proof = ctx.prove_trace_rows(rows, 2, 3, nq)
    roots2, proof2 = ctx.prove_trace(cols, 2, 3, nq)
    assert proof == proof2 and np.array_equal(roots, roots2)
    lde = oracle.fast_lde(cols[0], log_n, 2, 3)
    assert proof == oracle.fri_prove(lde, oracle.ff_prim_nth_root(4 << log_n), 3, 4, nq)["proof"]


def test_prove_trace_rows_fibonacci(ctx, oracle):
    """BASELINE config 1's Trace::fibonacci(64) column through the row-major entry point (trace.rs:36-49)"""
    fib = oracle.trace_fibonacci(64)
    rows = [[int(v)] for v in fib]
    roots, proof = ctx.prove_trace_rows(rows, 2, 3, 8)
    lde = oracle.lde(fib % np.uint64(P), 4, 3)
    assert proof == oracle.fri_prove(lde, oracle.ff_prim_nth_root(256), 3, 4, 8)["proof"]
    assert roots[0].tobytes() == oracle.merkle_commit(oracle.hash_leaves(lde))


def test_prove_trace_rejects_non_canonical_input(ctx, S):
    """stark_prove_trace checks canonical input without a mid-pipeline host round trip: the error must still surface"""
    col = np.arange(1 << 8, dtype=np.uint64)
    col[77] = P            # >= p: never silently reduced (the reference would hash / serialise the raw value)
    with pytest.raises(S.StarkPanic, match="non-canonical"):
        ctx.prove_trace(col, 2, 3, 8)
    ok_roots, ok_proof = ctx.prove_trace(np.arange(1 << 8, dtype=np.uint64), 2, 3, 8)   # the context is still usable
    assert len(ok_proof) > 0


def test_prove_randomised_parameters(ctx, oracle):
    """40 seeded random statements: size 2^4..2^13, expansion factor, query count, offset, codeword kind; every proof
    byte and every returned index must equal the oracle's.  Exercises the single-CTA tail alone (small n), tail + fused
    fold/leaf rounds, the multi-CTA climb (n >= 2^11) and the index-sampling reject rule with many queries."""
    rng = np.random.default_rng(20261018)
    for case in range(40):
        log_n = int(rng.integers(4, 14))
        n = 1 << log_n
        ef = int(rng.choice([4, 8, 16]))
        if ef >= n:
            ef = 4
        w = oracle.ff_prim_nth_root(n)
        offset = int(rng.integers(1, P))
        rounds_ok = [q for q in (1, 2, 3, 5, 8, 13, 21, 32, 40) if q <= n // ef // 2 or q <= 2]
        nq = int(rng.choice(rounds_ok))
        if rng.random() < 0.5:
            base = rng.integers(0, P, n // ef, dtype=np.uint64)
            cw = oracle.fast_eval_coset(base, offset, log_n)               # genuine low-degree codeword
        else:
            cw = rng.integers(0, P, n, dtype=np.uint64)
        try:
            ref = oracle.fri_prove(cw, w, offset, ef, nq)
        except oracle.OraclePanic:
            continue                                                        # parameters the reference rejects too
        proof, top = ctx.fri_prove(cw, offset, w, ef, nq)
        assert proof == ref["proof"], (case, log_n, ef, nq, offset)
        assert top == ref["top_indices"], (case, log_n, ef, nq)
