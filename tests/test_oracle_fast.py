"""fast_cpu.c (O(n log n) CPU checker / algorithm-matched baseline) vs the reference restatement."""
import numpy as np
import pytest

P = 998244353


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 7])
def test_eval_interpolate_coset(oracle, log_n):
    O = oracle
    n = 1 << log_n
    rng = np.random.default_rng(log_n)
    w = O.ff_prim_nth_root(n)
    for offset in (1, 3, 17):
        dom = [O.ff_mul(offset, O.ff_exp(w, i)) for i in range(n)]
        coeffs = rng.integers(0, P, n, dtype=np.uint64)
        ev = O.poly_eval_domain(coeffs, dom)
        assert np.array_equal(O.fast_eval_coset(coeffs, offset, log_n), ev)
        assert np.array_equal(O.fast_eval_coset(coeffs[: max(1, n // 2)], offset, log_n),
                              O.poly_eval_domain(coeffs[: max(1, n // 2)], dom))
        assert np.array_equal(O.fast_interpolate_coset(ev, offset, log_n), O.poly_interpolate_domain(dom, ev))


@pytest.mark.parametrize("log_n,log_b", [(1, 2), (3, 2), (5, 1), (6, 2)])
def test_lde(oracle, log_n, log_b):
    O = oracle
    col = np.random.default_rng(1).integers(0, P, 1 << log_n, dtype=np.uint64)
    assert np.array_equal(O.fast_lde(col, log_n, log_b, 3), O.lde(col, 1 << log_b, 3))


def test_poly_mul(oracle):
    O = oracle
    rng = np.random.default_rng(2)
    for na, nb in [(1, 1), (3, 5), (64, 64), (100, 29)]:
        a, b = rng.integers(1, P, na, dtype=np.uint64), rng.integers(1, P, nb, dtype=np.uint64)
        assert np.array_equal(O.fast_poly_mul(a, b), O.poly_mul(a, b))


def test_splitmix(oracle):
    # SURVEY 8(d) generator, first outputs for seed 0x5354524B checked against a scalar loop
    def ref(seed, n):
        out, s, M = [], seed, (1 << 64) - 1
        for _ in range(n):
            s = (s + 0x9E3779B97F4A7C15) & M
            z = s
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
            out.append((z ^ (z >> 31)) % P)
        return out
    assert list(oracle.splitmix64(0x5354524B, 16)) == ref(0x5354524B, 16)


def test_oracle_threads_do_not_change_results(oracle):
    """oracle_set_threads only splits the iteration range of the leaf-hash / Merkle-level / fold loops: same bytes for
    any thread count, and a panic raised inside a worker still surfaces with the reference's text (ff.rs:182)."""
    O = oracle
    n = 1 << 14
    cw = O.splitmix64(7, n)
    w = O.ff_prim_nth_root(n)
    base = (O.fri_prove(cw, w, 3, 4, 16)["proof"], O.merkle_build(O.hash_leaves(cw)).tobytes(), O.fri_fold(cw, 99, 3, w).tobytes(),
            O.hash_leaves(cw, 8).tobytes())
    for t in (2, 3, 8):
        old = O.set_threads(t)
        try:
            got = (O.fri_prove(cw, w, 3, 4, 16)["proof"], O.merkle_build(O.hash_leaves(cw)).tobytes(),
                   O.fri_fold(cw, 99, 3, w).tobytes(), O.hash_leaves(cw, 8).tobytes())
            assert got == base, t
            with pytest.raises(O.OraclePanic, match="no division by zero"):
                O.fri_fold(cw, 5, 0, w)
        finally:
            O.set_threads(old)


def test_baseline_digests_small_crosscheck(oracle):
    """the golden generator's code path (tests/golden/make_baseline_digests.py) on a small instance of config 5, so the
    committed full-size digests cannot silently drift from what the oracle computes"""
    import hashlib
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "mk", os.path.join(os.path.dirname(__file__), "golden", "make_baseline_digests.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    r = mk.cfg5(12)
    cw = oracle.splitmix64(mk.SEED + 12, 1 << 12)
    root = oracle.merkle_commit(oracle.hash_leaves(cw))
    assert r["root"] == root.hex() and r["alpha"] == oracle.fs_challenge(root)
    out = oracle.fast_fri_fold(cw, r["alpha"], 3, oracle.ff_prim_nth_root(1 << 23))
    assert r["folded_sha256"] == hashlib.sha256(out.astype("<u8").tobytes()).hexdigest()
