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
