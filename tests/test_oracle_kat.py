"""Oracle vs the known answers the REFERENCE's own tests pin.

Each case cites the reference test it is taken from.  CPU only.
"""
import numpy as np
import pytest

P = 998244353


def test_ff_add_sub_mul_neg(oracle):
    O = oracle
    assert O.ff_add(100, 200) == 300                       # ff.rs:344-352
    assert O.ff_add(P - 1, 5) == 4                         # ff.rs:355-362
    assert O.ff_sub(200, 100) == 100                       # ff.rs:389-398
    assert O.ff_sub(5, 10) == P - 5                        # ff.rs:401-408
    assert O.ff_sub(0, 123) == P - 123                     # ff.rs:411-421
    assert O.ff_mul(123, 456) == (123 * 456) % P           # ff.rs:424-433
    assert O.ff_mul(1000000, 2000000) == 2000000000000 % P  # ff.rs:462-469
    assert O.ff_neg(100) == P - 100                        # ff.rs:484-492
    assert O.ff_neg(0) == 0                                # ff.rs:495-501
    assert O.ff_add(123, O.ff_neg(123)) == 0               # ff.rs:504-511


def test_ff_inv_div(oracle):
    O = oracle
    assert O.ff_mul(O.ff_inv(123), 123) == 1               # ff.rs:525-533
    assert O.ff_inv(1) == 1
    assert O.ff_mul(O.ff_inv(P - 1), P - 1) == 1
    with pytest.raises(O.OraclePanic, match="no inverse"):  # ff.rs:552
        O.ff_inv(0)
    with pytest.raises(O.OraclePanic, match="no division by zero"):  # ff.rs:583
        O.ff_div(5, 0)
    assert O.ff_mul(O.ff_div(100, 7), 7) == 100


def test_ff_exp_roots(oracle):
    O = oracle
    assert O.ff_exp(3, 2) == 9 and O.ff_exp(12345, 0) == 1 and O.ff_exp(2, 10) == 1024   # ff.rs:592-625
    assert O.ff_g() == 3                                                                   # ff.rs:628-633
    w8 = O.ff_prim_nth_root(8)                                                             # ff.rs:645-660
    assert O.ff_exp(w8, 8) == 1 and all(O.ff_exp(w8, i) != 1 for i in range(1, 8))
    for k in range(1, 11):                                                                 # ff.rs:663-672
        n = 1 << k
        w = O.ff_prim_nth_root(n)
        assert O.ff_exp(w, n) == 1 and O.ff_exp(w, n // 2) == P - 1
    with pytest.raises(O.OraclePanic):
        O.ff_g(p=17)                                                                       # ff.rs:636-642
    with pytest.raises(O.OraclePanic):
        O.ff_prim_nth_root(8, p=17)
    with pytest.raises(O.OraclePanic, match="n must be a power of two"):                  # ff.rs:699
        O.ff_prim_nth_root(6)
    with pytest.raises(O.OraclePanic, match="n > 2\\^23 not supported"):                  # ff.rs:706
        O.ff_prim_nth_root(1 << 24)
    assert O.ff_sample([]) == 0 and O.ff_sample([42]) == 42                                # ff.rs:713-724


def test_ff_distributive(oracle):
    O = oracle
    a, b, c = 123, 456, 789                                                                # ff.rs:766-790
    assert O.ff_mul(a, O.ff_add(b, c)) == O.ff_add(O.ff_mul(a, b), O.ff_mul(a, c))


def test_poly_zerofier_scale(oracle):
    O = oracle
    assert list(O.poly_zerofier([5])) == [P - 5, 1]                                        # mod.rs:320-333
    assert list(O.poly_zerofier([2, 3])) == [6, P - 5, 1]
    assert list(O.poly_zerofier([1, 2, 3])) == [P - 6, 11, P - 6, 1]
    assert O.poly_eval(O.poly_zerofier([1, 2]), 5) == 12                                   # mod.rs:385-402
    assert list(O.poly_scale([2, 3], 5)) == [2, 15]                                        # mod.rs:427-440
    assert list(O.poly_scale([1, 2, 3], 2)) == [1, 4, 12]
    f, c, x = [3, 1, 4, 1, 5], 7, 11
    assert O.poly_eval(O.poly_scale(f, c), x) == O.poly_eval(f, O.ff_mul(c, x))            # mod.rs:470-488


def test_poly_interpolate(oracle):
    O = oracle
    assert list(O.poly_interpolate_domain([1, 2, 3], [1, 4, 9])) == [0, 0, 1]              # interpolate.rs:57-77
    assert list(O.poly_interpolate_domain([1, 3], [5, 9])) == [3, 2]                       # interpolate.rs:80-90
    assert list(O.poly_interpolate_domain([1, 2, 3], [2, 5, 10])) == [1, 0, 1]             # interpolate.rs:93-113
    r = O.poly_interpolate_domain([0, 1, 2, 4], [3, 7, 13, 35])                            # interpolate.rs:116-136
    assert [O.poly_eval(r, x) for x in (0, 1, 2, 4)] == [3, 7, 13, 35]
    assert list(O.poly_interpolate_domain([0, 1, P - 5], [P - 2, 6, 48])) == [P - 2, 5, 3]  # interpolate.rs:139-163
    with pytest.raises(O.OraclePanic, match="no inverse"):                                 # mod.rs:613-625
        O.poly_interpolate_domain([1, 1], [2, 3])
    # shape rules of SURVEY 3.5 (add.rs:7-12 + mul.rs:7-12)
    assert len(O.poly_interpolate_domain([1, 2, 3], [0, 0, 0])) == 0
    assert list(O.poly_interpolate_domain([4], [0])) == [0]
    assert len(O.poly_interpolate_domain([1, 2, 3, 4], [0, 7, 0, 0])) == 4


def test_poly_mul_eval_div_exp(oracle):
    O = oracle
    assert list(O.poly_mul([1, 1], [1, 1])) == [1, 2, 1]                                   # mul.rs:89-101
    assert list(O.poly_mul([1, 0, 2], [3, 0, 4])) == [3, 0, 10, 0, 8]                      # mul.rs:104-119
    assert list(O.poly_mul([P - 1], [2])) == [P - 2]                                       # mul.rs:182-195
    assert len(O.poly_mul([], [1, 2])) == 0 and len(O.poly_mul([0, 0], [1, 2])) == 0       # mul.rs:7-12
    assert len(O.poly_mul([1, 0], [1, 0, 0])) == 4                                         # length by vector length, mul.rs:14
    assert O.poly_eval([1, 2, 3, 4], 2) == 49                                              # eval.rs:83-95
    assert list(O.poly_eval_domain([1, 1], [0, 1, 2, 3])) == [1, 2, 3, 4]                  # eval.rs:98-118
    q, r = O.poly_div([2, 3, 1], [1, 1])                                                   # div.rs:83-100
    assert list(q) == [2, 1] and O.poly_deg(r) == -1
    q, r = O.poly_div([1, 0, 1], [1, 1])                                                   # div.rs:103-123
    assert O.poly_eval(r, 0) == 2 and O.poly_deg(r) == 0
    with pytest.raises(O.OraclePanic, match="No division by zero"):                        # div.rs:169
        O.poly_div([1, 2], [])
    assert list(O.poly_exp([1, 1], 3)) == [1, 3, 3, 1]                                     # exp.rs:84-100
    assert list(O.poly_exp([3], 4)) == [81]                                                # exp.rs:103-116
    assert list(O.poly_exp([5, 6], 0)) == [1]
    assert O.poly_deg([]) == -1 and O.poly_deg([0, 0]) == -1 and O.poly_deg([1, 2, 0]) == 1  # mod.rs:54-68
    assert O.poly_test_colinearity([1, 2, 3], [2, 4, 6]) and not O.poly_test_colinearity([1, 2, 3], [1, 4, 9])


def test_poly_add_sub(oracle):
    O = oracle
    assert list(O.poly_add([1, 2], [3, 4, 5])) == [4, 6, 5]
    assert list(O.poly_add([], [3, 4])) == [3, 4] and list(O.poly_add([3, 4], [0])) == [3, 4]   # add.rs:7-12
    assert list(O.poly_sub([1, 2], [3])) == [P - 2, 2]
    assert list(O.poly_sub([], [3, 4])) == [P - 3, P - 4]                                  # sub.rs:9-11


def test_merkle_asserts(oracle):
    O = oracle
    with pytest.raises(O.OraclePanic, match="Cannot create tree from empty leaves"):       # merkle.rs:12
        O.merkle_commit(np.zeros((0, 32), dtype=np.uint8))
    with pytest.raises(O.OraclePanic, match="Number of leaves must be power of 2"):        # merkle.rs:13-16
        O.merkle_commit(np.zeros((3, 32), dtype=np.uint8))
    with pytest.raises(O.OraclePanic, match="Index out of bounds"):                        # merkle.rs:68
        O.merkle_open(np.zeros((4, 32), dtype=np.uint8), 4)


def test_fri_asserts(oracle):
    O = oracle
    with pytest.raises(O.OraclePanic, match="Domain length must be power of 2"):           # fri.rs:37-40
        O.fri_num_rounds(48, 4, 2)
    with pytest.raises(O.OraclePanic, match="Expansion factor must be power of 2"):        # fri.rs:41-44
        O.fri_num_rounds(64, 6, 2)
    with pytest.raises(O.OraclePanic, match="Expansion factor must be at least 4"):        # fri.rs:45
        O.fri_num_rounds(64, 2, 2)
    assert O.fri_num_rounds(32, 4, 2) == 2 and O.fri_num_rounds(1 << 22, 4, 32) == 15      # fri.rs:93-103
    with pytest.raises(O.OraclePanic, match="initial codeword length does not match"):     # fri.rs:256-260
        O.fri_prove([1, 2, 3, 4], O.ff_prim_nth_root(32), 3, 4, 2, domain_length=32)
