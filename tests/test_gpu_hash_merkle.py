"""GPU parity: hash.rs / merkle.rs kernels through the C ABI vs the oracle."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P = 998244353
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "survey_vectors.json")))


def rf(seed, n):
    return np.random.default_rng(seed).integers(0, P, n, dtype=np.uint64)


def test_hash_bytes_all_lengths(ctx, oracle):
    rng = np.random.default_rng(5)
    assert ctx.hash_empty(3)[1].tobytes() == oracle.hash_from_bytes(b"")                # hash.rs:7-30, empty message
    for ln in list(range(1, 70)) + [95, 96, 97, 128, 200]:
        msgs = rng.integers(0, 256, 37 * ln, dtype=np.uint8)
        got = ctx.hash_bytes(msgs, ln)
        for k in (0, 17, 36):
            assert got[k].tobytes() == oracle.hash_from_bytes(msgs[k * ln:(k + 1) * ln].tobytes()), ln


def test_hash_golden(ctx):
    for msg_hex, digest in G["hash_from_bytes"].items():
        m = bytes.fromhex(msg_hex)
        if m:
            assert ctx.hash_bytes(m, len(m))[0].tobytes().hex() == digest
    assert ctx.hash_empty(1)[0].tobytes().hex() == G["hash_from_bytes"][""]
    assert ctx.hash_leaves([0])[0].tobytes().hex() == G["hash_from_u64_0"]
    assert ctx.hash_leaves([1])[0].tobytes().hex() == G["hash_from_field_elements_1"]
    h1, h2 = ctx.hash_bytes(b"hello", 5)[0], ctx.hash_bytes(b"hallo", 5)[0]
    assert int((h1 != h2).sum()) > 10                                                    # hash.rs:121-132


@pytest.mark.parametrize("width", [1, 2, 3, 4, 5, 8, 9, 16])
def test_hash_leaves(ctx, oracle, width):
    vals = np.concatenate([rf(width, 1000 * width), np.array([0, P - 1] * width, dtype=np.uint64)])
    got = ctx.hash_leaves(vals, width)
    assert np.array_equal(got, oracle.hash_leaves(vals, width))                          # hash.rs:32-35


@pytest.mark.parametrize("log_n", list(range(0, 14)))
def test_merkle_all_levels(ctx, oracle, log_n):
    n = 1 << log_n
    leaves = np.random.default_rng(log_n).integers(0, 256, (n, 32), dtype=np.uint8)
    t = ctx.merkle_build(leaves)
    assert t.num_levels == log_n + 1 and t.num_leaves == n                               # merkle.rs:103-108
    assert np.array_equal(t.nodes(), oracle.merkle_build(leaves))                        # merkle.rs:11-38
    assert t.get_root() == oracle.merkle_commit(leaves) == ctx.merkle_commit(leaves)     # merkle.rs:44-65
    for idx in {0, n - 1, n // 3}:
        path = t.open(idx)
        assert np.array_equal(path, oracle.merkle_open(leaves, idx))                     # merkle.rs:67-80
        assert oracle.merkle_verify(leaves[idx], idx, path, t.get_root())                # merkle.rs:82-96, 111-122


def test_merkle_golden_and_errors(ctx, oracle, S):
    l8 = ctx.hash_bytes(np.arange(8, dtype=np.uint8), 1)                                 # merkle.rs:105,114
    assert ctx.merkle_commit(l8[:4]).hex() == G["merkle_root_4"]
    assert ctx.merkle_commit(l8).hex() == G["merkle_root_8"]
    assert ctx.merkle_commit(np.zeros((2, 32), dtype=np.uint8)).hex() == G["combine_zero_zero"]
    with pytest.raises(S.StarkPanic, match="Cannot create tree from empty leaves"):      # merkle.rs:12
        ctx.merkle_build(np.zeros((0, 32), dtype=np.uint8))
    with pytest.raises(S.StarkPanic, match="Number of leaves must be power of 2"):       # merkle.rs:13-16
        ctx.merkle_build(np.zeros((3, 32), dtype=np.uint8))
    t = ctx.merkle_build(l8)
    with pytest.raises(S.StarkPanic, match="Index out of bounds"):                       # merkle.rs:68
        t.open(8)


@pytest.mark.parametrize("log_n,width", [(10, 1), (12, 8), (16, 1), (18, 1)])
def test_merkle_from_values(ctx, oracle, log_n, width):
    n = 1 << log_n
    vals = rf(log_n + width, n * width)
    t = ctx.merkle_build_from_values(vals, width)                                        # fri.rs:118-127
    leaves = oracle.hash_leaves(vals, width)
    assert np.array_equal(t.level(0), leaves)
    assert t.get_root() == oracle.merkle_commit(leaves)
    # device-resident, column-major variant (LDE output layout)
    cm = np.ascontiguousarray(vals.reshape(n, width).T).reshape(-1)
    t2 = ctx.merkle_build_from_buf(ctx.upload(cm), n, width)
    assert t2.get_root() == t.get_root()


def test_merkle_large_properties(ctx, oracle):
    """2^22 leaves (BASELINE config 3 size): size-independent checks -- every opened path verifies against the
    root with the oracle's MerkleTree::verify, subtree roots of the two halves combine to the root."""
    n = 1 << 22
    vals = rf(22, n)
    t = ctx.merkle_build_from_values(vals)
    root = t.get_root()
    for idx in (0, 1, n // 2 - 1, n // 2, n - 1, 1234567):
        leaf = oracle.hash_from_field_elements([int(vals[idx])])
        assert oracle.merkle_verify(leaf, idx, t.open(idx), root)
    lvl = t.level(21)
    assert oracle.hash_combine(lvl[0], lvl[1]) == root
    half = ctx.merkle_build_from_values(vals[: n // 2])
    assert half.get_root() == lvl[0].tobytes()                                           # disjoint subtrees
