"""Byte-level parity at the BASELINE sizes: the device results at the full configs 3, 4 and 5 against
the oracle run on the box's host cores on the same SURVEY 8(d) inputs (splitmix64, seed 0x5354524B + column index), and
against the committed digests of tests/golden/baseline_digests.json (made by tests/golden/make_baseline_digests.py).

The oracle's hashing / Merkle / fold loops are split over the host threads (oracle_set_threads: loop bodies unchanged,
results identical for any thread count), which keeps the three checks to about two minutes of wall clock."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 0x5354524B
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "baseline_digests.json")))


@pytest.fixture(scope="module")
def par_oracle(oracle):
    old = oracle.set_threads(os.cpu_count() or 1)
    yield oracle
    oracle.set_threads(old)


def test_cfg3_full_proof_bytes(ctx, par_oracle):
    """config 3, W = 1: 2^20-row column -> LDE x4 -> Merkle -> Fri::prove(ef 4, 32 queries): all 681 720 proof bytes and the
    column root equal the oracle's (fri.rs:250-311, merkle.rs:11-38, stream.rs:35-64)."""
    O = par_oracle
    col = O.splitmix64(SEED, 1 << 20)
    roots, proof = ctx.prove_trace(col, 2, 3, 32)
    assert len(proof) == 681720 == GOLD["cfg3"]["proof_len"]
    assert hashlib.sha256(proof).hexdigest() == GOLD["cfg3"]["proof_sha256"]
    assert roots[0].tobytes().hex() == GOLD["cfg3"]["column_root"]
    lde = O.fast_lde(col, 20, 2, 3)
    ref = O.fri_prove(lde, O.ff_prim_nth_root(1 << 22), 3, 4, 32)
    assert ref["top_indices"] == GOLD["cfg3"]["top_indices"]
    assert proof == ref["proof"], "GPU proof bytes differ from the oracle at N = 2^22"
    assert roots[0].tobytes() == ref["proof"][1:33]


def test_cfg3_16_columns_roots(ctx, par_oracle):
    """config 3, W = 16: every column root against the oracle's MerkleTree over its LDE, same proof as W = 1."""
    O = par_oracle
    cols = np.stack([O.splitmix64(SEED + c, 1 << 20) for c in range(16)])
    roots, proof = ctx.prove_trace(cols, 2, 3, 32)
    assert hashlib.sha256(proof).hexdigest() == GOLD["cfg3"]["proof_sha256"]
    for c in (0, 5, 15):
        lde = O.fast_lde(cols[c], 20, 2, 3)
        assert roots[c].tobytes() == O.merkle_commit(O.hash_leaves(lde)), c
    assert len({r.tobytes() for r in roots}) == 16


def test_cfg4_group_full_size(ctx, par_oracle):
    """config 4: ONE full-size group (8 columns x 2^22 rows, blowup 2 -> 2^23 leaves of 8 values) against the oracle; the
    other seven group roots against the golden digests; the 8-root commitment against the oracle's MerkleTree::new."""
    O = par_oracle
    log_n, lb, gw = 22, 1, 8
    n, N = 1 << log_n, 1 << (log_n + lb)
    roots = []
    for k in range(8):
        cols = np.concatenate([O.splitmix64(SEED + k * gw + c, n) for c in range(gw)])
        b = ctx.upload(cols)
        lde = ctx.lde_dev(b, gw, log_n, lb, 3)
        t = ctx.merkle_build_from_buf(lde, N, gw)
        roots.append(t.get_root())
        if k == 0:
            dev = lde.download().reshape(gw, N)
            rows = np.empty((N, gw), dtype=np.uint64)
            for c in range(gw):
                ref = O.fast_lde(cols[c * n:(c + 1) * n], log_n, lb, 3)
                assert np.array_equal(dev[c], ref), c
                rows[:, c] = ref
            assert roots[0] == O.merkle_commit(O.hash_leaves(rows.reshape(-1), gw))
        t.free(), lde.free(), b.free()
    assert [r.hex() for r in roots] == GOLD["cfg4"]["group_roots"]
    leaves = np.frombuffer(b"".join(roots), dtype=np.uint8)
    top = ctx.merkle_commit(leaves)
    assert top == O.merkle_commit(leaves)
    assert top.hex() == GOLD["cfg4"]["commitment"]


def test_cfg5_round_2_24(ctx, par_oracle):
    """config 5 at 2^24 (no 2^24-th root in this field: omega = w_2^23, offset 3 -- the fold and the hashes are still pure
    functions of their inputs): leaf hashes + tree root, alpha, and the whole folded codeword against the oracle's
    fold_codeword (fri.rs:57-91: per-element exp + two xgcd divisions)."""
    O = par_oracle
    k = 24
    cw = O.splitmix64(SEED + k, 1 << k)
    w = O.ff_prim_nth_root(1 << 23)
    b = ctx.upload(cw)
    t = ctx.merkle_build_from_buf(b, 1 << k, 1)
    root = t.get_root()
    assert root.hex() == GOLD["cfg5"]["root"]
    assert root == O.merkle_commit(O.hash_leaves(cw))
    alpha = O.fs_challenge(root)
    assert alpha == GOLD["cfg5"]["alpha"]
    out = ctx.fri_fold_dev(b, 1 << k, alpha, 3, w).download()
    assert hashlib.sha256(out.astype("<u8").tobytes()).hexdigest() == GOLD["cfg5"]["folded_sha256"]
    assert np.array_equal(out, O.fri_fold(cw, alpha, 3, w))
    t.free(), b.free()
