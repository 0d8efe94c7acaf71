"""The sharded prover behind the C ABI (include/stark_b200.h "groups of GPUs", csrc/mgpu.cu).

On ONE device the group is `world` virtual ranks whose contexts share a stream (stark_mgpu_create_local): every kernel of
the multi-GPU path runs -- range fold + peer stores + leaf hashes, subtree climb, root exchange through the windows, top
levels, transcript, cross-rank proof assembly -- in lock step, and every rank's outputs must equal the oracle's bytes.
With two or more devices the same checks run on a real peer-to-peer group (one host thread, fused signal + wait) and, in
a torchrun-style multi-process group (NCCL bootstrap + CUDA IPC), through tests/mgpu_worker.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SEED = 0x5354524B
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def par_oracle(oracle):
    old = oracle.set_threads(min(os.cpu_count() or 1, 16))
    yield oracle
    oracle.set_threads(old)


def virtual_group(S, ctx, world, max_codeword, shard_log=12):
    """`world` ranks on the device of `ctx`, all on ITS stream (strict host order: no kernel ever waits for a later one)"""
    ctxs = [ctx] + [S.Context(ctx.device, stream=ctx.stream) for _ in range(world - 1)]
    g = S.Group.local(ctxs, max_codeword)
    g.set_shard_log(shard_log)
    return g, ctxs


def close_group(g, ctxs):
    g.close()
    for c in ctxs[1:]:
        c.close()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("log_n", [13, 16])
def test_virtual_fri_prove_bytes(ctx, S, par_oracle, world, log_n):
    """Fri::prove on `world` virtual ranks: every rank's proof bytes and indices equal the oracle's (fri.rs:250-311)"""
    O = par_oracle
    n = 1 << log_n
    cw = O.fast_lde(O.splitmix64(SEED + log_n, n // 4), log_n - 2, 2, 3)
    w = O.ff_prim_nth_root(n)
    ref = O.fri_prove(cw, w, 3, 4, 16)
    g, ctxs = virtual_group(S, ctx, world, n)
    try:
        bufs = [c.upload(cw) for c in ctxs]
        for rep in range(2):                                  # twice: epochs, slots and arena are reused
            res = g.fri_prove_dev(bufs, n, 3, w, 4, 16)
            for k, (proof, top) in enumerate(res):
                assert proof == ref["proof"], (world, log_n, rep, k)
                assert top == ref["top_indices"]
        if world > 1:
            assert all(b > 0 for b in g.bytes_sent)
        # a transcript prefix (FiatShamir state before the call) and a different query count
        ref2 = O.fri_prove(cw, w, 3, 4, 5)
        assert all(p == ref2["proof"] for p, _ in g.fri_prove_dev(bufs, n, 3, w, 4, 5))
        for b in bufs:
            b.free()
    finally:
        close_group(g, ctxs)


def test_virtual_all_rounds_sharded(ctx, S, par_oracle):
    """shard_log low enough that every round down to 4 leaves per rank is sharded, and the panics of Fri::new / prove"""
    O = par_oracle
    n = 1 << 12
    cw = O.splitmix64(11, n)
    w = O.ff_prim_nth_root(n)
    g, ctxs = virtual_group(S, ctx, 4, n, shard_log=4)
    try:
        bufs = [c.upload(cw) for c in ctxs]
        for ef, nq in [(4, 2), (8, 3), (4, 64)]:
            ref = O.fri_prove(cw, w, 3, ef, nq)
            assert all(p == ref["proof"] for p, _ in g.fri_prove_dev(bufs, n, 3, w, ef, nq)), (ef, nq)
        with pytest.raises(S.StarkPanic, match="Expansion factor must be at least 4"):
            g.fri_prove_dev(bufs, n, 3, w, 2, 2)
        with pytest.raises(S.StarkPanic, match="initial codeword length does not match domain length"):
            g.fri_prove_dev(bufs, n, 3, w, 4, 2, domain_length=2 * n)
        # the group still works after a refused call
        ref = O.fri_prove(cw, w, 3, 4, 2)
        assert all(p == ref["proof"] for p, _ in g.fri_prove_dev(bufs, n, 3, w, 4, 2))
        for b in bufs:
            b.free()
    finally:
        close_group(g, ctxs)


@pytest.mark.parametrize("world,n_cols", [(2, 5), (4, 1), (8, 16)])
def test_virtual_prove_trace(ctx, S, par_oracle, world, n_cols):
    """config 3 sharded: LDE + column roots + proof, every rank gets all of it, equal to the single-GPU path and the oracle"""
    O = par_oracle
    log_n = 12
    cols = np.stack([O.splitmix64(SEED + c, 1 << log_n) for c in range(n_cols)])
    roots1, proof1 = ctx.prove_trace(cols, 2, 3, 16)
    lde0 = O.fast_lde(cols[0], log_n, 2, 3)
    ref = O.fri_prove(lde0, O.ff_prim_nth_root(1 << (log_n + 2)), 3, 4, 16)
    assert proof1 == ref["proof"]
    g, ctxs = virtual_group(S, ctx, world, 1 << (log_n + 2))
    try:
        owned = g.owned_columns(n_cols)
        assert sorted(c for o in owned for c in o) == list(range(1, n_cols))
        for rep in range(2):
            for k, (roots, proof) in enumerate(g.prove_trace(cols, 2, 3, 16)):
                assert proof == ref["proof"], (world, k)
                assert np.array_equal(roots, roots1), (world, k)
        c = n_cols - 1
        assert roots1[c].tobytes() == O.merkle_commit(O.hash_leaves(O.fast_lde(cols[c], log_n, 2, 3)))
    finally:
        close_group(g, ctxs)


def test_virtual_prove_trace_zero_rounds(ctx, S, oracle):
    """num_rounds() == 0 (N <= 4 nq): no FRI tree exists, column 0 is committed like the others (single GPU and group)"""
    O = oracle
    cols = np.stack([O.splitmix64(3 + c, 16) for c in range(3)])
    roots1, proof1 = ctx.prove_trace(cols, 2, 3, 16)
    lde = [O.fast_lde(c, 4, 2, 3) for c in cols]
    assert proof1 == O.fri_prove(lde[0], O.ff_prim_nth_root(64), 3, 4, 16)["proof"]
    for c in range(3):
        assert roots1[c].tobytes() == O.merkle_commit(O.hash_leaves(lde[c]))
    g, ctxs = virtual_group(S, ctx, 2, 1 << 12)
    try:
        for roots, proof in g.prove_trace(cols, 2, 3, 16):
            assert proof == proof1 and np.array_equal(roots, roots1)
    finally:
        close_group(g, ctxs)


@pytest.mark.parametrize("world", [2, 8])
def test_virtual_fold_commit_round(ctx, S, par_oracle, world):
    """config 5 round: root, alpha and the folded replica of every rank against the oracle (fri.rs:116-147)"""
    O = par_oracle
    for log_n, w_log in [(14, 14), (16, 12)]:        # the second: a degenerate domain, omega of smaller order
        n = 1 << log_n
        cw = O.splitmix64(SEED + log_n, n)
        w = O.ff_prim_nth_root(1 << w_log)
        root = O.merkle_commit(O.hash_leaves(cw))
        alpha = O.fs_challenge(root)
        ref = O.fri_fold(cw, alpha, 3, w)
        g, ctxs = virtual_group(S, ctx, world, n)
        try:
            bufs = [c.upload(cw) for c in ctxs]
            for r, a, folded in g.fold_commit_round(bufs, n, 3, w):
                assert r == root and a == alpha
                assert np.array_equal(folded.download(), ref)
                folded.free()
            for b in bufs:
                b.free()
        finally:
            close_group(g, ctxs)


@pytest.mark.parametrize("world", [1, 2, 4])
def test_virtual_lde_commit(ctx, S, par_oracle, world):
    """config 4 on a group: group roots (wide leaves, hash.rs:32-35) and the commitment over them against the oracle"""
    O = par_oracle
    log_n, lb, ng, gw = 10, 1, 4, 8
    n, N = 1 << log_n, 1 << (log_n + lb)
    cols = np.concatenate([O.splitmix64(SEED + c, n) for c in range(ng * gw)])
    want = []
    for k in range(ng):
        rows = np.empty((N, gw), dtype=np.uint64)
        for c in range(gw):
            rows[:, c] = O.fast_lde(cols[(k * gw + c) * n:(k * gw + c + 1) * n], log_n, lb, 3)
        want.append(O.merkle_commit(O.hash_leaves(rows.reshape(-1), gw)))
    commitment = O.merkle_commit(np.frombuffer(b"".join(want), dtype=np.uint8))
    g, ctxs = virtual_group(S, ctx, world, N)
    try:
        for rep in range(2):
            for roots, com in g.lde_commit(cols, ng, gw, log_n, lb, 3):
                assert [r.tobytes() for r in roots] == want
                assert com == commitment
    finally:
        close_group(g, ctxs)


def test_group_argument_errors(ctx, S):
    with pytest.raises(S.StarkPanic, match="world size must be 1, 2, 4 or 8"):
        S.Group.local([ctx, ctx, ctx], 1 << 12)
    g, ctxs = virtual_group(S, ctx, 2, 1 << 12)
    try:
        b = ctx.upload(np.arange(1 << 13, dtype=np.uint64))
        with pytest.raises(S.StarkPanic, match="larger than the group's window"):
            g.fold_commit_round([b, b], 1 << 13 << 1, 3, 5)
        b.free()
    finally:
        close_group(g, ctxs)


# ------------------------------------------------------------------------------------------- real multi-GPU groups

def n_devices():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_group_one_process(S, par_oracle, world):
    """one host thread, `world` devices with peer access: fused signal + wait kernels over NVLink"""
    if n_devices() < world:
        pytest.skip("needs %d GPUs" % world)
    O = par_oracle
    ctxs = [S.Context(d) for d in range(world)]
    g = S.Group.local(ctxs, 1 << 20)
    try:
        n = 1 << 18
        cw = O.fast_lde(O.splitmix64(SEED, n // 4), 16, 2, 3)
        w = O.ff_prim_nth_root(n)
        ref = O.fri_prove(cw, w, 3, 4, 32)
        bufs = [c.make_current().upload(cw) for c in ctxs]      # one thread, several devices
        for rep in range(3):
            for proof, top in g.fri_prove_dev(bufs, n, 3, w, 4, 32):
                assert proof == ref["proof"] and top == ref["top_indices"]
        cols = np.stack([O.splitmix64(SEED + c, 1 << 16) for c in range(5)])
        res = g.prove_trace(cols, 2, 3, 32)
        assert all(p == ref["proof"] for _, p in res)
        assert all(np.array_equal(r, res[0][0]) for r, _ in res)
        for c, b in zip(ctxs, bufs):
            c.make_current()
            b.free()
    finally:
        g.close()
        for c in reversed(ctxs):     # device 0 is current again when the test leaves (the module's `ctx` lives there)
            c.make_current()
            c.close()


@pytest.mark.parametrize("world", [2, 8])
def test_process_group_nccl_ipc(world):
    """one process per GPU (torchrun): NCCL bootstrap + CUDA IPC windows inside the library; tests/mgpu_worker.py checks
    every rank's proof bytes, column roots, config-4 commitment and config-5 round against the oracle"""
    if n_devices() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29500 + world + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("MGPU_WORKER_OK") == world, out.stdout[-3000:]


def test_virtual_prove_trace_rejects_non_canonical_column0(ctx, S, oracle):
    """column 0 reaches the other ranks through rank 0's window (one host copy, NVLink broadcast): a value >= p in it must
    fail on EVERY rank like it does on one GPU; without the broadcast (STARK_NO_BCAST0 path = world 1) the same text"""
    cols = np.stack([oracle.splitmix64(5 + c, 1 << 10) for c in range(3)])
    cols[0, 77] = S.P
    with pytest.raises(S.StarkPanic, match="non-canonical field element"):
        ctx.prove_trace(cols, 2, 3, 16)
    g, ctxs = virtual_group(S, ctx, 2, 1 << 12)
    try:
        with pytest.raises(S.StarkPanic, match="non-canonical field element"):
            g.prove_trace(cols, 2, 3, 16)
        cols[0, 77] = 1
        cols[2, 5] = S.P + 3                      # a column only rank 0 or rank 1 owns
        with pytest.raises(S.StarkPanic, match="non-canonical field element"):
            g.prove_trace(cols, 2, 3, 16)
        cols[2, 5] = 3
        ref = oracle.fri_prove(oracle.fast_lde(cols[0], 10, 2, 3), oracle.ff_prim_nth_root(1 << 12), 3, 4, 16)
        assert all(p == ref["proof"] for _, p in g.prove_trace(cols, 2, 3, 16))     # and the group still works
    finally:
        close_group(g, ctxs)
