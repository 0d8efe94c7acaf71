"""CudaBackend of the sharded prover on one B200: the per-rank pieces (leaf-range subtrees, output-range folds) must
reassemble to exactly the single-GPU / oracle results for G = 2, 4, 8, and ShardedFri at world size 1 must emit the
oracle's proof bytes.  (The collectives themselves are covered on CPU by tests/test_distributed_gloo.py and on
2-8 GPUs by benchmarks/sharded_sweep.py, which asserts the same equalities.)"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def backend(S):
    from stark_rs_b200 import distributed as D
    stream = torch.cuda.Stream()
    ctx = S.Context(0, stream=stream.cuda_stream)
    with torch.cuda.stream(stream):
        yield D.CudaBackend(ctx, "cuda:0")
    ctx.close()


@pytest.mark.parametrize("G", [2, 4, 8])
def test_subtrees_and_top_tree_equal_full_tree(backend, oracle, G):
    n = 1 << 12
    vals = oracle.splitmix64(11, n)
    cw = backend.upload(vals)
    full = oracle.merkle_build(oracle.hash_leaves(vals))
    roots = backend.new_hashes(G)
    per = n // G
    subs = []
    for g in range(G):
        t = backend.subtree(cw, g * per, per)
        roots[g].copy_(t.root)
        subs.append(t)
    top = backend.tree_from_hashes(roots)
    assert top.root_bytes() == full[-1].tobytes()
    # a sharded authentication path = owner's subtree path + top-tree path (merkle.rs:67-80)
    for idx in [0, 1, per - 1, per, n - 1, 1234]:
        g = idx // per
        lower = subs[g].open_batch([idx % per])[0]
        upper = top.open_batch([g])[0]
        want = oracle.merkle_open(full[:n], idx)
        assert np.array_equal(np.concatenate([lower, upper]), want)
    for t in subs + [top]:
        t.free()


@pytest.mark.parametrize("G", [2, 4, 8])
def test_fold_ranges_equal_full_fold(backend, oracle, G):
    n = 1 << 14
    vals = oracle.splitmix64(5, n)
    w = oracle.ff_prim_nth_root(n)
    alpha = 15764728482632548394          # raw, unreduced (fiat_shamir.rs:21-24)
    cw = backend.upload(vals)
    out = backend.new_codeword(n // 2)
    per = (n // 2) // G
    for g in range(G):
        backend.fold_range(cw, n, alpha, 3, w, g * per, per, out)
    assert np.array_equal(backend.download(out), oracle.fast_fri_fold(vals, alpha, 3, w))


def test_sharded_fri_world1_equals_oracle_and_single_gpu(backend, oracle, ctx):
    from stark_rs_b200 import distributed as D
    n, ef, nq = 1 << 12, 4, 16
    cw = oracle.fast_lde(oracle.splitmix64(9, n // ef), 10, 2, 3)
    w = oracle.ff_prim_nth_root(n)
    proof, top = D.ShardedFri(backend, w, 3, n, ef, nq).prove(backend.upload(cw))
    ref = oracle.fri_prove(cw, w, 3, ef, nq)
    assert proof == ref["proof"] and top == ref["top_indices"]
    single, top1 = ctx.fri_prove(cw, 3, w, ef, nq)
    assert single == proof and top1 == top


def test_lde_commit_world1(backend, oracle):
    from stark_rs_b200 import distributed as D
    log_n, lb, ng, gw = 8, 1, 2, 8
    cols = lambda k: backend.upload(np.concatenate([oracle.splitmix64(1000 * k + c, 1 << log_n) for c in range(gw)]))
    commitment, roots, _ = D.lde_commit_sharded(backend, D.Comm(), cols, ng, gw, log_n, lb, 3)
    want = []
    for k in range(ng):
        ldes = [oracle.fast_lde(oracle.splitmix64(1000 * k + c, 1 << log_n), log_n, lb, 3) for c in range(gw)]
        want.append(oracle.merkle_commit(oracle.hash_leaves(np.stack(ldes, axis=1).reshape(-1), gw)))
    assert roots.tobytes() == b"".join(want)
    assert commitment == oracle.merkle_commit(np.frombuffer(b"".join(want), dtype=np.uint8).reshape(ng, 32))


def test_lde_commit_overlapped_trees(backend, oracle, S):
    """the same commitment when every group's tree is built on an auxiliary context / stream while the next LDE runs"""
    from stark_rs_b200 import distributed as D
    log_n, lb, ng, gw = 12, 1, 4, 8
    cols = lambda k: backend.upload(np.concatenate([oracle.splitmix64(77 * k + c, 1 << log_n) for c in range(gw)]))
    plain = D.lde_commit_sharded(backend, D.Comm(), cols, ng, gw, log_n, lb, 3)
    aux_stream = torch.cuda.Stream()
    aux_ctx = S.Context(0, stream=aux_stream.cuda_stream)
    try:
        aux = D.CudaBackend(aux_ctx, "cuda:0")
        aux.stream = aux_stream
        both = D.CudaBackend(backend.ctx, "cuda:0", aux=aux)
        got = D.lde_commit_sharded(both, D.Comm(), cols, ng, gw, log_n, lb, 3)
        torch.cuda.synchronize()
        assert got[0] == plain[0] and got[1].tobytes() == plain[1].tobytes()
    finally:
        aux_ctx.close()


def test_fold_bcast_writes_every_replica(backend, oracle):
    """k_fri_fold_bcast with the P2P store path on one GPU: three 'replicas' (plain device buffers standing in for the
    peers' mapped memory) must all receive the folded range, equal to the oracle's fold."""
    n = 1 << 14
    vals = oracle.splitmix64(21, n)
    w = oracle.ff_prim_nth_root(n)
    alpha = (1 << 64) - 59
    cw = backend.upload(vals)
    reps = [torch.zeros(n // 2, dtype=torch.int32, device="cuda:0") for _ in range(3)]
    per = (n // 2) // 4
    for g in range(4):                                  # four ranks' ranges, each stored into all replicas
        backend.fold_bcast(cw, n, alpha, 3, w, g * per, per, [r.data_ptr() for r in reps], 0)
    want = oracle.fast_fri_fold(vals, alpha, 3, w)
    for r in reps:
        assert np.array_equal(backend.download(r), want)
    with pytest.raises(Exception):
        backend.fold_bcast(cw, n, alpha, 3, w, 2, per, [reps[0].data_ptr()], 0)      # range not a multiple of 4
