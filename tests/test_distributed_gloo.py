"""The N>1 path of the sharded prover (stark-rs_b200/distributed.py) on CPU: world_size-2 (and 4) gloo process groups,
arithmetic supplied by the oracle backend (tests/dist_oracle_backend.py).  Checks that leaf-range subtrees + gathered
roots, output-range folds + gathered slices, the query-phase path all-reduce and the host proof assembly give EXACTLY
the oracle's single-process Fri::prove bytes and Merkle roots, for every world size."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle as O
        from dist_oracle_backend import OracleBackend
        from stark_rs_b200 import distributed as D
        b = OracleBackend()
        if case["kind"] == "fri":
            n, off, ef, nq = case["n"], case["offset"], case["ef"], case["nq"]
            cw = O.fast_lde(O.splitmix64(case["seed"], n // ef), (n // ef).bit_length() - 1, ef.bit_length() - 1, off)
            w = O.ff_prim_nth_root(n)
            fri = D.ShardedFri(b, w, off, n, ef, nq, shard_min=case["shard_min"])
            proof, top = fri.prove(b.upload(cw))
            q.put((rank, proof, top, fri.comm.bytes_gathered))
        else:
            comm = D.Comm()
            log_n, lb, ng, gw = case["log_n"], case["log_blowup"], case["groups"], case["width"]
            cols = lambda k: b.upload(np.concatenate([O.splitmix64(1000 * k + c, 1 << log_n) for c in range(gw)]))
            commitment, roots, ldes = D.lde_commit_sharded(b, comm, cols, ng, gw, log_n, lb, 3)
            q.put((rank, commitment, roots.tobytes(), sorted(ldes)))
    finally:
        dist.destroy_process_group()


def _run(world, case):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(out)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_fri_prove_equals_oracle(world, oracle):
    case = dict(kind="fri", n=1 << 10, offset=3, ef=4, nq=8, seed=77, shard_min=1 << 6)
    cw = oracle.fast_lde(oracle.splitmix64(case["seed"], 1 << 8), 8, 2, 3)     # a genuine low-degree codeword
    ref = oracle.fri_prove(cw, oracle.ff_prim_nth_root(case["n"]), 3, 4, 8)
    for rank, proof, top, gathered in _run(world, case):
        assert proof == ref["proof"], "rank %d: sharded proof differs from the oracle's" % rank
        assert top == ref["top_indices"]
        assert gathered > 0                      # the sharded rounds really exchanged roots and slices
    ok, why = oracle.fri_verify(ref["proof"], oracle.ff_prim_nth_root(case["n"]), 3, case["n"], 4, 8)
    assert ok, why


def test_sharded_fri_all_rounds_replicated_when_small(oracle):
    # shard_min above the codeword: every round runs replicated, no data-path collective, same bytes
    case = dict(kind="fri", n=256, offset=17, ef=8, nq=5, seed=5, shard_min=1 << 20)
    cw = oracle.fast_lde(oracle.splitmix64(5, 32), 5, 3, 17)
    ref = oracle.fri_prove(cw, oracle.ff_prim_nth_root(256), 17, 8, 5)
    for rank, proof, top, gathered in _run(2, case):
        assert proof == ref["proof"] and gathered == 0


def test_sharded_lde_commit_equals_oracle(oracle):
    case = dict(kind="lde", log_n=6, log_blowup=1, groups=4, width=3)
    roots = []
    for k in range(4):
        cols = [oracle.fast_lde(oracle.splitmix64(1000 * k + c, 64), 6, 1, 3) for c in range(3)]
        rows = np.stack(cols, axis=1).reshape(-1)                       # leaf i = the 3 values of row i
        roots.append(oracle.merkle_commit(oracle.hash_leaves(rows, 3)))
    want = oracle.merkle_commit(np.frombuffer(b"".join(roots), dtype=np.uint8).reshape(4, 32))
    res = _run(2, case)
    for rank, commitment, group_roots, owned in res:
        assert commitment == want
        assert group_roots == b"".join(roots)
        assert owned == list(range(rank, 4, 2))                         # rank g owns groups g, g+G, ...


def test_single_process_path_needs_no_process_group(oracle):
    from dist_oracle_backend import OracleBackend
    from stark_rs_b200 import distributed as D
    cw = oracle.splitmix64(3, 128)
    w = oracle.ff_prim_nth_root(128)
    fri = D.ShardedFri(OracleBackend(), w, 13, 128, 4, 4)
    proof, top = fri.prove(OracleBackend().upload(cw))
    ref = oracle.fri_prove(cw, w, 13, 4, 4)
    assert proof == ref["proof"] and top == ref["top_indices"]
