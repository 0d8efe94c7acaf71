"""BASELINE configs 4 and 5 on 1/2/4/8 GPUs (torchrun, one process per GPU, NCCL):

  cfg5  "FRI fold + Merkle commit sweep, codeword 2^16..2^26": one FRI round = leaf hashes + tree (sharded by leaf range,
        all-gather of subtree roots) + alpha from the transcript + fold (sharded by output range, all-gather of slices).
        k <= 23 uses the genuine domain (omega = prim_nth_root(2^k)); k = 24..26 have no 2^k-th root in this field
        (ff.rs:218) and run with omega = prim_nth_root(2^23): throughput-only, degenerate domain.
  cfg4  "2^22-row multi-column trace, LDE and Merkle subtrees sharded": 8 fixed groups x 8 columns, blowup 2 (blowup 4
        would need a 2^24 domain), rank g owns groups {g, g+G, ...}.

usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
           benchmarks/sharded_sweep.py [--sizes 16,18,...] [--cfg4-log-n 22] [--check]
Rank 0 prints one JSON line per measurement; times are CUDA events on the device, max over ranks.  With --check the
results are compared with a single-GPU run of the same input on rank 0 (roots and folded codewords must be identical)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stark_rs_b200 as S  # noqa: E402
from stark_rs_b200 import distributed as D  # noqa: E402
from stark_rs_b200 import synthetic as G_  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="16,18,20,22,24,26")
ap.add_argument("--cfg4-log-n", type=int, default=0, help="rows (log2) of the config-4 trace; 0 = skip")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--check", action="store_true")
ap.add_argument("--fused", action="store_true",
                help="fold kernel stores straight into every rank's replica (symmetric memory: multicast / P2P) "
                     "instead of fold + all-gather")
ap.add_argument("--prove-log-n", type=int, default=0, help="also run a full sharded Fri::prove of this size; 0 = skip")
a = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream()
ctx = S.Context(local, stream=stream.cuda_stream)
# second context on a second stream: config 4 builds a group's tree there while the next group's LDE runs (distributed.py)
aux_stream = torch.cuda.Stream()
aux_ctx = S.Context(local, stream=aux_stream.cuda_stream)
aux_b = D.CudaBackend(aux_ctx, "cuda:%d" % local)
aux_b.stream = aux_stream
b = D.CudaBackend(ctx, "cuda:%d" % local, aux=None if os.environ.get("STARK_NO_OVERLAP") else aux_b)
comm = D.Comm()
P = S.P


def timed(fn, reps):
    ms = []
    for i in range(reps + 2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            out = fn()
            e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if i >= 2:
            ms.append(float(t.item()))
    return sorted(ms)[len(ms) // 2], out


arena = None
if a.fused and world > 1:
    try:
        arena = D.SymmetricArena(1 << (max(int(x) for x in a.sizes.split(",")) - 1), "cuda:%d" % local)
    except Exception as e:
        if rank == 0:
            print(json.dumps({"fused": "unavailable", "why": repr(e)}), flush=True)


def one_round(cw, n, omega, transcript=b""):
    """one Fri::commit round (fri.rs:116-147) on G ranks: returns (root bytes, folded codeword)"""
    tree = D.build_tree(b, comm, cw, n, shard_min=1 << 14)
    root = tree.root_bytes()
    alpha = S.fiat_shamir_challenge(transcript + root)
    h = n // 2
    if arena is not None and h % (4 * world) == 0:
        per = h // world
        arena.reset()
        arena.barrier()                       # peers are done reading the previous round's replica
        nxt, peers, mc = arena.carve(h)
        b.fold_bcast(cw, n, alpha, 3, omega, rank * per, per, peers, mc)
        arena.barrier()
    elif world > 1 and h % world == 0:
        nxt = b.new_codeword(h)
        per = h // world
        b.fold_range(cw, n, alpha, 3, omega, rank * per, per, nxt)
        comm.all_gather_inplace(nxt, rank * per, per)
    else:
        nxt = b.new_codeword(h)
        b.fold_range(cw, n, alpha, 3, omega, 0, h, nxt)
    tree.free()
    return root, nxt


with torch.cuda.stream(stream):
    for k in [int(x) for x in a.sizes.split(",")]:
        n = 1 << k
        omega = S.prim_nth_root(1 << min(k, 23))
        gen = torch.Generator(device="cuda").manual_seed(1234 + k)      # same replica on every rank
        cw = torch.randint(0, P, (n,), dtype=torch.int32, device="cuda", generator=gen)
        ms, (root, nxt) = timed(lambda: one_round(cw, n, omega), a.reps)
        line = {"config": "cfg5 fold+commit round", "log_n": k, "n_gpus": world, "ms": ms,
                "leaf_plus_node_hashes_per_s": (2 * n - 1) / (ms * 1e-3), "elements_per_s": n / (ms * 1e-3),
                "domain": "genuine" if k <= 23 else "degenerate (omega = w_2^23), throughput-only",
                "bytes_gathered_per_round": 32 * world + 4 * (n // 2) if world > 1 else 0,
                "fold_exchange": ("fused (multicast)" if arena.mc else "fused (P2P stores)") if arena is not None
                else ("nccl all_gather" if world > 1 else "none")}
        if a.check and rank == 0:
            solo = D.ShardedTree(n, b.subtree(cw, 0, n), None, D.Comm.__new__(D.Comm))
            ref_root = solo.sub.root_bytes()
            ref = b.new_codeword(n // 2)
            b.fold_range(cw, n, S.fiat_shamir_challenge(ref_root), 3, omega, 0, n // 2, ref)
            line["identical_to_single_gpu"] = bool(ref_root == root and torch.equal(ref, nxt))
            solo.sub.free()
        if rank == 0:
            print(json.dumps(line), flush=True)
        del cw, nxt

    if a.prove_log_n:
        import oracle as O
        O.build()
        k = a.prove_log_n
        col = O.splitmix64(4242, 1 << (k - 2))
        lde = O.fast_lde(col, k - 2, 2, 3)
        fri = D.ShardedFri(b, S.prim_nth_root(1 << k), 3, 1 << k, 4, 32)
        fused = fri.enable_fused_fold() if a.fused else False
        ms, (proof, top) = timed(lambda: fri.prove(b.upload(lde)), 2)
        line = {"config": "sharded Fri::prove (host-orchestrated)", "log_n": k, "n_gpus": world, "ms": ms,
                "fused_fold": bool(fused), "fused_rounds_per_proof": fri.fused_rounds // 4 if fused else 0,
                "proof_bytes": len(proof)}
        if a.check and rank == 0:
            ref = O.fri_prove(lde, S.prim_nth_root(1 << k), 3, 4, 32)
            line["identical_to_oracle"] = bool(proof == ref["proof"] and top == ref["top_indices"])
        if rank == 0:
            print(json.dumps(line), flush=True)

    if a.cfg4_log_n:
        log_n, lb, ng, gw = a.cfg4_log_n, 1, 8, 8
        n = 1 << log_n
        cols_cache = {}

        def cols(kg):
            if kg not in cols_cache:
                gen = torch.Generator(device="cuda").manual_seed(99 + kg)
                cols_cache[kg] = torch.randint(0, P, (gw * n,), dtype=torch.int32, device="cuda", generator=gen)
            return cols_cache[kg]

        for kg in range(rank, ng, world):
            cols(kg)
        ms, (commitment, roots, ldes) = timed(lambda: D.lde_commit_sharded(b, comm, cols, ng, gw, log_n, lb, 3)[:3], 3)
        if rank == 0:
            print(json.dumps({"config": "cfg4 LDE + Merkle, 64 columns in 8 groups of 8, blowup 2", "log_n": log_n,
                              "n_gpus": world, "ms": ms, "lde_out_elements_per_s": ng * gw * (n << lb) / (ms * 1e-3),
                              "commitment": commitment.hex()}), flush=True)
aux_ctx.close()
ctx.close()
if world > 1:
    dist.destroy_process_group()
