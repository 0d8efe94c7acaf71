"""BASELINE configs 4 and 5 on 1/2/4/8 GPUs (torchrun, one process per GPU) THROUGH THE C ABI's multi-GPU entry points
(stark_mgpu_fold_commit_round, stark_mgpu_lde_commit_dev, stark_mgpu_fri_prove_dev; include/stark_b200.h).
torch.distributed only carries the 128-byte NCCL id, the barrier and the max-over-ranks reduction.

  cfg5  "FRI fold + Merkle commit sweep, codeword 2^16..2^26": one Fri::commit round = leaf hashes + tree (sharded by leaf
        range, subtree roots exchanged inside the climb kernel) + alpha + fold (sharded by output range, every slice stored
        into all replicas).  k <= 23 uses the genuine domain (omega = prim_nth_root(2^k)); k = 24..26 have no 2^k-th root
        in this field (ff.rs:218) and run with omega = prim_nth_root(2^23): throughput-only, degenerate domain.
  cfg4  "2^22-row multi-column trace, LDE and Merkle subtrees sharded": 8 fixed groups x 8 columns, blowup 2 (blowup 4
        would need a 2^24 domain), rank g owns groups {g, g+G, ...}; group roots gathered with ncclAllGather.
Inputs are the SURVEY 8(d) splitmix64 columns (seed 0x5354524B + index), so the oracle reproduces them: --check compares
every rank's results with tests/golden/baseline_digests.json (cfg4 at 2^22, cfg5 at 2^24) and with the oracle for the
smaller sizes.

usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \\
           benchmarks/sharded_sweep.py [--sizes 16,18,...] [--cfg4-log-n 22] [--prove-log-n 22] [--check]
Rank 0 prints one JSON line per measurement; times are CUDA events on the device, max over ranks."""
import argparse
import hashlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stark_rs_b200 as S  # noqa: E402
from stark_rs_b200 import synthetic as G_  # noqa: E402

SEED = 0x5354524B
ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="16,18,20,22,24,26")
ap.add_argument("--cfg4-log-n", type=int, default=0, help="rows (log2) of the config-4 trace; 0 = skip")
ap.add_argument("--prove-log-n", type=int, default=0, help="also run a full sharded Fri::prove of this size; 0 = skip")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--check", action="store_true")
a = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream()
ctx = S.Context(local, stream=stream.cuda_stream)
sizes = [int(x) for x in a.sizes.split(",") if x]
max_cw = 1 << max(sizes + [a.prove_log_n, a.cfg4_log_n + 1, 12])
ids = [S.mgpu_unique_id() if rank == 0 else None]
if world > 1:
    dist.broadcast_object_list(ids, src=0)
    group = S.Group.init(ctx, ids[0], rank, world, max_cw)
else:
    group = S.Group.local([ctx], max_cw)
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "baseline_digests.json")))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps):
    ms, out = [], None
    for i in range(reps + 2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            flush.zero_()
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


def all_ok(flag):
    t = torch.tensor([1.0 if flag else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item() > 0.5)


def emit(line):
    if rank == 0:
        print(json.dumps(line), flush=True)


for k in sizes:
    n = 1 << k
    omega = S.prim_nth_root(1 << min(k, 23))
    cw = G_.splitmix64(SEED + k, n)
    buf = ctx.upload(cw)
    sent0 = group.bytes_sent[0]
    calls = [0]

    def one():
        calls[0] += 1
        (root, alpha, folded), = group.fold_commit_round([buf], n, 3, omega)
        return root, alpha, folded

    ms, (root, alpha, folded) = timed(one, a.reps)
    line = {"config": "cfg5 fold+commit round", "log_n": k, "n_gpus": world, "ms": ms,
            "leaf_plus_node_hashes_per_s": (2 * n - 1) / (ms * 1e-3), "elements_per_s": n / (ms * 1e-3),
            "domain": "genuine" if k <= 23 else "degenerate (omega = w_2^23), throughput-only",
            "bytes_sent_per_round_rank0": (group.bytes_sent[0] - sent0) / calls[0], "root": root.hex(), "alpha": alpha,
            "exchange": "peer-memory stores inside the climb / fold kernels (C ABI: stark_mgpu_fold_commit_round)"}
    if a.check:
        out = folded.download()
        sha = hashlib.sha256(out.astype("<u8").tobytes()).hexdigest()
        if k == GOLD["cfg5"]["log_n"]:
            ok = root.hex() == GOLD["cfg5"]["root"] and alpha == GOLD["cfg5"]["alpha"] and sha == GOLD["cfg5"]["folded_sha256"]
            line["identical_to_oracle_digest"] = all_ok(ok)
        elif k <= 20:
            import oracle as O
            O.set_threads(max(1, (os.cpu_count() or 8) // world))
            r0 = O.merkle_commit(O.hash_leaves(cw))
            a0 = O.fs_challenge(r0)
            ok = root == r0 and alpha == a0 and np.array_equal(out, O.fast_fri_fold(cw, a0, 3, omega))
            line["identical_to_oracle"] = all_ok(ok)
        else:
            # every rank must hold the same replica
            d = [hashlib.sha256(root + sha.encode()).digest()]
            if world > 1:
                dist.broadcast_object_list(d, src=0)
            line["identical_on_every_rank"] = all_ok(d[0] == hashlib.sha256(root + sha.encode()).digest())
    emit(line)
    folded.free(), buf.free()

if a.prove_log_n:
    k = a.prove_log_n
    import oracle as O
    col = G_.splitmix64(SEED, 1 << (k - 2))
    lde = O.fast_lde(col, k - 2, 2, 3)
    buf = ctx.upload(lde)
    w = S.prim_nth_root(1 << k)
    ms, res = timed(lambda: group.fri_prove_dev([buf], 1 << k, 3, w, 4, 32), a.reps)
    proof, top = res[0]
    line = {"config": "sharded Fri::prove (C ABI: stark_mgpu_fri_prove_dev)", "log_n": k, "n_gpus": world, "ms": ms,
            "proof_bytes": len(proof), "proof_sha256": hashlib.sha256(proof).hexdigest()}
    if a.check and k == 22:
        line["identical_to_oracle_digest"] = all_ok(line["proof_sha256"] == GOLD["cfg3"]["proof_sha256"])
    emit(line)
    buf.free()

if a.cfg4_log_n:
    log_n, lb, ng, gw = a.cfg4_log_n, 1, 8, 8
    n = 1 << log_n
    mine = list(range(rank, ng, world))
    bufs = [ctx.upload(np.concatenate([G_.splitmix64(SEED + kg * gw + c, n) for c in range(gw)])) for kg in mine]
    ms, res = timed(lambda: group.lde_commit_dev(bufs, ng, gw, log_n, lb, 3), 3)
    roots, com = res[0]
    line = {"config": "cfg4 LDE + Merkle, 64 columns in 8 groups of 8, blowup 2 (C ABI: stark_mgpu_lde_commit_dev)", "log_n": log_n,
            "n_gpus": world, "ms": ms, "lde_out_elements_per_s": ng * gw * (n << lb) / (ms * 1e-3), "commitment": com.hex()}
    if a.check and log_n == 22:
        ok = com.hex() == GOLD["cfg4"]["commitment"] and [r.tobytes().hex() for r in roots] == GOLD["cfg4"]["group_roots"]
        line["identical_to_oracle_digest"] = all_ok(ok)
    emit(line)
    for b in bufs:
        b.free()

group.close()
ctx.close()
if world > 1:
    dist.destroy_process_group()
