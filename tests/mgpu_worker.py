"""One rank of a multi-process GPU group (launched by tests/test_gpu_mgpu.py and usable by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tests/mgpu_worker.py)
torch.distributed only carries the 128-byte unique id from rank 0 to the others; the communicator, the CUDA IPC windows
and every exchange live inside libstark_b200.so (stark_mgpu_init).  Every rank checks ITS OWN outputs against the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
import stark_rs_b200 as S  # noqa: E402

SEED = 0x5354524B
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")                  # plumbing for the id only
ids = [S.mgpu_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
ctx = S.Context(local)
g = S.Group.init(ctx, ids[0], rank, world, 1 << 22)
O.set_threads(max(1, (os.cpu_count() or 8) // world))

# Fri::prove, sharded: 2^18 codeword
n = 1 << 18
cw = O.fast_lde(O.splitmix64(SEED, n // 4), 16, 2, 3)
w = O.ff_prim_nth_root(n)
ref = O.fri_prove(cw, w, 3, 4, 32)
buf = ctx.upload(cw)
for rep in range(3):
    (proof, top), = g.fri_prove_dev([buf], n, 3, w, 4, 32)
    assert proof == ref["proof"] and top == ref["top_indices"], "rank %d rep %d: proof differs" % (rank, rep)
buf.free()

# config 3: 5 columns x 2^16 rows
cols = np.stack([O.splitmix64(SEED + c, 1 << 16) for c in range(5)])
(roots, proof), = g.prove_trace(cols, 2, 3, 32)
assert proof == ref["proof"]
for c in range(5):
    assert roots[c].tobytes() == O.merkle_commit(O.hash_leaves(O.fast_lde(cols[c], 16, 2, 3))), c

# config 5 round at 2^20
n5 = 1 << 20
cw5 = O.splitmix64(SEED + 20, n5)
w5 = O.ff_prim_nth_root(n5)
root5 = O.merkle_commit(O.hash_leaves(cw5))
alpha5 = O.fs_challenge(root5)
b5 = ctx.upload(cw5)
(r5, a5, folded), = g.fold_commit_round([b5], n5, 3, w5)
assert r5 == root5 and a5 == alpha5
assert np.array_equal(folded.download(), O.fast_fri_fold(cw5, alpha5, 3, w5))
folded.free(), b5.free()

# config 4: 8 groups of 8 columns x 2^12 rows, blowup 2 (ncclAllGather of the group roots)
log_n, lb, ng, gw = 12, 1, 8, 8
nn, N = 1 << log_n, 1 << (log_n + lb)
c4 = np.concatenate([O.splitmix64(SEED + c, nn) for c in range(ng * gw)])
want = []
for k in range(ng):
    rows = np.empty((N, gw), dtype=np.uint64)
    for c in range(gw):
        rows[:, c] = O.fast_lde(c4[(k * gw + c) * nn:(k * gw + c + 1) * nn], log_n, lb, 3)
    want.append(O.merkle_commit(O.hash_leaves(rows.reshape(-1), gw)))
(groots, com), = g.lde_commit(c4, ng, gw, log_n, lb, 3)
assert [r.tobytes() for r in groots] == want
assert com == O.merkle_commit(np.frombuffer(b"".join(want), dtype=np.uint8))

sent = g.bytes_sent[0]
assert world == 1 or sent > 0
g.barrier()
g.close()
ctx.close()
dist.barrier()
print("MGPU_WORKER_OK rank %d of %d, %d bytes sent" % (rank, world, sent), flush=True)
dist.destroy_process_group()
