"""torchrun ... benchmarks/pipe_trace_mgpu.py with STARK_TRACE_PIPE=1: milestones of stark_mgpu_prove_trace on every rank"""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stark_rs_b200 as S
from stark_rs_b200 import synthetic as G
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cols, log_n = 16, 20
n, N = 1 << log_n, 1 << (log_n + 2)
stream = torch.cuda.Stream()
ctx = S.Context(local, stream=stream.cuda_stream)
ids = [S.mgpu_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
g = S.Group.init(ctx, ids[0], rank, world, N)
host = torch.empty(cols * n, dtype=torch.int64).pin_memory()
hv = host.numpy().view(np.uint64)
for c in range(cols):
    hv[c * n:(c + 1) * n] = G.splitmix64(0x5354524B + c, n)
proof = torch.empty(S.fri_proof_size(N, 4, 32), dtype=torch.uint8).pin_memory().numpy()
roots = torch.empty(cols * 32, dtype=torch.uint8).pin_memory().numpy().reshape(cols, 32)
for i in range(5):
    dist.barrier()
    torch.cuda.synchronize()
    if i == 4:
        sys.stderr.write("---- rank %d, call %d\n" % (rank, i))
        sys.stderr.flush()
    g.prove_trace_ptr(host.data_ptr(), cols, log_n, 2, 3, 32, [roots], [proof])
g.close()
dist.destroy_process_group()
