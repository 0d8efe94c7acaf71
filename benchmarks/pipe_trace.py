"""STARK_TRACE_PIPE=1 python benchmarks/pipe_trace.py : device timestamps of the config-3 pipeline, host input vs resident"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stark_rs_b200 as S
from stark_rs_b200 import synthetic as G
cols, log_n = int(os.environ.get("COLS", "16")), 20
n, N = 1 << log_n, 1 << (log_n + 2)
stream = torch.cuda.Stream()
ctx = S.Context(0, stream=stream.cuda_stream)
host = torch.empty(cols * n, dtype=torch.int64).pin_memory()
hv = host.numpy().view(np.uint64)
for c in range(cols):
    hv[c * n:(c + 1) * n] = G.splitmix64(0x5354524B + c, n)
proof = torch.empty(S.fri_proof_size(N, 4, 32), dtype=torch.uint8).pin_memory().numpy()
roots = torch.empty(cols * 32, dtype=torch.uint8).pin_memory().numpy().reshape(cols, 32)
dev = ctx.upload_ptr(host.data_ptr(), cols * n)
for name, fn in (("resident", lambda: ctx.prove_trace_dev(dev, cols, log_n, 2, 3, 32, roots, proof)),
                 ("host input", lambda: ctx.prove_trace_ptr(host.data_ptr(), cols, log_n, 2, 3, 32, roots, proof))):
    for i in range(4):
        torch.cuda.synchronize()
        if i == 3:
            sys.stderr.write("---- %s\n" % name)
            sys.stderr.flush()
        os.environ["X"] = "1"
        fn()
