"""Micro-driver: build one Merkle tree over 2^k one-value leaves a few times (for ncu captures / quick timing)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import stark_rs_b200 as S
from stark_rs_b200 import synthetic as G
k = int(sys.argv[1]) if len(sys.argv) > 1 else 22
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = S.Context(0)
buf = ctx.upload(G.splitmix64(1, 1 << k))
for r in range(reps):
    ctx.profile_begin()
    t = ctx.merkle_build_from_buf(buf, 1 << k, 1)
    prof = ctx.profile_end()
    t.free()
print({p["kernel"]: (p["launches"], round(p["ms"], 4)) for p in prof})
