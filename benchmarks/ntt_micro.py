"""NTT / LDE micro-benchmark (BASELINE config 2 sizes and the LDE of config 3): CUDA events on the context's stream.

usage: python benchmarks/ntt_micro.py [--logs 13,16,...] [--batch B] [--reps R] [--cold]
Prints one JSON line per size: microseconds per transform, field elements/s, and the fraction of the measured HBM
peak using the ALGORITHMIC bytes of DESIGN.md (2 passes x 8N bytes for 2^13 <= N <= 2^23)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import stark_rs_b200 as S  # noqa: E402
from stark_rs_b200 import synthetic as G  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--logs", default="13,14,15,16,17,18,19,20,21,22,23")
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--cold", action="store_true", help="evict L2 (256 MiB write) before every timed transform")
ap.add_argument("--lde", action="store_true", help="also time the coset LDE n -> 4n (config 3) for log_n <= 21")
a = ap.parse_args()

peak = 6542.7
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
stream = torch.cuda.Stream()
ctx = S.Context(0, stream=stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    with torch.cuda.stream(stream):
        for _ in range(3):
            fn()
        for e0, e1 in ev:
            if a.cold:
                flush.zero_()
            e0.record(stream)
            fn()
            e1.record(stream)
    torch.cuda.synchronize()
    t = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
    return t[len(t) // 2] * 1e3, t[0] * 1e3   # median, best (us)


for k in [int(x) for x in a.logs.split(",")]:
    n = 1 << k
    src = ctx.upload(G.splitmix64(7 + k, n * a.batch))
    dst = ctx.alloc(n * a.batch)
    for inverse in (False, True):
        med, best = timed(lambda: ctx.ntt_dev(src, dst, k, a.batch, inverse), a.reps)
        passes = 1 if k <= 12 else 2
        alg = passes * 8.0 * n * a.batch
        print(json.dumps({"op": "intt" if inverse else "ntt", "log_n": k, "batch": a.batch, "us": round(med, 2),
                          "us_best": round(best, 2), "elems_per_s": n * a.batch / (med * 1e-6),
                          "algorithmic_bytes": alg, "achieved_gbs": alg / (med * 1e-6) / 1e9,
                          "hbm_frac_of_measured": alg / (med * 1e-6) / 1e9 / peak, "cold_l2": a.cold}))
    if a.lde and k <= 21:
        out = ctx.alloc(4 * n * a.batch)
        med, best = timed(lambda: ctx.lde_dev(src, a.batch, k, 2, 3, out), a.reps)
        alg = (20 + 12 * 4) * n * a.batch
        print(json.dumps({"op": "lde_x4", "log_n": k, "batch": a.batch, "us": round(med, 2), "us_best": round(best, 2),
                          "out_elems_per_s": 4 * n * a.batch / (med * 1e-6), "algorithmic_bytes": alg,
                          "achieved_gbs": alg / (med * 1e-6) / 1e9,
                          "hbm_frac_of_measured": alg / (med * 1e-6) / 1e9 / peak, "cold_l2": a.cold}))
        out.free()
    src.free(), dst.free()
ctx.close()
