"""FRI fold micro-benchmark (BASELINE config 5, fold part): stark_fri_fold_dev at 2^16..2^26 inputs, CUDA events on the
context's stream, L2 evicted before every timed launch.  Algorithmic bytes = 12 per output."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import stark_rs_b200 as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="16,18,20,22,24,26")
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
peak = 6542.7
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
stream = torch.cuda.Stream()
ctx = S.Context(0, stream=stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for k in [int(x) for x in a.sizes.split(",")]:
    n = 1 << k
    omega = S.prim_nth_root(1 << min(k, 23))
    cw = torch.randint(0, S.P, (n,), dtype=torch.int32, device="cuda")
    out = torch.empty(n // 2, dtype=torch.int32, device="cuda")
    src, dst = ctx.wrap(cw.data_ptr(), n), ctx.wrap(out.data_ptr(), n // 2)
    for cold in (False, True):
        ts = []
        with torch.cuda.stream(stream):
            for i in range(a.reps + 3):
                if cold:
                    flush.zero_()
                ctx.profile_begin()
                ctx.fri_fold_dev(src, n, 15764728482632548394, 3, omega, dst)
                p = [x for x in ctx.profile_end() if x["kernel"] == "fri_fold"]
                if i >= 3:
                    ts.append(p[0]["ms"])
        ts.sort()
        ms = ts[len(ts) // 2]
        batched_us = None
        if not cold:
            # 10 launches between one event pair: removes the ~6 us event/launch floor of a single timed launch
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                for _ in range(10):
                    ctx.fri_fold_dev(src, n, 15764728482632548394, 3, omega, dst)
                e1.record(stream)
            torch.cuda.synchronize()
            batched_us = e0.elapsed_time(e1) * 100.0
        alg = 12.0 * (n // 2)
        print(json.dumps({"op": "fri_fold", "log_n": k, "cold_l2": cold, "us": ms * 1e3, "algorithmic_bytes": alg,
                          "achieved_gbs": alg / (ms * 1e-3) / 1e9, "hbm_frac_of_measured": alg / (ms * 1e-3) / 1e9 / peak,
                          "us_per_launch_10_back_to_back": batched_us,
                          "hbm_frac_back_to_back": (alg / (batched_us * 1e-6) / 1e9 / peak) if batched_us else None}))
ctx.close()
