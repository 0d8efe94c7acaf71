#!/usr/bin/env python3
"""Fri::verify on the device (stark_fri_verify): wall time per call from host proof bytes to verdict, next to the oracle's
CPU verifier (reference algorithm, 1 thread) on the same proof.  One JSON line per size."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402  (the checker and the CPU baseline, never the product path)
import stark_rs_b200 as S  # noqa: E402

P = 998244353


def main():
    ctx = S.Context(0)
    for log_n, nq, cpu in ((12, 16, True), (16, 32, True), (20, 32, True), (22, 32, True)):
        n, ef = 1 << log_n, 4
        w = O.ff_prim_nth_root(n)
        coeffs = np.random.default_rng(log_n).integers(0, P, n // ef, dtype=np.uint64)
        cw = O.fast_eval_coset(coeffs, 3, log_n)
        proof, _ = ctx.fri_prove(cw, 3, w, ef, nq)
        assert ctx.fri_verify(proof, w, 3, n, ef, nq) == (True, "")
        l0 = ctx.launches
        t = []
        for _ in range(10):
            t0 = time.perf_counter()
            ok, _why = ctx.fri_verify(proof, w, 3, n, ef, nq)
            t.append(time.perf_counter() - t0)
        launches = (ctx.launches - l0) // 10
        rec = {"op": "Fri::verify", "log_n": log_n, "nq": nq, "proof_bytes": len(proof), "gpu_ms_host_to_verdict": 1e3 * float(np.median(t)),
               "gpu_ms_best": 1e3 * min(t), "launches": int(launches), "ok": bool(ok)}
        if cpu:
            t0 = time.perf_counter()
            okc, _ = O.fri_verify(proof, w, 3, n, ef, nq)
            rec["cpu_reference_ms_1_thread"] = 1e3 * (time.perf_counter() - t0)
            assert okc
        print(json.dumps(rec), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
