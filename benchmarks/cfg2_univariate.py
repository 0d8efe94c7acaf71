"""BASELINE config 2: "univariate multiply / interpolate / evaluate microbench at degree 2^16 on 1 B200 vs reference".

GPU arm: Polynomial::mul (mul.rs:6-29), eval_domain (eval.rs:16-21) and interpolate_domain (interpolate.rs:6-44) on the
coset 3 * w^i through the C ABI, host buffers in / out (uint64, the reference's layout), wall-clock per call INCLUDING
the H2D / D2H copies; plus the device-resident transform times from CUDA events (what the roofline talks about).
CPU arm: the oracle's restatement of the reference algorithms (schoolbook O(n*m), Horner O(n*m), Lagrange O(n^3)) on a
BOUNDED size, single thread (the reference is single-threaded), extrapolated to 2^16 with the algorithm's exponent and
labelled as such; results of both arms are compared at the bounded size (parity) and the GPU result at 2^16 is checked
against the O(n log n) CPU NTT checker.  Prints one JSON line per operation."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402  (checker + CPU baseline only)
import stark_rs_b200 as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log-n", type=int, default=16)
ap.add_argument("--cpu-log-mul", type=int, default=14)
ap.add_argument("--cpu-log-eval", type=int, default=13)
ap.add_argument("--cpu-log-interp", type=int, default=9)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
O.build()
ctx = S.Context(0)
k, n = a.log_n, 1 << a.log_n
A, B = O.splitmix64(1, n), O.splitmix64(2, n)
w = O.ff_prim_nth_root(n)


def wall(fn, reps):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        t.append(time.perf_counter() - t0)
    return sorted(t)[len(t) // 2], out


def cpu(fn):
    t0 = time.perf_counter()
    out = fn()
    return time.perf_counter() - t0, out


# ---- Polynomial::mul
t_gpu, prod = wall(lambda: ctx.poly_mul(A, B), a.reps)
assert np.array_equal(prod, O.fast_poly_mul(A, B)), "GPU product differs from the CPU NTT checker"
m = 1 << a.cpu_log_mul
t_cpu, ref = cpu(lambda: O.poly_mul(A[:m], B[:m]))
assert np.array_equal(ctx.poly_mul(A[:m], B[:m]), ref), "GPU product differs from the reference algorithm"
t_cpu_ext = t_cpu * (n / m) ** 2
print(json.dumps({"op": "Polynomial::mul", "len": [n, n], "gpu_ms_host_to_host": t_gpu * 1e3,
                  "cpu_reference_s_at": {"len": m, "s": t_cpu, "threads": 1},
                  "cpu_reference_s_extrapolated_n2": t_cpu_ext, "speedup_vs_extrapolated": t_cpu_ext / t_gpu,
                  "parity": "bit-exact at len %d vs reference algorithm, at len %d vs CPU NTT checker" % (m, n)}))

# ---- eval_domain on the coset (n coefficients -> n points, and -> 4n points)
for logm in (k, k + 2):
    t_gpu, ev = wall(lambda: ctx.poly_eval_coset(A, 3, logm), a.reps)
    assert np.array_equal(ev, O.fast_eval_coset(A, 3, logm))
    e = 1 << a.cpu_log_eval
    we = O.ff_prim_nth_root(e)
    dom = np.array([3 * pow(we, i, O.P) % O.P for i in range(e)], dtype=np.uint64)
    t_cpu, ref = cpu(lambda: O.poly_eval_domain(A[:e], dom))
    assert np.array_equal(ctx.poly_eval_coset(A[:e], 3, a.cpu_log_eval), ref)
    t_cpu_ext = t_cpu * (n / e) * ((1 << logm) / e)
    print(json.dumps({"op": "Polynomial::eval_domain (coset)", "coeffs": n, "points": 1 << logm,
                      "gpu_ms_host_to_host": t_gpu * 1e3, "cpu_reference_s_at": {"coeffs": e, "points": e, "s": t_cpu, "threads": 1},
                      "cpu_reference_s_extrapolated_nm": t_cpu_ext, "speedup_vs_extrapolated": t_cpu_ext / t_gpu,
                      "parity": "bit-exact vs reference algorithm at %d, vs CPU NTT checker at full size" % e}))

# ---- interpolate_domain on the coset
vals = O.splitmix64(3, n)
t_gpu, co = wall(lambda: ctx.poly_interpolate_coset(vals, 3, k), a.reps)
assert np.array_equal(co, O.fast_interpolate_coset(vals, 3, k))
q = 1 << a.cpu_log_interp
wq = O.ff_prim_nth_root(q)
dom = np.array([3 * pow(wq, i, O.P) % O.P for i in range(q)], dtype=np.uint64)
t_cpu, ref = cpu(lambda: O.poly_interpolate_domain(dom, vals[:q]))
assert np.array_equal(ctx.poly_interpolate_coset(vals[:q], 3, a.cpu_log_interp), ref)
t_cpu_ext = t_cpu * (n / q) ** 3
print(json.dumps({"op": "Polynomial::interpolate_domain (coset)", "points": n, "gpu_ms_host_to_host": t_gpu * 1e3,
                  "cpu_reference_s_at": {"points": q, "s": t_cpu, "threads": 1},
                  "cpu_reference_s_extrapolated_n3": t_cpu_ext, "speedup_vs_extrapolated": t_cpu_ext / t_gpu,
                  "parity": "bit-exact vs reference algorithm at %d, vs CPU NTT checker at %d" % (q, n)}))

# ---- device-resident transform times (CUDA events through the library's per-kernel profile)
src, dst = ctx.upload(A), ctx.alloc(4 * n)
for logm in (k, k + 1, k + 2):
    for inverse in (False, True):
        ctx.ntt_dev(src if logm == k else dst, dst, logm, 1, inverse)
        ctx.profile_begin()
        for _ in range(a.reps):
            ctx.ntt_dev(dst, dst, logm, 1, inverse)
        prof = ctx.profile_end()
        ms = sum(p["ms"] for p in prof) / a.reps
        print(json.dumps({"op": "device-resident %s" % ("iNTT" if inverse else "NTT"), "log_n": logm,
                          "us_kernels_only": ms * 1e3, "elems_per_s": (1 << logm) / (ms * 1e-3),
                          "kernels": {p["kernel"]: round(p["ms"] / a.reps * 1e3, 2) for p in prof}}))
src.free(), dst.free()
ctx.close()
