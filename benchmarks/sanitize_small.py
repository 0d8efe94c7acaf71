"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python benchmarks/sanitize_small.py
Sizes are chosen to hit each code path once: multi-pass NTT (2 and 3 passes), LDE, hs2 leaves/levels, the multi-CTA
climb with its fused top (last-CTA ticket), the single-CTA top, the FRI tail, fold+leaf fusion, the query phase."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402  (checker)
import stark_rs_b200 as S  # noqa: E402

O.build()
ctx = S.Context(0)
for log_n in (10, 13, 19):
    c = O.splitmix64(log_n, 1 << log_n)
    ev = ctx.poly_eval_coset(c, 3, log_n)
    assert np.array_equal(ev, O.fast_eval_coset(c, 3, log_n))
    assert np.array_equal(ctx.poly_interpolate_coset(ev, 3, log_n), c)
vals = O.splitmix64(1, 1 << 13)
t = ctx.merkle_build_from_values(vals)
assert t.get_root() == O.merkle_commit(O.hash_leaves(vals))
t.free()
col = O.splitmix64(2, 1 << 11)
roots, proof = ctx.prove_trace(col, 2, 3, 16)
lde = O.fast_lde(col, 11, 2, 3)
assert proof == O.fri_prove(lde, O.ff_prim_nth_root(1 << 13), 3, 4, 16)["proof"]
# verifier (transcript replay, last-codeword tree + iNTT, rounds kernel) and trace ingestion (ragged 32 x 32 tiles)
w13 = O.ff_prim_nth_root(1 << 13)
assert ctx.fri_verify(proof, w13, 3, 1 << 13, 4, 16) == (True, "")
assert ctx.fri_verify(proof[:-7], w13, 3, 1 << 13, 4, 16)[0] is False
rows = [[(r * 7 + c) * (-1) ** r for c in range(5)] for r in range(1 << 7)]
got = ctx.trace_to_columns(rows).download().reshape(5, -1)
assert np.array_equal(got, O.trace_columns(rows) % np.uint64(998244353))
roots2, proof2 = ctx.prove_trace_rows(rows, 2, 3, 8)
assert ctx.fri_verify(proof2, O.ff_prim_nth_root(1 << 9), 3, 1 << 9, 4, 8)[0]
big = O.splitmix64(3, 1 << 18)
t = ctx.merkle_build_from_values(big)      # one throughput level launch + climb
assert t.get_root() == O.merkle_commit(O.hash_leaves(big))
t.free()
ctx.close()
print("sanitize_small ok")
