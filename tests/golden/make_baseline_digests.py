#!/usr/bin/env python
"""Generates tests/golden/baseline_digests.json: SHA-256 digests of the ORACLE's outputs at the full BASELINE sizes
(SURVEY 8(d) inputs: splitmix64, seed 0x5354524B + column index), so that the -m gpu parity tests can compare the device
results at those sizes without re-running minutes of CPU work, and so that any later change of the oracle is caught.

  cfg3   1 column x 2^20 rows -> fast_lde (blowup 4, offset 3) -> oracle Fri::prove(w_2^22, 3, ef 4, 32 queries):
         sha256 of the 681 720 proof bytes, column root, top indices
  cfg4   64 columns x 2^22 rows, blowup 2, 8 groups of 8: the 8 group roots (leaf i = Hash::from_field_elements(row i of
         the group's LDE), hash.rs:32-35) and MerkleTree::new over them (merkle.rs:11-38)
  cfg5   one fold+commit round at 2^24 (degenerate domain w = w_2^23, offset 3): root, alpha, sha256 of the folded codeword

The LDE uses oracle/fast_cpu.c (validated against the reference algorithm in tests/test_oracle_fast.py); hashing, Merkle,
Fiat-Shamir, fold and Fri::prove are the restatement of the reference (oracle/stark_oracle.c).  Takes several minutes.

usage: python tests/golden/make_baseline_digests.py [--skip-cfg4]
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as O  # noqa: E402

SEED = 0x5354524B


def cfg3():
    col = O.splitmix64(SEED, 1 << 20)
    lde = O.fast_lde(col, 20, 2, 3)
    r = O.fri_prove(lde, O.ff_prim_nth_root(1 << 22), 3, 4, 32)
    return {"proof_len": len(r["proof"]), "proof_sha256": hashlib.sha256(r["proof"]).hexdigest(),
            "column_root": r["proof"][1:33].hex(), "top_indices": r["top_indices"], "alpha0": r["alphas"][0]}


def cfg4_group_root(k, log_n=22, log_blowup=1, width=8):
    N = 1 << (log_n + log_blowup)
    rows = np.empty((N, width), dtype=np.uint64)
    for c in range(width):
        rows[:, c] = O.fast_lde(O.splitmix64(SEED + k * width + c, 1 << log_n), log_n, log_blowup, 3)
    return O.merkle_commit(O.hash_leaves(rows.reshape(-1), width))


def cfg4():
    roots = [cfg4_group_root(k) for k in range(8)]
    return {"group_roots": [r.hex() for r in roots], "commitment": O.merkle_commit(np.frombuffer(b"".join(roots), dtype=np.uint8)).hex()}


def cfg5(k=24):
    cw = O.splitmix64(SEED + k, 1 << k)
    root = O.merkle_commit(O.hash_leaves(cw))
    alpha = O.fs_challenge(root)
    out = O.fri_fold(cw, alpha, 3, O.ff_prim_nth_root(1 << 23))
    return {"log_n": k, "root": root.hex(), "alpha": alpha, "folded_sha256": hashlib.sha256(out.astype("<u8").tobytes()).hexdigest()}


if __name__ == "__main__":
    O.set_threads(os.cpu_count() or 1)
    out = {"generator": "tests/golden/make_baseline_digests.py (oracle, not the reference binary: no rustc in the image)",
           "seed": SEED}
    path = os.path.join(HERE, "baseline_digests.json")
    if os.path.exists(path):
        out.update(json.load(open(path)))
    for name, fn in (("cfg3", cfg3), ("cfg5", cfg5), ("cfg4", cfg4)):
        if name == "cfg4" and "--skip-cfg4" in sys.argv:
            continue
        t = time.time()
        out[name] = fn()
        print(name, "%.1f s" % (time.time() - t), flush=True)
        json.dump(out, open(path, "w"), indent=1)
