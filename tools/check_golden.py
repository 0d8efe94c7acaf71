#!/usr/bin/env python3
"""Compares the JSON printed by tools/emit_golden.rs (run inside the REAL stark-rs crate by someone with cargo) with this
repository's oracle and with tests/golden/survey_vectors.json.  A clean run turns the hash / Merkle / Fiat-Shamir / FRI
rows of the oracle from "parity unpinned" to pinned; the file can then be committed as tests/golden/reference_emitted.json
(tests/test_oracle_survey.py picks it up when present).

usage: python tools/check_golden.py reference_golden.json"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def check(ref, O=None, verbose=True):
    """-> list of mismatches (empty = the oracle reproduces everything the reference emitted)"""
    if O is None:
        import oracle as O
    bad = []

    def eq(name, got, want):
        if got != want:
            bad.append("%s: oracle %r != reference %r" % (name, got, want))
        elif verbose:
            print("ok  ", name)

    for m, d in ref["hash_from_bytes"].items():
        eq("Hash::from_bytes(%s)" % (m or "empty"), O.hash_from_bytes(bytes.fromhex(m)).hex(), d)
    eq("Hash::from_u64(0)", O.hash_from_u64(0).hex(), ref["hash_from_u64_0"])
    eq("Hash::from_field_elements([1])", O.hash_from_field_elements([1]).hex(), ref["hash_from_field_elements_1"])
    eq("Hash::from_field_elements(8 values)", O.hash_from_field_elements([1, 2, 3, 4, 5, 6, 7, 998244352]).hex(),
       ref["hash_from_field_elements_8"])
    eq("Hash::combine(0, 0)", O.hash_combine(bytes(32), bytes(32)).hex(), ref["combine_zero_zero"])
    for n in (4, 8, 16):
        leaves = np.stack([np.frombuffer(O.hash_from_bytes(bytes([i])), dtype=np.uint8) for i in range(n)])
        eq("MerkleTree root, %d leaves" % n, O.merkle_commit(leaves).hex(), ref["merkle_root_%d" % n])
        eq("MerkleTree::open(%d) of %d" % (n - 3, n), [bytes(h).hex() for h in O.merkle_open(leaves, n - 3)],
           ref["merkle_open_%d_%d" % (n, n - 3)])
    eq("FiatShamir::challenge(empty)", O.fs_challenge(b""), ref["challenge_empty"])
    eq("FiatShamir::challenge('stark-rs')", O.fs_challenge(b"stark-rs"), ref["challenge_stark_rs"])
    for k, v in ref["roots_of_unity"].items():
        eq("prim_nth_root(2^%s)" % k, O.ff_prim_nth_root(1 << int(k)), v)
    f = ref["fold_16"]
    eq("fold_codeword(16, unreduced alpha)", [int(x) for x in O.fri_fold(f["codeword"], f["alpha_raw"], 3, O.ff_prim_nth_root(16))], f["folded"])
    eq("sample_indices('abc', 64, 8, 5)", [int(x) for x in O.fri_sample_indices(b"abc", 64, 8, 5)], ref["sample_indices_seed_abc"])
    for c in ref["fri_proofs"]:
        name = "Fri::prove(n=%d, offset=%d, ef=%d, nq=%d)" % (c["n"], c["offset"], c["ef"], c["nq"])
        w = O.ff_prim_nth_root(c["n"])
        r = O.fri_prove(c["codeword"], w, c["offset"], c["ef"], c["nq"])
        eq(name + " proof bytes", r["proof"].hex(), c["proof_hex"])
        eq(name + " top indices", r["top_indices"], c["top"])
        eq(name + " reference verify", c["verify"], True)
    return bad


def check_survey(ref):
    """the survey-derived vectors against the reference-emitted ones (they were a second opinion; now they are checked)"""
    sv = json.load(open(os.path.join(ROOT, "tests", "golden", "survey_vectors.json")))
    bad = []
    for m, d in sv["hash_from_bytes"].items():
        if ref["hash_from_bytes"].get(m) != d:
            bad.append("survey hash_from_bytes[%s]" % m)
    for a, b in zip(sv["fri_proofs"], ref["fri_proofs"]):
        if hashlib.sha256(bytes.fromhex(b["proof_hex"])).hexdigest() != a["sha256"] or a["top"] != b["top"]:
            bad.append("survey fri proof n=%d" % a["n"])
    for k in ("merkle_root_4", "merkle_root_8", "combine_zero_zero", "hash_from_u64_0", "hash_from_field_elements_1"):
        if sv[k] != ref[k]:
            bad.append("survey " + k)
    return bad


if __name__ == "__main__":
    ref = json.load(open(sys.argv[1]))
    bad = check(ref) + check_survey(ref)
    for b in bad:
        print("MISMATCH", b)
    print("%d mismatches" % len(bad))
    sys.exit(1 if bad else 0)
