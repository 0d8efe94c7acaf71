"""Oracle vs (i) the reference's PROPERTY tests for hash / merkle / fri and (ii) the survey-derived
vectors (tests/golden/survey_vectors.json) plus a third, independent pure-Python restatement of
hash.rs kept in this file.  The reference pins no digest itself ("parity unpinned", oracle header)."""
import hashlib
import json
import os

import numpy as np
import pytest

P = 998244353
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "survey_vectors.json")))

PRIMES = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53]
RC = [0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1b, 0x36, 0x6c, 0xd8, 0xab, 0x4d, 0x9a, 0x2f,
      0x5e, 0xbc, 0x63, 0xc6, 0x97, 0x35, 0x6a, 0xd4, 0xb3, 0x7d, 0xfa, 0xef, 0xc5, 0x91, 0x39, 0x72]


def py_hash(msg):
    """Independent restatement of hash.rs:7-30, 59-94 written from the source text (pure Python)."""
    rotl = lambda b, n: ((b << n) | (b >> (8 - n))) & 0xFF
    s = [PRIMES[i % 16] for i in range(32)]

    def mix():
        for i in range(32):
            s[i] = rotl((s[i] * 251) & 0xFF, 1) ^ 0x63
        for g in range(8):
            t0, t1, t2, t3 = s[4 * g:4 * g + 4]
            s[4 * g:4 * g + 4] = [t0 ^ t1 ^ t3, t0 ^ t2 ^ t3, t0 ^ t1 ^ t2, t1 ^ t2 ^ t3]
        for i in range(32):
            s[i] = (s[i] + s[(i + 1) % 32] + s[31 if i == 0 else i - 1]) & 0xFF
        for i in range(32):
            s[i] = (s[i] + RC[i]) & 0xFF

    for off in range(0, len(msg), 32):
        for i, b in enumerate(msg[off:off + 32]):
            s[i] = rotl((s[i] + b) & 0xFF, 3)
            s[(i + 7) % 32] ^= s[i]
        mix()
    for _ in range(8):
        mix()
    return bytes(s)


def test_hash_vectors(oracle):
    O = oracle
    for msg_hex, digest in G["hash_from_bytes"].items():
        assert O.hash_from_bytes(bytes.fromhex(msg_hex)).hex() == digest
    assert O.hash_from_u64(0).hex() == G["hash_from_u64_0"]
    assert O.hash_from_field_elements([1]).hex() == G["hash_from_field_elements_1"]
    assert O.hash_combine(bytes(32), bytes(32)).hex() == G["combine_zero_zero"]
    assert [O.sbox(i) for i in range(4)] == G["sbox_0_4"]
    assert sorted(O.sbox(i) for i in range(256)) == list(range(256))  # permutation


def test_hash_vs_pure_python(oracle):
    O = oracle
    rng = np.random.default_rng(7)
    for n in [0, 1, 5, 8, 31, 32, 33, 36, 63, 64, 65, 96, 100, 257]:
        msg = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert O.hash_from_bytes(msg) == py_hash(msg), n


def test_hash_reference_properties(oracle):
    O = oracle
    h1, h2 = O.hash_from_bytes(b"hello"), O.hash_from_bytes(b"hello")
    assert h1 == h2                                                           # hash.rs:107-111
    assert h1 != O.hash_from_bytes(b"world")                                   # hash.rs:114-118
    ha = O.hash_from_bytes(b"hallo")
    assert sum(a != b for a, b in zip(h1, ha)) > 10                            # hash.rs:121-132
    assert len(O.hash_from_field_elements([1, 2, 3, 4, 5])) == 32              # hash.rs:135-139
    l, r = O.hash_from_bytes(b"left"), O.hash_from_bytes(b"right")
    c = O.hash_combine(l, r)
    assert c != l and c != r                                                   # hash.rs:142-149
    assert c == O.hash_from_bytes(l + r)                                       # hash.rs:41-46
    v = 0x0123456789ABCDEF
    assert O.hash_from_u64(v) == O.hash_from_bytes(v.to_bytes(8, "little"))    # hash.rs:32-39


def _leaves(O, n):
    return np.array([list(O.hash_from_bytes(bytes([i]))) for i in range(n)], dtype=np.uint8)


def test_merkle(oracle):
    O = oracle
    l4, l8 = _leaves(O, 4), _leaves(O, 8)
    assert O.merkle_commit(l4).hex() == G["merkle_root_4"]
    assert O.merkle_commit(l8).hex() == G["merkle_root_8"]
    nodes = O.merkle_build(l4)
    assert len(nodes) == 7 and nodes[-1].tobytes() == O.merkle_commit(l4)     # merkle.rs:103-108 (3 levels)
    root = O.merkle_commit(l8)
    for i in range(8):                                                         # merkle.rs:111-122
        proof = O.merkle_open(l8, i)
        assert len(proof) == 3
        assert O.merkle_verify(l8[i], i, proof, root)
    proof = O.merkle_open(l4, 0)                                               # merkle.rs:125-133
    assert not O.merkle_verify(O.hash_from_bytes(bytes([99])), 0, proof, O.merkle_commit(l4))
    assert O.merkle_commit(l4[:1]) == l4[0].tobytes()                          # single leaf: root = leaf


def test_field_vectors(oracle):
    O = oracle
    for k, w in G["roots_of_unity"].items():
        assert O.ff_prim_nth_root(1 << int(k)) == w
    assert O.ff_inv(2) == G["inv2"] and O.ff_inv(3) == G["inv3"]
    assert O.ff_sample([1, 2, 3]) == G["sample_1_2_3"]
    m = G["montgomery"]
    R = 1 << 32
    assert (-pow(P, -1, R)) % R == m["neg_pinv_mod_R"] and R % P == m["R_mod_p"] and R * R % P == m["R2_mod_p"]


def _statement(O, n, offset, coeffs):
    w = O.ff_prim_nth_root(n)
    dom = [O.ff_mul(offset, O.ff_exp(w, i)) for i in range(n)]                  # fri.rs:575-578
    return w, O.poly_eval_domain(coeffs, dom)


@pytest.mark.parametrize("case", G["fri_proofs"], ids=lambda c: "n%d" % c["n"])
def test_fri_prove_verify(oracle, case):
    """fri.rs:532-693: prove -> serialize -> deserialize -> verify == true; bytes vs survey vectors."""
    O = oracle
    n, off, ef, nq = case["n"], case["offset"], case["ef"], case["nq"]
    w, cw = _statement(O, n, off, case["coeffs"])
    r = O.fri_prove(cw, w, off, ef, nq)
    ok, why = O.fri_verify(r["proof"], w, off, n, ef, nq)
    assert ok, why
    assert len(r["proof"]) == case["bytes"] and O.stream_count(r["proof"]) == case["objects"]
    assert hashlib.sha256(r["proof"]).hexdigest() == case["sha256"]
    assert r["top_indices"] == case["top"]
    # tampering with a codeword value or a path byte must break verification
    bad = bytearray(r["proof"])
    bad[-1] ^= 1
    assert not O.fri_verify(bytes(bad), w, off, n, ef, nq)[0]


def test_fri_test1_details(oracle):
    O = oracle
    t = G["fri_test1"]
    w, cw = _statement(O, 32, 3, [5])
    r = O.fri_prove(cw, w, 3, 4, 2)
    assert r["rounds"] == t["rounds"]
    proof = r["proof"]
    assert proof[0] == 0 and proof[1:33].hex() == t["roots"][0] and proof[33] == 0 and proof[34:66].hex() == t["roots"][1]
    assert r["alphas"] == [t["alpha0_raw"]] and r["alphas"][0] % P == t["alpha0_mod_p"]
    assert r["seed_challenge"] == t["seed_challenge_raw"] and r["top_indices"] == t["top_indices"]
    # transcript semantics (fiat_shamir.rs:19-25): alpha0 = LE u64 of Hash(root0)[0..8], unreduced
    assert O.fs_challenge(bytes.fromhex(t["roots"][0])) == t["alpha0_raw"]


def test_fri_rejects_high_degree(oracle):
    """A codeword that is not low degree must fail Fri::verify (fri.rs:389-397)."""
    O = oracle
    n, off = 64, 7
    w, cw = _statement(O, n, off, list(range(1, 40)))    # degree 38 > 64/4 - 1
    r = O.fri_prove(cw, w, off, 4, 3)
    assert not O.fri_verify(r["proof"], w, off, n, 4, 3)[0]


def test_fold_matches_closed_form(oracle):
    O = oracle
    rng = np.random.default_rng(3)
    for n in (2, 4, 64, 512):
        w = O.ff_prim_nth_root(n)
        cw = rng.integers(0, P, n, dtype=np.uint64)
        alpha = int(rng.integers(0, 2**63, dtype=np.uint64)) * 2 + 1    # unreduced, like fiat_shamir.rs:21-24
        assert np.array_equal(O.fri_fold(cw, alpha, 3, w), O.fast_fri_fold(cw, alpha, 3, w))


def test_trace_fibonacci(oracle):
    col = oracle.trace_fibonacci(64)                                        # trace.rs:36-49
    assert list(col[:6]) == [1, 1, 2, 3, 5, 8] and int(col[63]) == 10610209857723
    with pytest.raises(oracle.OraclePanic):
        oracle.trace_fibonacci(200)                                         # i128 overflow (debug panic)


def test_trace_columns(oracle):
    """trace.rs:21-34: `e as u64` keeps the low 64 bits of the two's-complement i128; get_col transposes"""
    rows = [[1, -1, 1 << 64], [(1 << 127) - 1, -(1 << 127), 998244353], [7, (1 << 64) + 5, -998244353]]
    cols = oracle.trace_columns(rows)
    assert cols.shape == (3, 3)
    m = (1 << 64) - 1
    assert [int(x) for x in cols[0]] == [1, m, 7]
    assert [int(x) for x in cols[1]] == [m, 0, 5]
    assert [int(x) for x in cols[2]] == [0, 998244353, (-998244353) & m]
    fib = [[int(v)] for v in oracle.trace_fibonacci(64)]                    # Trace::fibonacci(64).get_col(0)
    assert np.array_equal(oracle.trace_columns(fib)[0], oracle.trace_fibonacci(64))



def _emitter_shaped_json(O):
    """what tools/emit_golden.rs prints, computed with the oracle instead of the reference crate"""
    def fri_case(n, off, ef, nq, coeffs):
        w = O.ff_prim_nth_root(n)
        dom = [O.ff_mul(off, O.ff_exp(w, i)) for i in range(n)]
        cw = [int(x) for x in O.poly_eval_domain(coeffs, dom)]
        r = O.fri_prove(cw, w, off, ef, nq)
        return {"n": n, "offset": off, "ef": ef, "nq": nq, "coeffs": coeffs, "codeword": cw, "bytes": len(r["proof"]),
                "top": r["top_indices"], "verify": True, "proof_hex": r["proof"].hex()}
    msgs = [b"hello", b"world", b"", b"\0", bytes([7] * 32), bytes([9] * 33)]
    ref = {"hash_from_bytes": {m.hex(): O.hash_from_bytes(m).hex() for m in msgs},
           "hash_from_u64_0": O.hash_from_u64(0).hex(), "hash_from_field_elements_1": O.hash_from_field_elements([1]).hex(),
           "hash_from_field_elements_8": O.hash_from_field_elements([1, 2, 3, 4, 5, 6, 7, 998244352]).hex(),
           "combine_zero_zero": O.hash_combine(bytes(32), bytes(32)).hex(),
           "challenge_empty": O.fs_challenge(b""), "challenge_stark_rs": O.fs_challenge(b"stark-rs"),
           "roots_of_unity": {str(k): O.ff_prim_nth_root(1 << k) for k in (3, 16, 20, 22, 23)}}
    for n in (4, 8, 16):
        leaves = np.stack([np.frombuffer(O.hash_from_bytes(bytes([i])), dtype=np.uint8) for i in range(n)])
        ref["merkle_root_%d" % n] = O.merkle_commit(leaves).hex()
        ref["merkle_open_%d_%d" % (n, n - 3)] = [bytes(h).hex() for h in O.merkle_open(leaves, n - 3)]
    cw = [(i * 1234567 + 89) % 998244353 for i in range(16)]
    a = 15764728482632548394
    ref["fold_16"] = {"codeword": cw, "alpha_raw": a, "folded": [int(x) for x in O.fri_fold(cw, a, 3, O.ff_prim_nth_root(16))]}
    ref["sample_indices_seed_abc"] = [int(x) for x in O.fri_sample_indices(b"abc", 64, 8, 5)]
    ref["fri_proofs"] = [fri_case(32, 3, 4, 2, [5]), fri_case(64, 7, 4, 3, [5, 3]), fri_case(128, 13, 4, 4, [1, 3, 2]),
                         fri_case(256, 17, 8, 5, [1, 2, 5, 3, 7, 4, 1, 2])]
    return ref


def test_reference_golden_checker(oracle):
    """tools/check_golden.py (the checker for the output of tools/emit_golden.rs, which only someone with cargo can run):
    it accepts an emitter-shaped JSON built from the oracle, catches a flipped byte, agrees with the survey vectors -- and
    if a reference-emitted file has been committed (tests/golden/reference_emitted.json) the oracle must reproduce it."""
    import importlib.util
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("check_golden", os.path.join(root, "tools", "check_golden.py"))
    C = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(C)
    ref = _emitter_shaped_json(oracle)
    assert C.check(ref, oracle, verbose=False) == [] and C.check_survey(ref) == []
    broken = json.loads(json.dumps(ref))
    broken["fri_proofs"][2]["proof_hex"] = broken["fri_proofs"][2]["proof_hex"][:-2] + "00"
    broken["combine_zero_zero"] = "00" * 32
    assert len(C.check(broken, oracle, verbose=False)) == 2
    emitted = os.path.join(root, "tests", "golden", "reference_emitted.json")
    if os.path.exists(emitted):
        real = json.load(open(emitted))
        assert C.check(real, oracle, verbose=False) == [] and C.check_survey(real) == []
