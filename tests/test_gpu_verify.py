"""GPU parity: Fri::verify (fri.rs:313-505) on the device against the oracle's restatement -- same verdict AND the same
`println!` reason on honest proofs, tampered proofs, truncated streams and wrong parameters."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P = 998244353
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "survey_vectors.json")))


def rf(seed, n):
    return np.random.default_rng(seed).integers(0, P, n, dtype=np.uint64)


def low_degree_codeword(O, n, ef, seed, offset=3):
    return O.fast_eval_coset(rf(seed, n // ef), offset, n.bit_length() - 1)


def parse(proof):
    """ProofStream::deserialize layout (stream.rs:66-168) of an honest proof: [(tag, count, header_off, payload_off)]"""
    objs, i = [], 0
    while i < len(proof):
        tag = proof[i]
        if tag == 0:
            objs.append((0, 1, i, i + 1))
            i += 33
        else:
            cnt = int.from_bytes(proof[i + 1:i + 9], "little")
            objs.append((tag, cnt, i, i + 9))
            i += 9 + cnt * (8 if tag == 2 else 32)
    return objs


def same_verdict(ctx, O, S, proof, w, off, n, ef, nq):
    """device verdict == oracle verdict (bool and reason); a reference panic must surface as StarkPanic"""
    try:
        want = O.fri_verify(proof, w, off, n, ef, nq)
    except O.OraclePanic as e:
        with pytest.raises(S.StarkPanic):
            ctx.fri_verify(proof, w, off, n, ef, nq)
        return ("panic", str(e))
    got = ctx.fri_verify(proof, w, off, n, ef, nq)
    assert got == want, (got, want)
    return want


@pytest.mark.parametrize("case", G["fri_proofs"], ids=lambda c: "n%d" % c["n"])
def test_verify_reference_statements(ctx, oracle, case):
    """the four prove -> verify statements of fri.rs:532-693, verified on the device"""
    n, off, ef, nq = case["n"], case["offset"], case["ef"], case["nq"]
    w = oracle.ff_prim_nth_root(n)
    dom = [oracle.ff_mul(off, oracle.ff_exp(w, i)) for i in range(n)]
    cw = oracle.poly_eval_domain(case["coeffs"], dom)
    proof, top = ctx.fri_prove(cw, off, w, ef, nq)
    ok, why, d = ctx.fri_verify(proof, w, off, n, ef, nq, details=True)
    assert ok and why == ""
    assert d["top"] == top == case["top"]
    # polynomial_values (fri.rs:437-441): (a, cw[a]), (b, cw[b]) per query, in query order
    want = []
    for t in top:
        a = t % (n // 2)
        want += [(a, int(cw[a])), (a + n // 2, int(cw[a + n // 2]))]
    assert d["polynomial_values"] == want
    roots = [proof[o[3]:o[3] + 32] for o in parse(proof) if o[0] == 0]
    assert [d["roots"][r].tobytes() for r in range(len(roots))] == roots
    if n == 32:
        assert [r.hex() for r in roots] == G["fri_test1"]["roots"]


@pytest.mark.parametrize("log_n,ef,nq", [(5, 4, 2), (8, 8, 5), (10, 4, 8), (13, 4, 16), (16, 4, 32)])
def test_verify_honest_proofs(ctx, oracle, S, log_n, ef, nq):
    n = 1 << log_n
    w = oracle.ff_prim_nth_root(n)
    for off in (3, 7):
        proof, _ = ctx.fri_prove(low_degree_codeword(oracle, n, ef, log_n, off), off, w, ef, nq)
        assert same_verdict(ctx, oracle, S, proof, w, off, n, ef, nq) == (True, "")


def test_verify_large_proof(ctx, oracle):
    """BASELINE config 3 size: the 2^22-point proof (15 rounds, 32 queries, 682 KB) verified on the device"""
    n = 1 << 22
    w = oracle.ff_prim_nth_root(n)
    proof, _ = ctx.fri_prove(low_degree_codeword(oracle, n, 4, 22), 3, w, 4, 32)
    assert ctx.fri_verify(proof, w, 3, n, 4, 32) == (True, "")
    bad = bytearray(proof)
    bad[len(bad) - 5] ^= 1                 # last authentication path of the last round
    assert ctx.fri_verify(bytes(bad), w, 3, n, 4, 32) == (False, "merkle authentication path verification fails for cc")


def test_verify_rejects_high_degree(ctx, oracle, S):
    """a consistent FRI transcript of a codeword that is NOT low degree fails only at the degree check (fri.rs:393-399)"""
    n, ef, nq = 1 << 9, 4, 4
    w = oracle.ff_prim_nth_root(n)
    proof, _ = ctx.fri_prove(rf(5, n), 3, w, ef, nq)
    assert same_verdict(ctx, oracle, S, proof, w, 3, n, ef, nq) == \
        (False, "last codeword does not correspond to polynomial of low enough degree")


def test_verify_tampered_proofs(ctx, oracle, S):
    """one mutation at a time, every kind of object: verdict and reason must equal the oracle's"""
    n, ef, nq, off = 1 << 10, 4, 6, 3
    w = oracle.ff_prim_nth_root(n)
    proof, _ = ctx.fri_prove(low_degree_codeword(oracle, n, ef, 77), off, w, ef, nq)
    objs = parse(proof)
    R = sum(1 for o in objs if o[0] == 0)
    rng = np.random.default_rng(11)
    seen = set()

    def run(mut):
        seen.add(same_verdict(ctx, oracle, S, bytes(mut), w, off, n, ef, nq)[1])

    for k, (tag, cnt, hdr, pay) in enumerate(objs):
        size = 32 if tag in (0, 3) else 8
        if tag == 2 and k > R and rng.random() < 0.5:
            continue                       # half of the triples, all roots / paths / the last codeword
        m = bytearray(proof)
        item = int(rng.integers(0, cnt))
        byte = int(rng.integers(0, 4 if tag == 2 else 32))     # low bytes of a value stay below 2^32
        m[pay + size * item + byte] ^= 1 << int(rng.integers(0, 8))
        run(m)
    # unknown tag: the stream ends there (stream.rs `_ => break`)
    for k in (0, R - 1, R, R + 1, R + nq, R + nq + 1, R + nq + 2, len(objs) - 1):
        m = bytearray(proof)
        m[objs[k][2]] = 9
        run(m)
    # a triple that is not a triple (count 2): the third value's first byte then reads as a tag -- pick one that ends the stream
    for k in range(R + 1, R + 1 + nq):
        if proof[objs[k][3] + 16] > 3:
            m = bytearray(proof)
            m[objs[k][2] + 1] = 2
            run(m)
            break
    # truncations: inside a root, the last codeword, a triple, a path, and exactly at object boundaries
    for cut in (0, 1, 20, 33 * R + 3, 33 * R + 9 + 8 * 5, objs[R + 1][2], objs[R + 1][3] + 9, objs[R + nq + 1][3] + 40,
                objs[-1][2], len(proof) - 1, len(proof) - 32):
        run(bytearray(proof[:cut]))
    # values that are congruent but not canonical: colinearity holds, the leaf hash of the raw value does not
    m = bytearray(proof)
    a = int.from_bytes(proof[objs[R + 1][3]:objs[R + 1][3] + 8], "little")
    m[objs[R + 1][3]:objs[R + 1][3] + 8] = (a + P).to_bytes(8, "little")
    run(m)
    # a raw value the reference's u128 subtraction underflows on (ff.rs:154-160): debug panic
    m = bytearray(proof)
    m[objs[R + 3][3]:objs[R + 3][3] + 8] = ((1 << 64) - 1).to_bytes(8, "little")
    run(m)
    assert {"Failed to extract Merkle root", "Failed to extract last codeword", "last codeword is not well formed",
            "Failed to extract triple values", "Expected triple of values", "colinearity check failure",
            "merkle authentication path verification fails for aa", "merkle authentication path verification fails for bb",
            "merkle authentication path verification fails for cc", "Failed to extract path for aa"} <= seen, seen


def test_verify_wrong_parameters(ctx, oracle, S):
    n, ef, nq = 1 << 8, 4, 4
    w = oracle.ff_prim_nth_root(n)
    proof, _ = ctx.fri_prove(low_degree_codeword(oracle, n, ef, 3), 3, w, ef, nq)
    assert same_verdict(ctx, oracle, S, proof, w, 3, n, ef, nq) == (True, "")
    assert not same_verdict(ctx, oracle, S, proof, w, 5, n, ef, nq)[0]          # other coset
    assert not same_verdict(ctx, oracle, S, proof, w, 3, n, ef, nq + 1)[0]      # other query count
    assert not same_verdict(ctx, oracle, S, proof, w, 3, n, 8, nq)[0]           # other expansion factor
    w2 = oracle.ff_prim_nth_root(2 * n)
    assert not same_verdict(ctx, oracle, S, proof, w2, 3, 2 * n, ef, nq)[0]     # other domain
    for bad, msg in (((proof, w, 3, 100, ef, nq), "Domain length must be power of 2"),
                     ((proof, w, 3, n, 3, nq), "Expansion factor must be power of 2"),
                     ((proof, w, 3, n, 2, nq), "Expansion factor must be at least 4")):
        with pytest.raises(S.StarkPanic, match=msg):                             # Fri::new asserts, fri.rs:37-45
            ctx.fri_verify(*bad)
    with pytest.raises(S.StarkPanic, match="unsupported domain"):
        ctx.fri_verify(proof, oracle.ff_mul(w, w), 3, n, ef, nq)
    assert ctx.fri_verify(b"", w, 3, n, ef, nq) == (False, "Failed to extract Merkle root")


def test_verify_with_transcript_prefix(ctx, oracle):
    """the caller's FiatShamir state before prove / verify (fiat_shamir.rs:4-13), chunk-aligned or not"""
    n, ef, nq = 1 << 9, 4, 4
    w = oracle.ff_prim_nth_root(n)
    cw = low_degree_codeword(oracle, n, ef, 9)
    for prefix in (b"abc", bytes(range(32)), bytes(range(45))):
        proof, _ = ctx.fri_prove(cw, 3, w, ef, nq, transcript=prefix)
        assert ctx.fri_verify(proof, w, 3, n, ef, nq, transcript=prefix) == (True, "")
        assert not ctx.fri_verify(proof, w, 3, n, ef, nq)[0]
