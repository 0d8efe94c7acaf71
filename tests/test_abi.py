"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/stark_b200.h declares; host-only entry points behave like the reference; and the product fails
LOUDLY without a GPU instead of falling back to the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def S():
    import stark_rs_b200 as S
    S.build_library()
    return S


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "stark_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(stark_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(S):
    names = _declared_symbols()
    assert len(names) > 50
    lib = ctypes.CDLL(S.lib_path)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_header_cites_reference():
    text = open(os.path.join(ROOT, "include", "stark_b200.h")).read()
    for cite in ["ff.rs:138", "mul.rs:6-29", "eval.rs:16-21", "interpolate.rs:6-44", "hash.rs:7-30",
                 "merkle.rs:11-38", "merkle.rs:67-80", "fri.rs:57-91", "fri.rs:105-156", "fri.rs:250-311",
                 "stream.rs:35-64", "fri.rs:313-505", "merkle.rs:82-97", "trace.rs:4-34"]:
        assert cite in text, cite


def test_host_only_entry_points(S, oracle):
    for k in range(0, 24):
        assert S.prim_nth_root(1 << k) == oracle.ff_prim_nth_root(1 << k)           # ff.rs:215-223
    with pytest.raises(S.StarkPanic, match="n must be a power of two"):
        S.prim_nth_root(6)
    with pytest.raises(S.StarkPanic, match="n > 2\\^23 not supported by this modulus"):
        S.prim_nth_root(1 << 24)
    for n, ef, nq in [(32, 4, 2), (64, 4, 3), (128, 4, 4), (256, 8, 5), (1 << 22, 4, 32), (16, 4, 8), (4, 4, 1)]:
        assert S.fri_num_rounds(n, ef, nq) == oracle.fri_num_rounds(n, ef, nq)       # fri.rs:93-103
    with pytest.raises(S.StarkPanic, match="Domain length must be power of 2"):
        S.fri_num_rounds(48, 4, 2)
    with pytest.raises(S.StarkPanic, match="Expansion factor must be power of 2"):
        S.fri_num_rounds(64, 6, 2)
    with pytest.raises(S.StarkPanic, match="Expansion factor must be at least 4"):
        S.fri_num_rounds(64, 2, 2)


def test_proof_size_matches_oracle(S, oracle):
    for n, off, ef, nq in [(32, 3, 4, 2), (64, 7, 4, 3), (128, 13, 4, 4), (256, 17, 8, 5), (1024, 3, 4, 16)]:
        w = oracle.ff_prim_nth_root(n)
        cw = oracle.splitmix64(n, n)
        assert S.fri_proof_size(n, ef, nq) == len(oracle.fri_prove(cw, w, off, ef, nq)["proof"])


def test_sample_indices_host(S, oracle):
    seed = oracle.hash_from_u64(12345)
    for size, red, num in [(16, 8, 2), (1 << 21, 256, 32), (64, 16, 16), (128, 32, 5)]:
        assert list(S.fri_sample_indices(seed, size, red, num)) == list(oracle.fri_sample_indices(seed, size, red, num))
    with pytest.raises(S.StarkPanic, match="not enough entropy"):
        S.fri_sample_indices(seed, 64, 4, 9)
    with pytest.raises(S.StarkPanic, match="cannot sample more indices"):
        S.fri_sample_indices(seed, 64, 4, 5)


def test_verify_reason_texts_are_the_references(S):
    """stark_fri_verify_reason returns the `println!` lines of fri.rs:313-505 -- the strings the oracle's verifier
    reports too (oracle/stark_oracle.c fri_verify), so GPU and oracle verdicts can be compared textually"""
    lib = ctypes.CDLL(S.lib_path)
    lib.stark_fri_verify_reason.restype = ctypes.c_char_p
    texts = [lib.stark_fri_verify_reason(ctypes.c_uint32(i)).decode() for i in range(18)]
    assert texts[0] == "" and texts[17] == "unknown"
    ref = open("/root/reference/src/fri.rs").read() if os.path.exists("/root/reference/src/fri.rs") else None
    port = open(os.path.join(ROOT, "oracle", "stark_oracle.c")).read()
    for t in texts[1:17]:
        assert '"%s"' % t in port, t
        if ref is not None:
            assert '"%s"' % t in ref, t


def test_no_cpu_fallback(S):
    """Without a usable CUDA device the product must raise, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(S.StarkError, match="no CPU fallback"):
        S.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "stark-rs_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".rs", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "liboracle" not in text and "stark_oracle" not in text, f


def test_mgpu_column_partition_host_logic():
    """stark_mgpu_prove_trace's column assignment (column 0 on every rank, columns 1.. round robin) is pure host logic:
    for every group size the ranks' sets are disjoint, cover 1 .. n_cols - 1, are balanced within one column, and a group
    of one owns everything"""
    import stark_rs_b200 as S
    for world in (1, 2, 4, 8):
        for n_cols in (1, 2, 5, 16, 64):
            sets = [S.mgpu_columns_of_rank(r, world, n_cols) for r in range(world)]
            flat = sorted(c for s in sets for c in s)
            assert flat == list(range(1, n_cols)), (world, n_cols)
            assert max(len(s) for s in sets) - min(len(s) for s in sets) <= 1
            assert all(s == sorted(s) for s in sets)
    assert S.mgpu_columns_of_rank(0, 1, 16) == list(range(1, 16))
    assert S.mgpu_columns_of_rank(3, 2, 16) == [] and S.mgpu_columns_of_rank(-1, 2, 16) == []
