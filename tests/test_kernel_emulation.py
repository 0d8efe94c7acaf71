"""CPU emulation of the device index math (the round functions of ntt_core.cuh / ntt_pass.cuh are host+device
inline): every thread of every CTA is run sequentially, phase by phase, and the result is compared with a plain
O(n log n) NTT done with u64 %.  Catches indexing, range (lazy [0,2p) invariants) and shared-memory bank-conflict
regressions without a GPU.  The parity tests proper (-m gpu) run the real kernels."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(src, arg, tmp_path):
    exe = str(tmp_path / "emul")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "stark-rs_b200", "csrc"),
                           os.path.join(ROOT, "tests", "emul", src), "-o", exe])
    out = subprocess.run([exe, str(arg)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    return out.stdout


def test_single_pass_ntt_emulation(tmp_path):
    assert "OK" in _run("ntt_emul.cpp", 12, tmp_path)          # N <= 2^12: one CTA per transform


def test_multi_pass_ntt_emulation(tmp_path):
    out = _run("ntt2_emul.cpp", 22, tmp_path)                  # 2^13..2^22: every pass plan but 2^23's, all scale modes
    assert "bank-conflicted quarter-warp accesses: 0" in out
    assert "smem range violations (>= 2p): 0" in out
    assert "OK" in out


def test_hash_forms_on_host_vs_oracle(tmp_path, oracle):
    """hash.cuh is host+device inline: the one-hash (hs) and two-hashes-per-thread (hs2, lane 1 in bits 24-31) forms
    are compared with the oracle on the CPU for random and all-ones inputs; so are the
    8-lanes-per-hash form (hso) and the 4-lanes-per-hash form (hsq), their lanes running as eight / four host threads with
    shuffles emulated through a shared slot array."""
    exe = str(tmp_path / "hash_emul")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "stark-rs_b200", "csrc"),
                           os.path.join(ROOT, "tests", "emul", "hash_emul.cpp"), os.path.join(ROOT, "oracle", "liboracle.so"),
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "hs2 OK" in out.stdout and "hso OK" in out.stdout and "hsq OK" in out.stdout and "\nOK" in out.stdout, out.stdout + out.stderr
