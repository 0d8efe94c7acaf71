"""Static instruction budgets of the NTT pass kernels, read from the SASS of the built library (no GPU needed).

The passes are bound by the FMA-heavy pipe, on which IMAD.WIDE / IMAD.HI take two slots, so a
compiler or source change that brings back a spurious instruction per product, or moves the butterfly sums back to the
FMA pipe, shows up here as a budget overrun long before anybody looks at a profile.  Budgets = shipped counts + ~3 %."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "stark-rs_b200", "build", "ntt.o")

# kernel -> (max instructions, max FMA slots, max ALU-pipe instructions) per thread = per 32 elements
BUDGET = {
    "k_ntt2_pass<8, 0, 0, 12>": (1330, 780, 620),    # FIRST, radix 2^8 (four-step twiddles)
    "k_ntt2_pass<7, 1, 0, 12>": (1000, 542, 465),    # MIDDLE, radix 2^7
    "k_ntt2_pass<7, 2, 0, 12>": (950, 480, 465),     # LAST, radix 2^7, no post-scale
    "k_ntt2_pass<7, 2, 1, 12>": (1010, 540, 495),    # LAST with the constant post-scale (iNTT)
    "k_ntt2_pass<7, 2, 2, 12>": (1190, 705, 570),    # LAST with the geometric post-scale (coset interpolation, LDE)
}


@pytest.mark.skipif(not os.path.exists(OBJ), reason="library objects not built")
def test_ntt_pass_instruction_budgets():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_hist.py"), OBJ, "k_ntt2_pass"],
                         capture_output=True, text=True, check=True).stdout
    seen = {}
    name = None
    for line in out.splitlines():
        m = re.match(r"void (k_ntt2_pass<\d+, \d+, \d+, \d+>)", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"total=(\d+) fma_pipe=\d+ \(fma_slots=(\d+)\) alu_pipe=(\d+)", line)
        if m and name:
            seen[name] = tuple(int(x) for x in m.groups())
    for k, (mt, mf, ma) in BUDGET.items():
        assert k in seen, "kernel %s not found in the SASS" % k
        t, f, a = seen[k]
        assert t <= mt and f <= mf and a <= ma, "%s: %d instructions, %d FMA slots, %d ALU (budget %d / %d / %d)" % (k, t, f, a, mt, mf, ma)


def test_hash_instruction_table_matches_build():
    """bench.py takes the per-hash instruction counts of the hash kernels from profiles/r2_hash_instr.json (an ncu capture
    of this round, tools/hash_instr_from_ncu.py).  The file records the static SASS size of each kernel at capture time:
    if the shipped build has drifted from it (more than 1 %) the dynamic counts are stale and must be re-captured."""
    import json
    path = os.path.join(ROOT, "profiles", "r2_hash_instr.json")
    if not os.path.exists(path) or not os.path.exists(os.path.join(ROOT, "stark-rs_b200", "build", "merkle.o")):
        pytest.skip("no capture / no objects")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import hash_instr_from_ncu as H
    now = H.static_counts(os.path.join(ROOT, "stark-rs_b200", "build"))
    for tag, e in json.load(open(path))["kernels"].items():
        if e.get("static_sass_instr"):
            assert abs(now[e["kernel"]] - e["static_sass_instr"]) <= 0.01 * e["static_sass_instr"], (tag, now[e["kernel"]], e["static_sass_instr"])
