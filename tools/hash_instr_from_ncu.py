#!/usr/bin/env python3
"""Per-hash dynamic instruction counts of the shipped hash kernels, from an ncu capture of bench.py, for bench.py's
integer-pipe roofline (profiles/r2_hash_instr.json).

usage: tools/hash_instr_from_ncu.py gpurun_out/prof.ncu-rep profiles/r2_hash_instr.json [stark-rs_b200/build]
The capture must carry smsp__thread_inst_executed.sum and sm__inst_executed_pipe_alu.sum (both in `--set full`, or add
them with --metrics).  Per kernel the LARGEST launch is used (the one the roofline quotes).  The static SASS instruction
count of the same kernels in the built objects is recorded beside the dynamic counts: tests/test_sass_budget.py fails when
the shipped build no longer matches it, i.e. when this file is stale."""
import csv
import json
import re
import subprocess
import sys

# kernel name fragment -> (profile tag used by bench.py, hashes per thread, algorithmic bytes per hash)
KERNELS = {"k_merkle_level": ("merkle_level", 2, 96.0), "k_leaf_hash1": ("leaf_hash", 2, 36.0),
           "k_fold_leaf1": ("fold_leaf", 2, 44.0), "k_mg_fold_leaf": ("mg_fold_leaf", 2, 44.0)}


def static_counts(build_dir):
    out = {}
    for obj in ("merkle.o", "fri.o"):
        txt = subprocess.run(["cuobjdump", "-sass", "%s/%s" % (build_dir, obj)], capture_output=True, text=True).stdout
        fn = None
        for line in txt.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                fn = next((k for k in KERNELS if k in m.group(1)), None)
                if fn:
                    out[fn] = 0
                continue
            if fn and re.match(r"\s+/\*[0-9a-f]{4}\*/\s+\S", line):
                out[fn] += 1
    return out


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    build = sys.argv[3] if len(sys.argv) > 3 else "stark-rs_b200/build"
    if rep.endswith(".csv"):     # `ncu -i x.ncu-rep --page raw --csv` made on the GPU box (the report itself may exceed the copy-back limit)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "s": 1e9, "ms": 1e6, "us": 1e3, "ns": 1.0}

    def scaled(r, name):     # bytes, or nanoseconds
        v = num(r, name)
        return None if v is None else v * SCALE.get(units[col[name]], 1.0)

    def num(r, name):
        return float(r[col[name]].replace(",", "")) if name in col and r[col[name]] not in ("", "n/a") else None

    best = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        k = next((k for k in KERNELS if k in name), None)
        if not k:
            continue
        threads = num(r, "launch__grid_size") * num(r, "launch__block_size")
        if k not in best or threads > best[k][0]:
            best[k] = (threads, r)
    stat = static_counts(build)
    out = {"source": rep, "note": "dynamic counts per hash from the largest launch of each kernel; alu = warp-level "
                                  "sm__inst_executed_pipe_alu.sum x 32 / hashes", "kernels": {}}
    for k, (threads, r) in best.items():
        tag, per_thread, bytes_per_hash = KERNELS[k]
        hashes = threads * per_thread
        ti, alu = num(r, "smsp__thread_inst_executed.sum"), num(r, "sm__inst_executed_pipe_alu.sum")
        if ti is None and num(r, "smsp__inst_executed.sum"):
            ti = 32 * num(r, "smsp__inst_executed.sum")      # warp-level count x 32 lanes (these kernels run full warps)
        fma = num(r, "sm__inst_executed_pipe_fma.sum")
        rd, wr = scaled(r, "dram__bytes_read.sum"), scaled(r, "dram__bytes_write.sum")
        out["kernels"][tag] = {
            "kernel": k, "hashes_in_launch": hashes, "thread_instr_per_hash": ti / hashes if ti else None,
            "alu_instr_per_hash": alu * 32 / hashes if alu else None, "fma_instr_per_hash": fma * 32 / hashes if fma else None,
            "algorithmic_bytes_per_hash": bytes_per_hash,
            "alu_pipe_pct_ncu": num(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "issue_active_pct_ncu": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "dram_bytes_per_launch": (rd or 0) + (wr or 0), "duration_ns": scaled(r, "gpu__time_duration.sum"),
            "static_sass_instr": stat.get(k)}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
