#!/usr/bin/env python3
"""SASS opcode histogram per kernel (static instruction counts; the NTT / hash kernels are fully unrolled, so static
counts are dynamic counts per thread).  usage: sass_hist.py file.o [substring-filter]"""
import collections
import re
import subprocess
import sys

FMA_PIPE = {"IMAD", "FFMA", "FMUL", "FADD"}                       # IMAD (incl. .MOV/.SHL/.IADD/.WIDE) issue on the FMA pipe
ALU_PIPE = {"IADD3", "IADD", "LOP3", "VIADDMNMX", "VIADD", "VIMNMX", "ISETP", "SEL", "SHF", "PRMT", "LEA", "MOV", "IABS", "FSEL", "ISCADD"}


def main():
    obj = sys.argv[1]
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    names = {}
    fn = None
    hist = collections.defaultdict(collections.Counter)
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and fn:
            op, mods = m.group(1), m.group(2)
            hist[fn][op + (".WIDE" if ".WIDE" in mods else "") + (".HI" if ".HI" in mods and op == "IMAD" else "") + (".MOV" if ".MOV" in mods else "") + (".SHL" if ".SHL" in mods else "") + (".IADD" if ".IADD" in mods else "")] += 1
    dem = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
    for f, d in zip(hist, dem):
        if flt not in d:
            continue
        h = hist[f]
        tot = sum(h.values())
        fma = sum(v for k, v in h.items() if k.split(".")[0] in FMA_PIPE)
        alu = sum(v for k, v in h.items() if k.split(".")[0] in ALU_PIPE)
        # measured on B200 (stark_bench_mul_peak): IMAD.WIDE / IMAD.HI issue at half the rate of a 32-bit IMAD, i.e. they
        # occupy the FMA-heavy pipe for two slots
        slots = fma + sum(v for k, v in h.items() if k.startswith("IMAD.WIDE") or k.startswith("IMAD.HI"))
        print("%s\n  total=%d fma_pipe=%d (fma_slots=%d) alu_pipe=%d other=%d" % (d[:150], tot, fma, slots, alu, tot - fma - alu))
        print("   " + " ".join("%s=%d" % kv for kv in sorted(h.items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    main()
