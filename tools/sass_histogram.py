#!/usr/bin/env python3
"""SASS opcode histogram of every kernel in cuda_sdr_b200/libb200sdr.so (cuobjdump -sass; runs without a GPU):

  python tools/sass_histogram.py > profiles/r2_sass_histogram.md

Per kernel: static instruction count and the counts of the opcodes that tell which hardware path it uses -- UTCIMMA (tcgen05.mma
kind::i8), LDTM (tcgen05.ld from TMEM), UTMALDG (TMA tensor copy), UBLKCP (TMA bulk copy), SYNCS (mbarrier), IMMA (legacy
mma.sync s8), LDSM (ldmatrix), DFMA (fp64), FFMA / FFMA2 (fp32 scalar / packed), MUFU."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cuda_sdr_b200", "libb200sdr.so")
KEYS = ["UTCIMMA", "LDTM", "UTCBAR", "UTMALDG", "UBLKCP", "SYNCS", "IMMA", "LDSM", "DFMA", "DADD", "DMUL", "FFMA2", "FFMA", "MUFU", "PRMT", "LDS", "STS",
        "LDG", "STG", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = re.sub(r"\(.*\)$", "", name).replace("b200sdr::", "").replace("(anonymous namespace)::", "")
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and name:
            op, mods = m.group(1), m.group(2)
            if op == "FFMA2" or (op == "FFMA" and ".F32x2" in mods):
                op = "FFMA2"
            kernels[name][op] += 1
            kernels[name]["__total__"] += 1
    print("# SASS opcode histogram of `cuda_sdr_b200/libb200sdr.so` (sm_100a), static counts per kernel\n")
    print("Made by `tools/sass_histogram.py` (cuobjdump -sass).  `UTCIMMA` = tcgen05.mma kind::i8, `LDTM` = tcgen05.ld (TMEM -> registers), "
          "`UTMALDG` = TMA tensor copy, `UBLKCP` = TMA bulk copy, `SYNCS` = mbarrier, `IMMA` = legacy mma.sync s8, `LDSM` = ldmatrix.\n")
    print("| kernel | instr | " + " | ".join(KEYS) + " |")
    print("|---|---:|" + "---:|" * len(KEYS))
    for k, c in kernels.items():
        if c["__total__"] < 40:
            continue
        print(f"| `{k[:90]}` | {c['__total__']} | " + " | ".join(str(c[key]) if c[key] else "" for key in KEYS) + " |")


if __name__ == "__main__":
    sys.exit(main())
