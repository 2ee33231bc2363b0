#!/usr/bin/env python3
"""Print the frac table of a `bench.py --workload firsweep` JSON line:  python tools/show_sweep.py gpurun_out/x_firsweep.json [--md]"""
import json
import statistics
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
c = d["cells"]
md = "--md" in sys.argv
Ts = sorted({x["taps"] for x in c})
Ds = sorted({x["decimation"] for x in c})
cell = {(x["taps"], x["decimation"]): x for x in c}
if md:
    print("| T \\\\ D | " + " | ".join(str(D) for D in Ds) + " |")
    print("|---:|" + "---:|" * len(Ds))
    for T in Ts:
        print(f"| {T} | " + " | ".join(f"{cell[T, D]['frac']:.2f} {cell[T, D]['bound'][0]}" for D in Ds) + " |")
else:
    print("T\\D ", *[f"{D:>6}" for D in Ds])
    for T in Ts:
        print(f"{T:>4}", *[f"{cell[T, D]['frac']:6.2f}{cell[T, D]['bound'][0]}" for D in Ds])
fr = [x["frac"] for x in c]
print(f"\nmin {min(fr):.3f}  median {statistics.median(fr):.3f}  max {max(fr):.3f}  cells >= 0.70: {sum(f >= 0.7 for f in fr)} of {len(fr)}; >= 0.50: {sum(f >= 0.5 for f in fr)}")
