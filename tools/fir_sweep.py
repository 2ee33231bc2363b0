#!/usr/bin/env python3
"""BASELINE config C4: FIR roofline sweep through the gsdr C-ABI (gsdrFirFC: real taps x complex float data, decimating).

  T in {32 .. 4096} x D in {1 .. 64} on complex-float blocks of 2^26 samples (512 MiB, > L2), one B200.

Per cell: input Msps, algorithmic GB/s (8 + 8/D bytes per input sample) and TFLOP/s (4*T/D flop per input sample), and the
fraction of the binding roof = max(bytes / HBM peak, flop / FFMA peak) / time.  HBM peak from MEASURED_PEAKS.json; FFMA
peak measured by tools/microbench.cu on this pool (71 TFLOP/s at 1.965 GHz).  Writes a markdown table + JSON lines.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_sdr_b200 as sdr  # noqa: E402
from cuda_sdr_b200 import ops, taps  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2", type=int, default=26)
    ap.add_argument("--taps", type=int, nargs="*", default=[32, 64, 128, 256, 512, 1024, 2048, 4096])
    ap.add_argument("--decims", type=int, nargs="*", default=[1, 2, 4, 8, 16, 32, 64])
    ap.add_argument("--budget-ms", type=float, default=400.0, help="skip repeats once a cell has used this much GPU time")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "fir_sweep"))
    args = ap.parse_args()
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    ffma = 71.0  # TFLOP/s, tools/microbench.cu (ffma_rrr) on this pool
    n = 1 << args.log2
    dev = torch.device("cuda", 0)
    x = torch.view_as_complex(torch.randn(n, 2, device=dev, dtype=torch.float32))
    rows = []
    for T in args.taps:
        h = torch.from_numpy(taps.lowpass(T, 0.2, 1.0)).to(dev)
        for D in args.decims:
            n_out = ops.fir_num_outputs(n, T, D)
            ops.fir("fc", h, x, D, n_out)  # warm-up (and variant selection)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            times = []
            spent = 0.0
            while len(times) < 5 and (spent < args.budget_ms or len(times) < 1):
                e0.record()
                ops.fir("fc", h, x, D, n_out)
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
                spent += times[-1]
            ms = min(times)
            gbs = n * (8.0 + 8.0 / D) / (ms * 1e-3) / 1e9
            tflops = n * 4.0 * T / D / (ms * 1e-3) / 1e12
            t_roof = max(n * (8.0 + 8.0 / D) / (hbm * 1e9), n * 4.0 * T / D / (ffma * 1e12)) * 1e3
            rows.append({"taps": T, "decimation": D, "ms": ms, "msps": n / (ms * 1e-3) / 1e6, "gbs": gbs, "tflops": tflops,
                         "bound": "hbm" if gbs / hbm >= tflops / ffma else "ffma", "frac_of_binding_roof": t_roof / ms})
            print(json.dumps(rows[-1]), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out + ".jsonl", "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
    with open(args.out + ".md", "w") as f:
        f.write(f"# FIR roofline sweep (gsdrFirFC, 2^{args.log2} complex-float samples, one B200)\n\n")
        f.write(f"Cell = fraction of the binding roof (HBM {hbm:.0f} GB/s measured copy peak, FFMA {ffma:.0f} TFLOP/s measured); "
                "`h` = HBM-bound cell, `f` = FFMA-bound cell; second line = input Msps.\n\n")
        f.write("| taps \\ D | " + " | ".join(str(d) for d in args.decims) + " |\n|---|" + "---|" * len(args.decims) + "\n")
        for T in args.taps:
            cells = []
            for D in args.decims:
                r = next(r for r in rows if r["taps"] == T and r["decimation"] == D)
                cells.append(f"{r['frac_of_binding_roof']:.2f}{r['bound'][0]}<br>{r['msps']:.0f}")
            f.write(f"| {T} | " + " | ".join(cells) + " |\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
