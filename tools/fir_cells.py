#!/usr/bin/env python3
"""Run chosen cells of the FIR sweep once each through the gsdr C-ABI (for ncu):  python tools/fir_cells.py 4096x1 1024x64 32x64"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from cuda_sdr_b200 import ops  # noqa: E402

taps = bench.load_taps_module()
dev = torch.device("cuda", 0)
n = 1 << 26
x = torch.view_as_complex(torch.randn(n, 2, device=dev, dtype=torch.float32))
for cell in sys.argv[1:]:
    T, D = (int(v) for v in cell.split("x"))
    h = torch.from_numpy(taps.lowpass(T, 0.2, 1.0)).to(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.fir("fc", h, x, D, ops.fir_num_outputs(n, T, D))
    e1.record()
    torch.cuda.synchronize()
    print(f"T={T} D={D}: {e0.elapsed_time(e1):.3f} ms", flush=True)
