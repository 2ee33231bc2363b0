#!/usr/bin/env python3
"""Turn an `ncu --set full` report (+ optionally the launch list CSV) into the text summary kept under profiles/.

  python tools/ncu_summary.py gpurun_out/r1a_prof.ncu-rep [gpurun_out/r1a_launches.csv] > profiles/r1a_summary.md

Runs here (no GPU needed): it only reads the report with `ncu -i`.
"""
from __future__ import annotations

import csv
import io
import subprocess
import sys
from collections import defaultdict

METRICS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def ncu_csv(rep: str, page: str) -> list[list[str]]:
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main() -> int:
    rep = sys.argv[1]
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary of `{rep}`\n")
    print("Per-launch values (cold-cache, serialised, ~40 replays per launch: compare shares, not absolutes).\n")
    for r in rows[2:]:
        print(f"## {r[col['Kernel Name']]}  (launch id {r[col['ID']]})\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for m in METRICS:
            if m in col and r[col[m]] != "":
                print(f"| `{m}` | {r[col[m]]} | {units[col[m]]} |")
        if "dram__bytes_read.sum" in col:
            def to_bytes(name):
                v, u = float(r[col[name]]), units[col[name]].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            traffic = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
            dur = float(r[col["gpu__time_duration.sum"]])
            dur_s = dur * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(units[col["gpu__time_duration.sum"]], 1e-9)
            print(f"\nDRAM traffic per launch = {traffic / 1e6:.2f} MB -> {traffic / dur_s / 1e9:.0f} GB/s under ncu.\n")

    # SASS page: instruction mix by opcode and the hottest instructions by warp-stall samples
    try:
        src = ncu_csv(rep, "source")
    except subprocess.CalledProcessError:
        src = []
    kernels, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    for k in kernels:
        c = {name: i for i, name in enumerate(k["hdr"])}
        if "Instructions Executed" not in c or "Source" not in c:
            continue
        mix, stalls = defaultdict(float), []
        for r in k["rows"]:
            try:
                n = float(r[c["Instructions Executed"]])
                smp = float(r[c["# Samples"]])
            except (ValueError, IndexError):
                continue
            text = r[c["Source"]].strip()
            op = text.split()[1] if text.startswith("@") and len(text.split()) > 1 else text.split()[0] if text else "?"
            mix[op.split(".")[0]] += n
            stalls.append((smp, text))
        total = sum(mix.values()) or 1.0
        print(f"## SASS instruction mix of `{k['name'][:90]}` (warp-instructions executed)\n")
        print("| opcode | warp-instructions | share |\n|---|---:|---:|")
        for op, n in sorted(mix.items(), key=lambda kv: -kv[1])[:16]:
            print(f"| {op} | {n:.0f} | {100 * n / total:.1f}% |")
        tot_s = sum(s for s, _ in stalls) or 1.0
        print("\nHottest instructions by warp-stall samples:\n\n| share | SASS |\n|---:|---|")
        for smp, text in sorted(stalls, key=lambda kv: -kv[0])[:10]:
            print(f"| {100 * smp / tot_s:.1f}% | `{text[:110]}` |")
        print()

    if len(sys.argv) > 2:
        print(f"\n## Launch list `{sys.argv[2]}` (ncu --metrics gpu__time_duration.sum)\n")
        per = defaultdict(list)
        for r in csv.DictReader(l for l in open(sys.argv[2]) if l.startswith('"')):
            if r.get("Metric Name") == "gpu__time_duration.sum":
                per[r["Kernel Name"]].append(float(r["Metric Value"]))
        total = sum(sum(v) for v in per.values()) or 1.0
        print("| kernel | launches | mean ns | share of listed GPU time |\n|---|---:|---:|---:|")
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            print(f"| `{k}` | {len(v)} | {sum(v) / len(v):.0f} | {100 * sum(v) / total:.1f}% |")
    return 0


if __name__ == "__main__":
    sys.exit(main())
