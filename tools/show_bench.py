#!/usr/bin/env python3
"""Print the numbers of one or more bench.py JSON lines that matter when reading a multi-GPU run."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.load(open(path))
    except Exception as e:
        print(path, "unreadable:", e)
        try:
            print(open(path.replace(".json", ".err")).read()[-1500:])
        except Exception:
            pass
        continue
    print(f"{path}: n={d['n_gpus']} value={d['value']:.0f} ms/step={d['ms_per_step']:.5f} comm_exposed_ms={d.get('comm_exposed_ms')}")
    for p in d.get("per_rank", []):
        c = p.get("clocks", {})
        print(f"   rank {p['rank']}: {p['ms_per_step']:.5f} kernels {p['kernel_ms_per_step']:.5f} exposed {p['comm_exposed_ms']:.4f} sm {c.get('sm_mhz')} {c.get('reasons')} {c.get('power_w_max')}")
    if d.get("gather"):
        print("   gather:", d["gather"])
    if d.get("balance"):
        print("   shares:", [round(v, 4) for v in d["balance"]["segment_samples_over_average"]])
    c = d.get("channelizer")
    if c:
        print(f"   C5: value={c['value']:.0f} ms/step={c['ms_per_step']:.4f} speedup={c.get('speedup_vs_single_gpu_same_run')} single={c.get('single_gpu_same_run')} "
              f"exposed={c.get('comm_exposed_ms')} gather={c.get('gather')}")
        for p in c.get("per_rank", []):
            print(f"      rank {p['rank']}: {p['ms_per_step']:.4f} kernels {p['kernel_ms_per_step']:.4f} exposed {p['comm_exposed_ms']:.4f}")
    if d.get("e2e"):
        print("   e2e:", d["e2e"].get("value"), d["e2e"].get("per_rank_msps"), (d["e2e"].get("c_abi") or {}).get("value"))
