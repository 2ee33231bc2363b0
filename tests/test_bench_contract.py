"""bench.py contract checks that need no GPU: the reference (CPU) arm prints one well-formed JSON line, names the same workload
as the GPU arm would, and never maps the product library (the arm is the oracle port, nothing of cuda_sdr_b200 on its path)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_and_isolation():
    code = (
        "import sys, io, contextlib; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0']\n"
        "sys.path.insert(0, %r); import bench\n"
        "bench.main()\n"
        "maps = open('/proc/self/maps').read()\n"
        "print('MAPPED_PRODUCT=' + str('libb200sdr' in maps or 'libgpusdrpipeline' in maps))\n" % ROOT)
    env = dict(os.environ, OMP_NUM_THREADS="2")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [line for line in res.stdout.splitlines() if line.strip()]
    assert lines[-1] == "MAPPED_PRODUCT=False", lines[-1]
    rec = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in rec, key
    assert rec["impl"] == "reference" and rec["cpu_baseline"]["kind"] == "port" and rec["cpu_baseline"]["value"] == rec["value"]
    assert rec["e2e"] == {"value": rec["value"], "unit": rec["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert rec["warmup"] == 0 and rec["steps"] == 1 and rec["gpu_launches"] == 0
    # the workload string is the GPU arm's (no block size in it); the block a step really ran is its own key
    sys.path.insert(0, ROOT)
    import bench
    wl = bench.workload("am")
    assert rec["config"]["workload"] == bench.config_dict(wl, 28)["workload"]
    assert rec["config"]["samples_per_gpu_per_step"] <= 1 << 26 and rec["config"]["bounded_sample_of_samples_per_gpu_per_step"] == 1 << 28
    assert f"2^{rec['config']['samples_per_gpu_per_step'].bit_length() - 1} samples" in rec["cpu_baseline"]["sample"]


def test_gather_slab_schedule():
    """bench.py's slab lengths: they cover the run exactly, never exceed the slab capacity, never grow, and end with one step."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.gather_slab_schedule(20, 5) == [5, 5, 5, 2, 1, 1, 1]
    for steps in (1, 2, 3, 7, 20, 200, 1000):
        for longest in (1, 5, 32):
            sizes = bench.gather_slab_schedule(steps, longest)
            assert sum(sizes) == steps and max(sizes) <= longest and sizes[-1] == 1
            assert all(a >= b for a, b in zip(sizes, sizes[1:]))
