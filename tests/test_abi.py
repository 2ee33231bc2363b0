"""CPU-side checks of the C-ABI: the shared library loads and exports every symbol include/*.h
declares, and the pure host arithmetic (counts, phase step, segments) matches the oracle.  No compute
call is made here (no GPU)."""
import ctypes
import os
import re
import subprocess

import pytest

from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    import __graft_entry__ as g
    g.build()
    from cuda_sdr_b200 import _native
    return _native


def _declared(header, macro):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith("#"))
    return set(re.findall(macro + r"\s+[\w\s\*]+?\b(\w+)\s*\(", src))


def test_library_exports_every_declared_symbol(native):
    declared = _declared("gsdr/gsdr.h", "GSDR_EXPORT") | _declared("gsdr/conversion.h", "GSDR_EXPORT") | \
        _declared("b200sdr/b200sdr.h", "B200SDR_EXPORT")
    assert len(declared) >= 14 + 15, declared
    exported = set(subprocess.run(["nm", "-D", "--defined-only", native.LIB_PATH], capture_output=True, text=True,
                                  check=True).stdout.split())
    missing = sorted(d for d in declared if d not in exported)
    assert not missing, missing
    # and the ctypes table binds all of them
    bound = set(native.GSDR_SYMBOLS) | set(native.B200SDR_SYMBOLS)
    assert declared <= bound, sorted(declared - bound)


def test_cubin_is_sm_100a(native):
    out = subprocess.run(["cuobjdump", "-lelf", native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out[:400]


def test_count_rules_match_oracle(native):
    f = native.lib.b200sdr_fir_num_outputs
    for T in (1, 2, 3, 63, 101, 545):
        for D in (0, 1, 2, 10, 40, 80):
            for n in list(range(0, 50)) + [1 << 20, (1 << 20) + 7]:
                assert f(n, T, D) == orc.fir_num_outputs(n, T, max(1, D)), (n, T, D)
    assert f(1 << 20, 63, 10) == 104851


def test_phase_step_matches_oracle(native):
    for f, fs in [(0, 1e6), (250e3, 1e6), (-250e3, 1e6), (1.234e6, 19.2e6), (-2.5e6, 19.2e6), (19.2e6, 19.2e6), (1e-3, 19.2e6)]:
        assert native.lib.b200sdr_phase_step(f, fs) == orc.phase_step(f, fs), (f, fs)


def test_chain_create_rejects_bad_arguments_without_gpu(native):
    cfg = native.ChainConfig()
    handle = ctypes.c_void_p()
    assert native.lib.b200sdr_chain_create(None, ctypes.byref(handle)) == 4  # Status_InvalidArgument
    cfg.struct_size = 3
    assert native.lib.b200sdr_chain_create(ctypes.byref(cfg), ctypes.byref(handle)) == 4
    cfg.struct_size = ctypes.sizeof(native.ChainConfig)
    cfg.input_type = 1  # float (real) input is not a chain input type
    assert native.lib.b200sdr_chain_create(ctypes.byref(cfg), ctypes.byref(handle)) == 4
    assert b"input_type" in native.lib.b200sdr_last_error()


def test_no_cpu_fallback(native):
    """Without a CUDA device the product refuses to run rather than falling back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from cuda_sdr_b200 import Chain, ops
    with pytest.raises(ValueError):
        ops.int8_to_norm_float(torch.zeros(16, dtype=torch.int8))
    with pytest.raises(native.NativeError):
        Chain(1e6, 1e3, np.ones(8, np.float32), 4)


def test_gather_create_rejects_bad_arguments_without_gpu(native):
    """b200sdr_gather_create validates its configuration before it touches CUDA or NCCL (include/b200sdr/b200sdr.h)."""
    cfg = native.GatherConfig()
    handle = ctypes.c_void_p()
    counts = (ctypes.c_size_t * 2)(16, 16)
    create = native.lib.b200sdr_gather_create
    assert create(None, ctypes.byref(handle)) == 4
    cfg.struct_size = 5
    assert create(ctypes.byref(cfg), ctypes.byref(handle)) == 4 and b"struct_size" in native.lib.b200sdr_last_error()
    cfg.struct_size = ctypes.sizeof(native.GatherConfig)
    cfg.rank, cfg.world, cfg.slabs, cfg.floats_per_rank = 2, 2, 3, counts   # rank out of range
    assert create(ctypes.byref(cfg), ctypes.byref(handle)) == 4
    cfg.rank, cfg.slabs = 0, 1                                              # a ring needs two slabs
    assert create(ctypes.byref(cfg), ctypes.byref(handle)) == 4
    cfg.slabs, cfg.mode = 3, 3                                              # unknown transport
    assert create(ctypes.byref(cfg), ctypes.byref(handle)) == 4 and b"mode" in native.lib.b200sdr_last_error()
    cfg.mode = 0                                                            # NCCL transport without the unique id
    assert create(ctypes.byref(cfg), ctypes.byref(handle)) == 4 and b"nccl_unique_id" in native.lib.b200sdr_last_error()
    assert not handle.value


def test_config_structs_match_the_header(native, tmp_path):
    """sizeof / offsetof of the C-ABI's configuration structs as gcc sees the header == the ctypes mirrors in _native.py."""
    probe = tmp_path / "layout.c"
    probe.write_text(
        '#include <stddef.h>\n#include <stdio.h>\n#include <b200sdr/b200sdr.h>\n'
        'int main(void) {\n'
        '  printf("%zu %zu %zu %zu %d %d %d\\n", sizeof(b200sdr_gather_config), offsetof(b200sdr_gather_config, floats_per_rank),\n'
        '         offsetof(b200sdr_gather_config, nccl_unique_id), offsetof(b200sdr_gather_config, mode),\n'
        '         (int)B200SDR_GATHER_NCCL, (int)B200SDR_GATHER_PEER, (int)B200SDR_GATHER_PEER_COPY);\n'
        '  printf("%zu\\n", sizeof(b200sdr_chain_config));\n  return 0;\n}\n')
    exe = tmp_path / "layout"
    cuda_inc = "/usr/local/cuda/include"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-I", cuda_inc, str(probe), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    G = native.GatherConfig
    assert [int(v) for v in out[:4]] == [ctypes.sizeof(G), G.floats_per_rank.offset, G.nccl_unique_id.offset, G.mode.offset]
    from cuda_sdr_b200 import sharding
    assert [int(v) for v in out[4:7]] == [sharding.Gather.NCCL, sharding.Gather.PEER, sharding.Gather.PEER_COPY]
    assert int(out[7]) == ctypes.sizeof(native.ChainConfig)
