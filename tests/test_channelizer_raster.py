"""CPU check of the route decision of the wideband channelizer (csrc/channelizer.cu: onRaster, through the C-ABI hook
b200sdr_channelizer_raster): which channel sets take the filter-bank route, on which raster, and with which bins."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def lib():
    import cuda_sdr_b200 as m
    return m._native.lib


def raster(lib, freqs, fs):
    f = np.ascontiguousarray(freqs, dtype=np.float64)
    bins = np.zeros(f.size, dtype=np.int32)
    n = lib.b200sdr_channelizer_raster(f.ctypes.data_as(C.POINTER(C.c_double)), f.size, fs, bins.ctypes.data_as(C.POINTER(C.c_int32)))
    return int(n), bins.tolist()


def test_c5_raster_is_one_over_256_with_offset(lib):
    fs = 153.6e6
    freqs = [(c - 128) * 600e3 + 100e3 for c in range(256)]  # bench.py --workload channelizer
    n, bins = raster(lib, freqs, fs)
    assert n == 256
    assert bins == [c for c in range(256)]  # bins are relative to channel 0 (the common offset goes into the taps), mod N


def test_subsets_resolve_to_the_coarsest_raster_that_holds_them(lib):
    fs = 1.024e6
    assert raster(lib, [7e3 + b * fs / 64 for b in (0, 8, 16, 40)], fs) == (8, [0, 1, 2, 5])
    assert raster(lib, [7e3 + b * fs / 16 for b in (0, 3, 5, 15, 9, 8)], fs) == (16, [0, 3, 5, 15, 9, 8])
    assert raster(lib, [5e3, 5e3, 5e3], fs) == (4, [0, 0, 0])            # one frequency: any raster
    n, bins = raster(lib, [0.0, -fs / 4, fs / 4, fs / 2], fs)             # negative frequencies wrap: turns are mod 1
    assert (n, bins) == (4, [0, 3, 1, 2])


def test_frequencies_off_any_raster_keep_the_per_channel_route(lib):
    fs = 1.024e6
    assert raster(lib, [-300e3, -111e3, 50e3, 222e3, 333e3, -7e3], fs)[0] == 0
    assert raster(lib, [(-5 + i) * 90e3 + 1e3 for i in range(11)], fs)[0] == 0     # fs * 45/512: finer than 1/256
    assert raster(lib, [0.0, fs / 256 * (1 + 1e-9)], fs)[0] == 0                    # 1e-9 of a bin off: not on the raster
    assert raster(lib, [0.0, fs / 256], fs)[0] == 256


def test_bad_arguments(lib):
    assert lib.b200sdr_channelizer_raster(None, 3, 1.0, None) == 0
    f = np.zeros(2)
    assert lib.b200sdr_channelizer_raster(f.ctypes.data_as(C.POINTER(C.c_double)), 2, 0.0, None) == 0
