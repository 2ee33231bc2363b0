#!/usr/bin/env python3
"""Extract the known-answer vectors the reference's own gtests hold for the hot path into JSON fixtures.

Run in the build container (where /root/reference is mounted):  python tests/golden/make_golden.py
The GPU box has no /root/reference; tests read only the committed JSON.

Sources:
  /root/reference/tests/FirTests.cpp:8-94     (T=2, D=2, two commits of 3+2 samples -> 2 outputs)
  /root/reference/tests/FirTests.cpp:96-221   (T=3, D=2, 8 samples, output room 1 then 2)
  /root/reference/tests/CosineSourceTests.cpp:8-56 (fs=100, f=1, 101 complex samples, 1e-4)
"""
import json
import os
import re

REF = "/root/reference/tests"
HERE = os.path.dirname(os.path.abspath(__file__))
CPLX = re.compile(r"make_cuComplex\(\s*([-0-9.]+)f?\s*,\s*([-0-9.]+)f?\s*\)")


def cplx_list(block: str):
    return [[float(a), float(b)] for a, b in CPLX.findall(block)]


def braces_after(src: str, marker: str) -> str:
    i = src.index(marker)
    j = src.index("{", i)
    k = src.index("};", j)
    return src[j:k]


def taps_of(body: str):
    return [float(v) for v in re.findall(r"taps\[\d+\]\s*=\s*([-0-9.]+)f;", body)]


def main():
    src = open(os.path.join(REF, "FirTests.cpp")).read()
    t1, t2 = src.split("TEST(")[1:3]
    fir = {
        "source": "kernrj/cuda-sdr tests/FirTests.cpp",
        "tolerance": 1e-3,
        "cases": [
            {
                "name": "WhenThereAre4InputsWithDecimation2AndTwoInputCommits.ItProduces2CorrectOutputs",
                "cite": "tests/FirTests.cpp:8-94",
                "tap_type": "float", "element_type": "float_complex",
                "taps": taps_of(t1),
                "decimation": int(re.search(r"decimation = (\d+);", t1).group(1)),
                "commits": [cplx_list(braces_after(t1, "cpuData1[]")), cplx_list(braces_after(t1, "cpuData2[]"))],
                "request_bytes": [4 * 8, 2 * 8],
                "reads": [{"room_elements": None, "expected": cplx_list(braces_after(t1, "expectedValues["))}],
            },
            {
                "name": "WhenTheFirstReadCantFitAllAvailableInputs.ItDoesntSkipAnyInputValues",
                "cite": "tests/FirTests.cpp:96-221",
                "tap_type": "float", "element_type": "float_complex",
                "taps": taps_of(t2),
                "decimation": int(re.search(r"decimation = (\d+);", t2).group(1)),
                "commits": [cplx_list(braces_after(t2, "cpuData[]"))],
                "request_bytes": [8 * 8],
                "reads": [
                    {"room_elements": 1, "expected": cplx_list(braces_after(t2, "expectedValues1["))},
                    {"room_elements": 2, "expected": cplx_list(braces_after(t2, "expectedValues2["))},
                ],
            },
        ],
    }
    with open(os.path.join(HERE, "fir_kat.json"), "w") as f:
        json.dump(fir, f, indent=1)

    cs = open(os.path.join(REF, "CosineSourceTests.cpp")).read()
    fs = int(re.search(r"sampleRate = (\d+);", cs).group(1))
    freq = float(re.search(r"frequency = ([0-9.]+)f;", cs).group(1))
    tol = float(re.search(r"maxError = ([0-9.]+)f;", cs).group(1))
    cos = {
        "source": "kernrj/cuda-sdr tests/CosineSourceTests.cpp",
        "cite": "tests/CosineSourceTests.cpp:8-56",
        "sample_type": "float_complex",
        "sample_rate": fs,
        "frequency": freq,
        "output_value_count": fs + 1,
        "tolerance": tol,
        "expected_formula": "theta = float(i) * frequency / sampleRate * pi_f * 2 (float32); (cos(theta), sin(theta))",
        "note": "the output buffer is allocated as 101*8 = 808 B and rounded up by the 32-byte-aligning allocator "
                "(CudaAllocator.cpp:54-55) to 832 B = 104 samples, so the source emits 104 samples with "
                "phiEnd = 104*delta; the first 101 are checked",
        "allocator_alignment": 32,
    }
    with open(os.path.join(HERE, "cosine_kat.json"), "w") as f:
        json.dump(cos, f, indent=1)
    print("wrote fir_kat.json, cosine_kat.json")


if __name__ == "__main__":
    main()
