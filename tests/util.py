"""Shared helpers of the parity tests."""
import numpy as np

REL_TOL = 1e-5  # north_star: FIR / mixer / demod within 1e-5 relative (float32) of the fp64 golden


def rel_err(got, ref) -> float:
    """||got - ref||_inf / ||ref||_inf over one block (the gate SURVEY.md section 8(d) names)."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    scale = float(np.max(np.abs(ref)))
    return float(np.max(np.abs(got.astype(ref.dtype) - ref))) / (scale if scale > 0 else 1.0)


def assert_close(got, ref, tol=REL_TOL, what=""):
    e = rel_err(got, ref)
    assert e <= tol, f"{what}: relative error {e:.3e} > {tol:.1e}"


def assert_fm_close(got, ref, gain, tol=REL_TOL, what=""):
    """FM discriminator outputs compared modulo 2*pi*gain (a sample sitting on the +-pi branch cut may
    legitimately come out on the other side in float32)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if ref.size == 0:
        return
    period = 2.0 * np.pi * abs(float(gain))
    d = got - ref
    d -= period * np.rint(d / period)
    scale = max(float(np.max(np.abs(ref))), 1e-30)
    e = float(np.max(np.abs(d))) / scale
    assert e <= tol, f"{what}: relative error {e:.3e} > {tol:.1e}"
