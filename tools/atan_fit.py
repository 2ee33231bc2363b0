#!/usr/bin/env python3
"""Coefficients of p2Atan2 (cuda_sdr_b200/csrc/pfb256_kernels.cuh): atan(z) ~ z * P(z^2) on [0, 1], P of degree 7, fitted by
Lawson-reweighted least squares on Chebyshev nodes; prints the float32 coefficients and the max error of the float32 Horner form."""
import numpy as np

z = np.cos(np.linspace(0, np.pi, 20001)) * 0.5 + 0.5
z = z[z > 1e-9]
t, f, n = z * z, np.arctan(z) / z, 7
w = np.ones_like(t)
for _ in range(60):
    c = np.linalg.lstsq(np.vander(t, n + 1, increasing=True) * w[:, None], f * w, rcond=None)[0]
    err = np.abs(np.polyval(c[::-1], t) * z - np.arctan(z))
    w = w * (1 + err / err.max())
    w /= w.max()
zz = np.linspace(0, 1, 200001)
c32, z32 = c.astype(np.float32), zz.astype(np.float32)
p = np.full_like(z32, c32[-1])
for k in range(n - 1, -1, -1):
    p = p * (z32 * z32) + c32[k]
print([float(x) for x in c32], "max error (float32 Horner):", np.abs((p * z32).astype(np.float64) - np.arctan(zz)).max())
