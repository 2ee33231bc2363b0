"""Tap design used by the harness.  The reference designs taps with the un-vendored `remez` library
(src/filters/factories/RfToPcmAudioFactory.cpp:49-122); tap VALUES do not change throughput and parity
uses identical taps on both sides, so the harness uses a Hamming-windowed sinc (fp64 -> fp32, unity DC
gain), with the reference's tap-count heuristic available for sizing."""
from __future__ import annotations

import math

import numpy as np


def lowpass(num_taps: int, cutoff_hz: float, sample_rate_hz: float) -> np.ndarray:
    n = np.arange(num_taps, dtype=np.float64) - (num_taps - 1) / 2.0
    fc = cutoff_hz / sample_rate_hz
    h = 2.0 * fc * np.sinc(2.0 * fc * n)
    h *= 0.54 - 0.46 * np.cos(2.0 * np.pi * np.arange(num_taps) / max(1, num_taps - 1))
    h /= h.sum()
    return h.astype(np.float32)


def fred_harris_tap_count(db_attenuation: float, transition_hz: float, sample_rate_hz: float) -> int:
    """RfToPcmAudioFactory.cpp:44-47 (same float expression)."""
    return int(round(math.ceil(-db_attenuation / (22.0 * (transition_hz / sample_rate_hz)))))
