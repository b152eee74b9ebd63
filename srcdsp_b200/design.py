"""Integer tap design for the benchmark workloads and the tools (host side, numpy only).

The reference ships no filter designer: its users hand `setCoeffs` / `setCoefficients` integer taps scaled to
`coeffScaling = 15` (dsptl_dnsampling_filters.h:215, upsampling_filters.h:204-208).  These two helpers produce such
taps for BASELINE.json's shapes.  `oracle.design_*` are the test suite's twins (`tests/test_design.py` keeps the two
equal), so that the benchmark's device arm does not reach into `oracle/`.
"""
from __future__ import annotations

import numpy as np

__all__ = ["design_lowpass_taps", "design_interp_taps"]


def _windowed_sinc(ntaps: int, ratio: int) -> np.ndarray:
    t = np.arange(ntaps, dtype=np.float64) - (ntaps - 1) / 2.0
    return np.sinc(t / ratio) * np.hamming(ntaps)


def design_lowpass_taps(ntaps: int, ratio: int, target_sum: int = 49152, pad_to: int | None = None) -> np.ndarray:
    """Hamming-windowed sinc with cut-off at the decimated Nyquist, as int32 with 32768 <= sum|c| <= 65535: the
    accumulator `sum c_k x_k` cannot overflow int32 for any int16 input, and `>> 15` leaves about unit gain.
    `pad_to` appends zero taps (same outputs; the non-obsolete header wants a multiple of M, :122)."""
    h = _windowed_sinc(ntaps, ratio)
    c = np.round(h * (target_sum / np.abs(h).sum())).astype(np.int32)
    total = int(np.abs(c).sum())
    if not 32768 <= total <= 65535:
        raise ValueError(f"sum|c| = {total} is outside [32768, 65535] for {ntaps} taps at /{ratio}")
    if pad_to is not None and pad_to > ntaps:
        c = np.concatenate([c, np.zeros(pad_to - ntaps, np.int32)])
    return c


def design_interp_taps(ntaps: int, L: int) -> np.ndarray:
    """Interpolation prototype with gain L in Q(15 - log2 L): every polyphase branch sums to about 32768 / L, i.e.
    unit pass-band gain after the upsampler's `>> (15 - log2 L)` (upsampling_filters.h:204-208)."""
    h = _windowed_sinc(ntaps, L)
    h *= L / h.sum()
    return np.round(h * (32768.0 / L)).astype(np.int32)
