"""Parity oracle for the SrcDsp DDC hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (srcdsp_b200) never does.

Three checkers, strongest last:
  * numpy restatement (np_* functions below)            -- closed forms of SURVEY.md Appendix A
  * liborc.so  (oracle/srcdsp_oracle.c, plain C)        -- `corc()`
  * _ref/libsrcdsp_ref.so (the unmodified reference headers compiled from /root/reference by
    oracle/Makefile)                                    -- `ref()`, None when not built

Parity status: PINNED against the compiled reference (tests/test_oracle.py) and against the
fixtures under tests/golden/ that were generated from it (tests/golden/make_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_i16p = C.POINTER(C.c_int16)
_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)


def build(quiet: bool = True) -> None:
    """Compile liborc.so and, when the reference checkout is present, _ref/."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _p16(a: np.ndarray):
    assert a.dtype == np.int16 and a.flags.c_contiguous
    return a.ctypes.data_as(_i16p)


def _p32(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_i32p)


def _pf(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_float))


def as_taps(taps) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(taps, dtype=np.int32))


# ----------------------------------------------------------------------------------------------
# plain-C restatement
# ----------------------------------------------------------------------------------------------
class _COracle:
    def __init__(self, path: str):
        L = self.lib = C.CDLL(path)
        L.orc_hash32.restype = C.c_uint32
        L.orc_hash32.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
        L.orc_synth_fill.argtypes = [_i16p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_size_t, C.c_int]
        L.orc_mixer_table.argtypes = [_i16p, C.c_uint]
        L.orc_mixer_set_frequency.restype = C.c_int
        L.orc_mixer_set_frequency.argtypes = [C.c_float, C.c_uint]
        L.orc_mixer_adjust_nominal.restype = C.c_float
        L.orc_mixer_adjust_nominal.argtypes = [C.c_float, C.c_float]
        L.orc_mixer_step.argtypes = [_i16p, C.c_uint, C.POINTER(C.c_int), C.c_int, _i16p, _i16p, C.c_size_t]
        L.orc_dec_coeff_scaling.restype = C.c_int
        L.orc_dec_coeff_scaling.argtypes = [_i32p, C.c_int]
        L.orc_dec_step.argtypes = [_i32p, C.c_int, C.c_int, C.c_uint, _i16p, _i16p, C.c_size_t, _i16p]
        L.orc_fir_step.argtypes = [_i32p, C.c_int, C.c_uint, _i16p, _i16p, C.c_size_t, _i16p]
        fp = C.POINTER(C.c_float)
        L.orc_decf_coeff_scaling.restype = C.c_uint
        L.orc_decf_coeff_scaling.argtypes = [fp, C.c_int]
        L.orc_decf_step.argtypes = [fp, C.c_int, C.c_int, C.c_uint, fp, fp, C.c_size_t, fp]
        L.orc_up_left_shift_factor.restype = C.c_int
        L.orc_up_left_shift_factor.argtypes = [C.c_int]
        L.orc_up_length.restype = C.c_int
        L.orc_up_length.argtypes = [_i32p, C.c_int]
        L.orc_up_step.argtypes = [_i32p, C.c_int, C.c_int, C.c_uint, _i16p, _i16p, C.c_size_t, C.c_size_t, _i16p]

    # -- synthetic input ----------------------------------------------------------------------
    def synth(self, seed: int, channel: int, n0: int, n: int, amp_shift: int = 0) -> np.ndarray:
        out = np.empty((n, 2), np.int16)
        self.lib.orc_synth_fill(_p16(out), seed, channel, n0, n, amp_shift)
        return out

    # -- mixer -------------------------------------------------------------------------------
    def mixer_table(self, n_table: int = 4096) -> np.ndarray:
        t = np.empty(n_table, np.int16)
        self.lib.orc_mixer_table(_p16(t), n_table)
        return t

    def mixer_set_frequency(self, f: float, n_table: int = 4096) -> int:
        return self.lib.orc_mixer_set_frequency(f, n_table)

    def mixer_adjust_nominal(self, nominal: float, adjust: float) -> float:
        return self.lib.orc_mixer_adjust_nominal(nominal, adjust)

    def mixer_step(self, x: np.ndarray, phi: int, freq: int, n_table: int = 4096):
        x = np.ascontiguousarray(x, np.int16).reshape(-1, 2)
        out = np.empty_like(x)
        p = C.c_int(phi)
        self.lib.orc_mixer_step(_p16(self.mixer_table(n_table)), n_table, C.byref(p), freq,
                                _p16(x), _p16(out), x.shape[0])
        return out, p.value

    # -- decimator ---------------------------------------------------------------------------
    def dec_coeff_scaling(self, taps) -> int:
        t = as_taps(taps)
        return self.lib.orc_dec_coeff_scaling(_p32(t), t.size)

    def dec_step(self, taps, M: int, x: np.ndarray, history: np.ndarray | None = None,
                 left_shift: int = 0):
        t = as_taps(taps)
        x = np.ascontiguousarray(x, np.int16).reshape(-1, 2)
        h = np.zeros((t.size - 1, 2), np.int16) if history is None else np.array(history, np.int16).reshape(-1, 2)
        assert h.shape[0] == t.size - 1 and x.shape[0] % M == 0
        out = np.empty((x.shape[0] // M, 2), np.int16)
        shift = self.dec_coeff_scaling(t) - left_shift
        self.lib.orc_dec_step(_p32(t), t.size, M, shift, _p16(h), _p16(x), x.shape[0], _p16(out))
        return out, h

    # -- float decimator (the reference's complex<float> instantiation) -----------------------------
    def decf_coeff_scaling(self, taps) -> int:
        t = np.ascontiguousarray(taps, np.float32)
        return self.lib.orc_decf_coeff_scaling(_pf(t), t.size)

    def decf_step(self, taps, M: int, x: np.ndarray, history: np.ndarray | None = None, left_shift: int = 0):
        t = np.ascontiguousarray(taps, np.float32)
        x = np.ascontiguousarray(x, np.float32).reshape(-1, 2)
        h = np.zeros((t.size - 1, 2), np.float32) if history is None else np.array(history, np.float32).reshape(-1, 2)
        assert h.shape[0] == t.size - 1 and x.shape[0] % M == 0
        out = np.empty((x.shape[0] // M, 2), np.float32)
        shift = (self.decf_coeff_scaling(t) - left_shift) & 0xFFFFFFFF
        self.lib.orc_decf_step(_pf(t), t.size, M, shift, _pf(h), _pf(x), x.shape[0], _pf(out))
        return out, h

    def fir_step(self, taps, x: np.ndarray, history: np.ndarray | None = None):
        t = as_taps(taps)
        x = np.ascontiguousarray(x, np.int16).reshape(-1, 2)
        h = np.zeros((t.size - 1, 2), np.int16) if history is None else np.array(history, np.int16).reshape(-1, 2)
        out = np.empty_like(x)
        self.lib.orc_fir_step(_p32(t), t.size, self.dec_coeff_scaling(t), _p16(h), _p16(x), x.shape[0], _p16(out))
        return out, h

    # -- upsampler ---------------------------------------------------------------------------
    def up_length(self, taps) -> int:
        t = as_taps(taps)
        return self.lib.orc_up_length(_p32(t), t.size)

    def up_step(self, taps, L: int, x: np.ndarray, history: np.ndarray | None = None,
                flush: bool = False, shift_mode: int = 0):
        t = as_taps(taps)
        assert t.size % L == 0
        x = np.ascontiguousarray(x, np.int16).reshape(-1, 2)
        H = t.size // L
        h = np.zeros((H - 1, 2), np.int16) if history is None else np.array(history, np.int16).reshape(-1, 2)
        n_flush = self.up_length(t) // L if flush else 0
        out = np.empty((L * (x.shape[0] + n_flush), 2), np.int16)
        shift = 15 - self.lib.orc_up_left_shift_factor(L) if shift_mode == 0 else 0
        self.lib.orc_up_step(_p32(t), t.size, L, shift, _p16(h), _p16(x), x.shape[0], n_flush, _p16(out))
        return out, h


_corc = None


class _OrcCorr(C.Structure):
    _fields_ = [("N", C.c_int), ("S", C.c_int), ("top", C.c_int), ("coeff_scaling", C.c_int),
                ("coeffs_energy", C.c_uint32), ("corr_value", C.c_uint32 * 3), ("energy_value", C.c_uint32 * 3),
                ("threshold_factor", C.c_double), ("history", C.c_void_p), ("coeffs", C.c_void_p), ("bits", C.c_void_p)]


class OrcCorrelator:
    """Sequential C restatement of FixedPatternCorrelator<int16_t, int32_t, N, S> (oracle/srcdsp_oracle.c)."""

    def __init__(self, oracle: "_COracle", N: int = 32, S: int = 4):
        self._l = oracle.lib
        self._l.orc_corr_step.argtypes = [C.POINTER(_OrcCorr), _i16p, C.c_size_t, C.POINTER(C.c_int)]
        self._l.orc_corr_set_pattern.argtypes = [C.POINTER(_OrcCorr), _i32p, C.c_double]
        self._c = _OrcCorr()
        self.N, self.S = N, S
        assert self._l.orc_corr_init(C.byref(self._c), N, S) == 0

    def __del__(self):
        if getattr(self, "_c", None) is not None and self._c.history:
            self._l.orc_corr_free(C.byref(self._c))

    def setPattern(self, pattern, thresholdCoeff=0.8):
        p = np.ascontiguousarray(pattern, np.int32).reshape(self.N, 2)
        if self._l.orc_corr_set_pattern(C.byref(self._c), _p32(p), thresholdCoeff) != 0:
            raise ValueError("pattern energy above 1073217600 (correlators.h:183)")

    def reset(self):
        self._l.orc_corr_reset(C.byref(self._c))

    def step(self, x):
        x = np.ascontiguousarray(x, np.int16).reshape(-1, 2)
        idx = C.c_int(0)
        found = self._l.orc_corr_step(C.byref(self._c), _p16(x), x.shape[0], C.byref(idx))
        return bool(found), idx.value

    def getRefBitSamples(self):
        return np.ctypeslib.as_array((C.c_int16 * (2 * self.N)).from_address(self._c.bits)).reshape(self.N, 2).copy()

    def getStatus(self):
        return dict(energyValue=list(self._c.energy_value), corrValue=list(self._c.corr_value),
                    coeffsEnergy=self._c.coeffs_energy, coeffScaling=self._c.coeff_scaling,
                    thresholdFactor=self._c.threshold_factor)


class RefCorrelator:
    """The reference's FixedPatternCorrelator<int16_t, int32_t, N, S> ((N, S) compiled into ref_harness.cpp:
    (32, 4), (16, 2), (8, 1), (64, 8))."""

    def __init__(self, lib: "_RefLib", N: int = 32, S: int = 4):
        self._l, self.N = lib.lib, N
        self._h = self._l.ref_corr_create(N, S)
        if not self._h:
            raise ValueError(f"ref_harness.cpp has no FixedPatternCorrelator<.., {N}, {S}>")

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.ref_corr_destroy(self._h)
            self._h = None

    def setPattern(self, pattern, thresholdCoeff=0.8):
        p = np.ascontiguousarray(pattern, np.int32).reshape(self.N, 2)
        self._l.ref_corr_set_pattern(self._h, _p32(p), thresholdCoeff)

    def reset(self):
        self._l.ref_corr_reset(self._h)

    def step(self, x):
        x = np.ascontiguousarray(x, np.int16).reshape(-1, 2)
        idx = C.c_int(0)
        found = self._l.ref_corr_step(self._h, _p16(x), x.shape[0], C.byref(idx))
        return bool(found), idx.value

    def getRefBitSamples(self):
        out = np.zeros((self.N, 2), np.int16)
        self._l.ref_corr_bits(self._h, _p16(out), self.N)
        return out

    def getStatus(self):
        e, c = (C.c_uint32 * 3)(), (C.c_uint32 * 3)()
        ce, cs = C.c_uint32(), C.c_int()
        self._l.ref_corr_status(self._h, e, c, C.byref(ce), C.byref(cs))
        return dict(energyValue=list(e), corrValue=list(c), coeffsEnergy=ce.value, coeffScaling=cs.value)


def corc() -> _COracle:
    global _corc
    if _corc is None:
        path = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(path):
            build()
        _corc = _COracle(path)
    return _corc


# ----------------------------------------------------------------------------------------------
# the compiled reference
# ----------------------------------------------------------------------------------------------
class _RefLib:
    def __init__(self, path: str):
        L = self.lib = C.CDLL(path)
        vp = C.c_void_p
        L.ref_mixer_create.restype = vp
        L.ref_mixer_create.argtypes = [C.c_uint]
        L.ref_mixer_destroy.argtypes = [vp]
        L.ref_mixer_set_frequency.argtypes = [vp, C.c_float]
        L.ref_mixer_reset.argtypes = [vp, C.c_float]
        L.ref_mixer_adjust_frequency.argtypes = [vp, C.c_float]
        L.ref_mixer_phi.argtypes = [vp]
        L.ref_mixer_freq.argtypes = [vp]
        L.ref_mixer_step.argtypes = [vp, _i16p, _i16p, C.c_size_t]
        L.ref_dec_create.restype = vp
        L.ref_dec_create.argtypes = [C.c_int, C.c_int, _i32p, C.c_int]
        L.ref_dec_destroy.argtypes = [vp]
        L.ref_dec_reset.argtypes = [vp]
        L.ref_dec_set_left_shift.argtypes = [vp, C.c_int]
        if hasattr(L, "ref_dec_set_coeffs"):
            L.ref_dec_set_coeffs.argtypes = [vp, C.c_int, _i32p, C.c_int]
        L.ref_dec_step.argtypes = [vp, _i16p, C.c_size_t, C.c_int, _i16p]
        fp = C.POINTER(C.c_float)
        L.ref_decf_create.restype = vp
        L.ref_decf_create.argtypes = [C.c_int, C.c_int, fp, C.c_int]
        L.ref_decf_destroy.argtypes = [vp]
        L.ref_decf_reset.argtypes = [vp]
        L.ref_decf_set_left_shift.argtypes = [vp, C.c_int]
        L.ref_decf_step.argtypes = [vp, fp, C.c_size_t, C.c_int, fp]
        L.ref_firf_create.restype = vp
        L.ref_firf_create.argtypes = [fp, C.c_int]
        L.ref_firf_destroy.argtypes = [vp]
        L.ref_firf_reset.argtypes = [vp]
        L.ref_firf_step.argtypes = [vp, fp, C.c_size_t, fp]
        L.ref_fir_create.restype = vp
        L.ref_fir_create.argtypes = [_i32p, C.c_int]
        L.ref_fir_destroy.argtypes = [vp]
        L.ref_fir_reset.argtypes = [vp]
        L.ref_fir_step.argtypes = [vp, _i16p, C.c_size_t, _i16p]
        L.ref_up_create.restype = vp
        L.ref_up_create.argtypes = [C.c_int, _i32p, C.c_int]
        L.ref_up_destroy.argtypes = [vp]
        L.ref_up_reset.argtypes = [vp]
        L.ref_up_set_coefficients.argtypes = [vp, _i32p, C.c_int]
        L.ref_up_get_length.argtypes = [vp]
        L.ref_up_get_imp_length.argtypes = [vp]
        L.ref_up_get_ratio.argtypes = [vp]
        L.ref_up_step.argtypes = [vp, _i16p, C.c_size_t, _i16p, C.c_int, C.c_int]
        L.ref_bench_bank.restype = C.c_double
        L.ref_bench_bank.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_int, _i16p, _i16p,
                                     C.c_int, _i32p, C.c_int, C.c_int, _i32p, C.c_int, _f32p]
        L.ref_build_info.restype = C.c_char_p
        # fifo, binary files, correlator
        L.ref_fifo_create.restype = vp
        L.ref_fifo_create.argtypes = [C.c_size_t, C.c_double]
        L.ref_fifo_destroy.argtypes = [vp]
        L.ref_fifo_write.argtypes = [vp, _i16p, C.c_size_t, C.c_uint, C.c_double]
        L.ref_fifo_read.argtypes = [vp, _i16p, C.c_size_t, C.POINTER(C.c_uint64)]
        L.ref_fifo_count.restype = C.c_size_t
        L.ref_fifo_count.argtypes = [vp]
        L.ref_fifo_reset.argtypes = [vp]
        L.ref_fifo_abs_time.argtypes = [vp, C.c_uint64, C.c_double, C.POINTER(C.c_uint), C.POINTER(C.c_double)]
        L.ref_fifo_set_time.argtypes = [vp, C.c_uint64, C.c_uint64]
        L.ref_fifo_state.argtypes = [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        L.ref_save_binary.argtypes = [C.c_char_p, _i16p, C.c_size_t]
        L.ref_read_binary.restype = C.c_size_t
        L.ref_read_binary.argtypes = [C.c_char_p, _i16p, C.c_size_t, C.c_size_t]
        L.ref_corr_create.restype = vp
        L.ref_corr_create.argtypes = [C.c_int, C.c_int]
        L.ref_corr_destroy.argtypes = [vp]
        L.ref_corr_set_pattern.argtypes = [vp, _i32p, C.c_double]
        L.ref_corr_reset.argtypes = [vp]
        L.ref_corr_step.argtypes = [vp, _i16p, C.c_size_t, C.POINTER(C.c_int)]
        L.ref_corr_bits.argtypes = [vp, _i16p, C.c_size_t]
        L.ref_corr_status.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int)]

    def build_info(self) -> str:
        return self.lib.ref_build_info().decode()

    def bench_bank(self, kind: int, x: np.ndarray, block_len: int, n_threads: int,
                   M1: int, taps1, M2: int = 1, taps2=None, lo_freq=None, want_out: bool = False):
        """x: [C, n, 2] int16.  Returns (seconds inside the reference's step() loop, out|None)."""
        x = np.ascontiguousarray(x, np.int16)
        Cn, n = x.shape[0], x.shape[1]
        t1 = as_taps(taps1)
        t2 = as_taps(taps2 if taps2 is not None else [1])
        lf = None if lo_freq is None else np.ascontiguousarray(lo_freq, np.float32)
        out = None
        if want_out:
            n_out = n * M1 if kind == 3 else n // (M1 * (M2 if kind == 2 else 1))
            out = np.zeros((Cn, n_out, 2), np.int16)
        secs = self.lib.ref_bench_bank(kind, Cn, n, block_len, n_threads, _p16(x),
                                       _p16(out) if out is not None else None,
                                       M1, _p32(t1), t1.size, M2, _p32(t2), t2.size,
                                       lf.ctypes.data_as(_f32p) if lf is not None else None)
        if secs < 0:
            raise RuntimeError(f"ref_bench_bank failed ({secs})")
        return secs, out


class RefMixer:
    def __init__(self, lib: _RefLib, n_table: int = 4096):
        self._l = lib.lib
        self._h = self._l.ref_mixer_create(n_table)
        assert self._h

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.ref_mixer_destroy(self._h)
            self._h = None

    def setFrequency(self, f):
        assert self._l.ref_mixer_set_frequency(self._h, f) == 0

    def reset(self, f=0.0):
        self._l.ref_mixer_reset(self._h, f)

    def adjustFrequency(self, f=0.0):
        self._l.ref_mixer_adjust_frequency(self._h, f)

    @property
    def phi(self):
        return self._l.ref_mixer_phi(self._h)

    @property
    def freq(self):
        return self._l.ref_mixer_freq(self._h)

    def step(self, x):
        x = np.ascontiguousarray(x, np.int16).reshape(-1, 2)
        out = np.empty_like(x)
        self._l.ref_mixer_step(self._h, _p16(x), _p16(out), x.shape[0])
        return out


class RefDecimator:
    """variant 0 = dnsampling_filters.h (no taps%M assert), 1 = dsptl_dnsampling_filters.h."""

    def __init__(self, lib: _RefLib, M: int, taps, variant: int = 0):
        self._l = lib.lib
        self.M = M
        t = as_taps(taps)
        self._h = self._l.ref_dec_create(variant, M, _p32(t), t.size)
        if not self._h:
            raise ValueError("reference would assert / unsupported M")

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.ref_dec_destroy(self._h)
            self._h = None

    def reset(self):
        self._l.ref_dec_reset(self._h)

    def setLeftShiftBy2(self, s):
        self._l.ref_dec_set_left_shift(self._h, s)

    def setCoeffs(self, taps):
        """setCoeffs() on the live object (dsptl_dnsampling_filters.h:114-134; variant 1 only)."""
        t = as_taps(taps)
        if self._l.ref_dec_set_coeffs(self._h, self.M, _p32(t), t.size) != 0:
            raise ValueError("reference would assert, or the obsolete header (no setCoeffs)")

    def step(self, x):
        x = np.ascontiguousarray(x, np.int16).reshape(-1, 2)
        assert x.shape[0] % self.M == 0
        out = np.empty((x.shape[0] // self.M, 2), np.int16)
        self._l.ref_dec_step(self._h, _p16(x), x.shape[0], self.M, _p16(out))
        return out


class RefDecF:
    """The reference's FilterDnsamplingFir<complex<float>, complex<float>, complex<float>, float, M>, compiled."""

    def __init__(self, lib: _RefLib, M: int, taps, obsolete: bool = False):
        self._l = lib.lib
        self.M = M
        t = np.ascontiguousarray(taps, np.float32)
        self._h = self._l.ref_decf_create(0 if obsolete else 1, M, _pf(t), t.size)
        if not self._h:
            raise ValueError("reference would assert / unsupported M")

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.ref_decf_destroy(self._h)
            self._h = None

    def reset(self):
        self._l.ref_decf_reset(self._h)

    def setLeftShiftBy2(self, s):
        self._l.ref_decf_set_left_shift(self._h, s)

    def step(self, x):
        x = np.ascontiguousarray(x, np.float32).reshape(-1, 2)
        assert x.shape[0] % self.M == 0
        out = np.empty((x.shape[0] // self.M, 2), np.float32)
        self._l.ref_decf_step(self._h, _pf(x), x.shape[0], self.M, _pf(out))
        return out


class RefFirF:
    """The reference's FilterFir<complex<float>, complex<float>, complex<float>, float>, compiled."""

    def __init__(self, lib: _RefLib, taps):
        self._l = lib.lib
        t = np.ascontiguousarray(taps, np.float32)
        self._h = self._l.ref_firf_create(_pf(t), t.size)

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.ref_firf_destroy(self._h)
            self._h = None

    def reset(self):
        self._l.ref_firf_reset(self._h)

    def step(self, x):
        x = np.ascontiguousarray(x, np.float32).reshape(-1, 2)
        out = np.empty_like(x)
        self._l.ref_firf_step(self._h, _pf(x), x.shape[0], _pf(out))
        return out


class RefFir:
    def __init__(self, lib: _RefLib, taps):
        self._l = lib.lib
        t = as_taps(taps)
        self._h = self._l.ref_fir_create(_p32(t), t.size)

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.ref_fir_destroy(self._h)
            self._h = None

    def reset(self):
        self._l.ref_fir_reset(self._h)

    def step(self, x):
        x = np.ascontiguousarray(x, np.int16).reshape(-1, 2)
        out = np.empty_like(x)
        self._l.ref_fir_step(self._h, _p16(x), x.shape[0], _p16(out))
        return out


class RefUpsampler:
    def __init__(self, lib: _RefLib, L: int, taps):
        self._l = lib.lib
        self.L = L
        t = as_taps(taps)
        self._h = self._l.ref_up_create(L, _p32(t), t.size)
        if not self._h:
            raise ValueError("reference would assert / unsupported L")

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.ref_up_destroy(self._h)
            self._h = None

    def reset(self):
        self._l.ref_up_reset(self._h)

    def setCoefficients(self, taps):
        t = as_taps(taps)
        assert self._l.ref_up_set_coefficients(self._h, _p32(t), t.size) == 0

    def getLength(self):
        return self._l.ref_up_get_length(self._h)

    def getImpLength(self):
        return self._l.ref_up_get_imp_length(self._h)

    def getUpsamplingRatio(self):
        return self._l.ref_up_get_ratio(self._h)

    def step(self, x, flush=False, shift_mode=0):
        x = np.ascontiguousarray(x, np.int16).reshape(-1, 2)
        n_out = self.L * (x.shape[0] + (self.getLength() // self.L if flush else 0))
        out = np.empty((n_out, 2), np.int16)
        self._l.ref_up_step(self._h, _p16(x), x.shape[0], _p16(out), int(flush), shift_mode)
        return out


_ref = {}


class RefFifo:
    """The reference's FifoWithTimeTrack<cs16, N> (capacities compiled into ref_harness.cpp: 16, 100, 1024, 65536)."""

    def __init__(self, lib: _RefLib, capacity: int, fs: float = 0.0):
        self._l = lib.lib
        self._h = self._l.ref_fifo_create(capacity, fs)
        if not self._h:
            raise ValueError(f"ref_harness.cpp has no FifoWithTimeTrack<cs16, {capacity}>")

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.ref_fifo_destroy(self._h)
            self._h = None

    def write(self, x, seconds=0, frac=0.0):
        x = np.ascontiguousarray(x, np.int16).reshape(-1, 2)
        self._l.ref_fifo_write(self._h, _p16(x), x.shape[0], seconds, frac)

    def read(self, n, start):
        out = np.zeros((n, 2), np.int16)
        st = C.c_uint64(start)
        err = self._l.ref_fifo_read(self._h, _p16(out), n, C.byref(st))
        return bool(err), st.value, (None if err else out)

    def count(self):
        return int(self._l.ref_fifo_count(self._h))

    def reset(self):
        self._l.ref_fifo_reset(self._h)

    def getAbsoluteTime(self, tp, frac=0.0):
        s, f = C.c_uint(), C.c_double()
        self._l.ref_fifo_abs_time(self._h, tp, frac, C.byref(s), C.byref(f))
        return s.value, f.value

    def _set_time(self, time_start, time_end):
        """Places the reference object's PRIVATE time counters (ref_harness.cpp reaches them through an explicit
        template instantiation, the reference source is untouched): the only way to drive buffers.h:179-207."""
        self._l.ref_fifo_set_time(self._h, time_start, time_end)

    def state(self):
        wp, a, b, r = C.c_size_t(), C.c_uint64(), C.c_uint64(), C.c_int()
        self._l.ref_fifo_state(self._h, C.byref(wp), C.byref(a), C.byref(b), C.byref(r))
        return wp.value, a.value, b.value, bool(r.value)


class PyFifo:
    """Plain-Python restatement of FifoWithTimeTrack's bookkeeping (buffers.h:139-217, 262-276, 284-352,
    361-377, 396-459).  TEST INFRASTRUCTURE ONLY."""
    U64 = (1 << 64) - 1

    def __init__(self, capacity: int, fs: float = 0.0):
        self.N, self.fs = capacity, fs
        self.storage = np.zeros((capacity, 2), np.int16)
        self.writePtr = self.timeStart = self.timeEnd = 0
        self.rollover = False
        self.ref = (0, 0, 0.0)

    def write(self, x, seconds=0, frac=0.0):  # buffers.h:139-217
        x = np.asarray(x, np.int16).reshape(-1, 2)
        n, N = x.shape[0], self.N
        assert n < N
        idx = (self.writePtr + np.arange(n)) % N
        self.storage[idx] = x
        self.writePtr = (self.writePtr + n) % N
        diff = self.U64 - self.timeEnd
        self.ref = ((self.timeEnd + 1) & self.U64, seconds, frac)
        if diff >= n:
            self.timeEnd += n
        else:
            self.timeEnd = n - diff
            self.rollover = True
        if not self.rollover:  # uint64 arithmetic, as the reference: after a wrap timeEnd < timeStart and the differences wrap too
            self.timeStart = (self.timeEnd - N + 1) & self.U64 if ((self.timeEnd - self.timeStart + 1) & self.U64) > N else 1
        else:
            d2 = self.U64 - self.timeStart
            self.timeStart = self.timeStart + n if d2 >= n else n - d2
            self.rollover = False

    def read(self, n, start):  # buffers.h:284-352
        if start < self.timeStart:
            start = self.timeStart
        if ((start + n - 1) & self.U64) > self.timeEnd:
            return True, start, None
        sp = ((self.writePtr + self.N - (self.timeEnd - start) - 1) & self.U64) % self.N
        return False, start, self.storage[(sp + np.arange(n)) % self.N].copy()

    def count(self):  # buffers.h:361-377
        if not self.rollover:
            return ((self.timeEnd - self.timeStart) + 1) & self.U64
        return ((self.U64 - self.timeStart) + self.timeEnd + 1) & self.U64

    def reset(self):  # buffers.h:262-276
        self.writePtr = self.timeStart = self.timeEnd = 0
        self.rollover = False

    def _set_time(self, time_start, time_end):  # test hook, see RefFifo._set_time
        self.timeStart, self.timeEnd = time_start, time_end

    def state(self):
        return self.writePtr, self.timeStart, self.timeEnd, self.rollover

    def getAbsoluteTime(self, tp, frac=0.0):  # buffers.h:396-459
        import math
        d = (tp - self.ref[0]) & self.U64
        if d >= 1 << 63:
            d -= 1 << 64
        td = d / self.fs
        ti = int(math.floor(td))
        s = (self.ref[1] + ti) & 0xFFFFFFFF
        fr = self.ref[2] + (td - math.floor(td)) + frac / self.fs
        t = int(fr)
        return (s + t) & 0xFFFFFFFF, fr - t


def ref_save_binary(lib: _RefLib, path: str, x: np.ndarray):
    x = np.ascontiguousarray(x, np.int16).reshape(-1, 2)
    lib.lib.ref_save_binary(path.encode(), _p16(x), x.shape[0])


def ref_read_binary(lib: _RefLib, path: str, cap: int, prefill: int = 0) -> np.ndarray:
    """What the reference's readBinarySamples leaves in a vector that held `prefill` elements (7, -7)."""
    out = np.zeros((cap, 2), np.int16)
    n = lib.lib.ref_read_binary(path.encode(), _p16(out), cap, prefill)
    return out[: min(n, cap)], n


def ref(opt: str = "O2"):
    """The compiled reference, or None when oracle/_ref was never built (fresh clone, no
    /root/reference).  opt='O0' selects the build with the reference makefile's flags."""
    if opt not in _ref:
        name = "libsrcdsp_ref.so" if opt == "O2" else "libsrcdsp_ref_O0.so"
        path = os.path.join(_HERE, "_ref", name)
        _ref[opt] = _RefLib(path) if os.path.exists(path) else None
    return _ref[opt]


# ----------------------------------------------------------------------------------------------
# numpy restatement -- independent of the C code above
# ----------------------------------------------------------------------------------------------
def np_hash32(seed: int, channel: int, n: np.ndarray) -> np.ndarray:
    n = np.asarray(n, np.uint64)
    with np.errstate(over="ignore"):
        x = (np.uint32(seed) ^ (np.uint32(channel) * np.uint32(0x9E3779B1))
             ^ ((n & np.uint64(0xFFFFFFFF)).astype(np.uint32) * np.uint32(0x85EBCA6B))
             ^ ((n >> np.uint64(32)).astype(np.uint32) * np.uint32(0xC2B2AE35)))
        x ^= x >> np.uint32(16)
        x *= np.uint32(0x7FEB352D)
        x ^= x >> np.uint32(15)
        x *= np.uint32(0x846CA68B)
        x ^= x >> np.uint32(16)
    return x


def np_synth(seed: int, channel: int, n0: int, n: int, amp_shift: int = 0) -> np.ndarray:
    h = np_hash32(seed, channel, np.arange(n0, n0 + n, dtype=np.uint64))
    re = (h & np.uint32(0xFFFF)).astype(np.uint16).view(np.int16) >> amp_shift
    im = (h >> np.uint32(16)).astype(np.uint16).view(np.int16) >> amp_shift
    return np.stack([re, im], axis=1).astype(np.int16)


def _cl16s(v):
    return np.clip(v, -32767, 32767).astype(np.int16)


def _cl16a(v):
    return np.clip(v, -32768, 32767).astype(np.int16)


def _wrap32(v):
    return v.astype(np.int64).astype(np.uint32).view(np.int32) if v.dtype != np.int32 else v


def np_mixer_table(n_table: int = 4096) -> np.ndarray:
    k = np.arange(n_table, dtype=np.float64)
    return np.trunc(16383 * np.sin(2 * np.pi * k / n_table)).astype(np.int16)


def np_mixer_set_frequency(f: float, n_table: int = 4096) -> int:
    f = np.float32(f)
    N = np.float32(n_table)

    def rnd(v):  # round half away from zero, float32
        return np.float32(np.sign(v) * np.floor(np.abs(v) + np.float32(0.5)))

    if f >= 0:
        fr = int(rnd(f * N / np.float32(2)))
    else:
        fr = int(rnd(N - rnd(-f * N / np.float32(2))))
        if fr == n_table:
            fr = 0
    return fr


def np_mixer_step(x: np.ndarray, phi: int, freq: int, n_table: int = 4096):
    x = np.asarray(x, np.int16).reshape(-1, 2).astype(np.int64)
    T = np_mixer_table(n_table).astype(np.int64)
    n = np.arange(x.shape[0], dtype=np.int64)
    ph = (phi + n * freq) % n_table
    c = T[(ph + n_table // 4) % n_table]
    s = T[ph]
    re = (x[:, 0] * c - x[:, 1] * s) >> 14
    im = (x[:, 1] * c + x[:, 0] * s) >> 14
    out = np.stack([_cl16s(re), _cl16s(im)], axis=1)
    return out, int((phi + x.shape[0] * freq) % n_table)


def np_dec_coeff_scaling(taps) -> int:
    return int(np.floor(np.log2(np.sum(np.abs(np.asarray(taps, np.float64))))))


def np_dec_step(taps, M: int, x: np.ndarray, history: np.ndarray | None = None, left_shift: int = 0):
    t = np.asarray(taps, np.int64)
    x = np.asarray(x, np.int16).reshape(-1, 2)
    Nt = t.size
    h = np.zeros((Nt - 1, 2), np.int16) if history is None else np.asarray(history, np.int16).reshape(-1, 2)
    xx = np.concatenate([h, x]).astype(np.int64)
    n_out = x.shape[0] // M
    idx = (Nt - 1) + np.arange(n_out)[:, None] * M - np.arange(Nt)[None, :]
    acc = np.einsum("k,ikc->ic", t, xx[idx])
    acc = acc.astype(np.uint64).astype(np.uint32).view(np.int32).astype(np.int64)  # int32 wrap
    sh = np_dec_coeff_scaling(taps) - left_shift
    out = _cl16s(acc >> sh)
    return out, xx[xx.shape[0] - (Nt - 1):].astype(np.int16)


def np_decf_coeff_scaling(taps) -> int:
    """dsptl_dnsampling_filters.h:128-132 with float taps: abs() is ::abs(int) there."""
    s = float(np.sum(np.abs(np.trunc(np.asarray(taps, np.float32)).astype(np.int64))))
    return int(np.floor(np.log2(s))) & 0xFFFFFFFF if s >= 1 else 0x80000000


def np_decf_step(taps, M: int, x: np.ndarray, history: np.ndarray | None = None, left_shift: int = 0):
    """Float instantiation of the decimator (dsptl_dnsampling_filters.h:188-219): float32 multiply and add per tap,
    in tap order, vectorised over the outputs (each numpy float32 operation is one IEEE rounding)."""
    t = np.asarray(taps, np.float32)
    x = np.asarray(x, np.float32).reshape(-1, 2)
    Nt = t.size
    h = np.zeros((Nt - 1, 2), np.float32) if history is None else np.asarray(history, np.float32).reshape(-1, 2)
    xx = np.concatenate([h, x])
    n_out = x.shape[0] // M
    base = (Nt - 1) + np.arange(n_out) * M
    acc = np.zeros((n_out, 2), np.float32)
    for k in range(Nt):
        acc = acc + t[k] * xx[base - k]
    sh = ((np_decf_coeff_scaling(taps) - left_shift) & 0xFFFFFFFF) & 31
    with np.errstate(invalid="ignore"):  # cvttss2si: 0x80000000 for NaN and every |y| >= 2^31 (x86-64 build of the reference)
        inr = np.abs(acc) < np.float32(2147483648.0)
        v = np.where(inr, np.trunc(np.where(inr, acc, 0)).astype(np.int64), -(1 << 31)) >> sh
    out = np.clip(v, -32767, 32767).astype(np.float32)
    return out, xx[xx.shape[0] - (Nt - 1):].copy()


def np_up_step(taps, L: int, x: np.ndarray, history: np.ndarray | None = None,
               flush: bool = False, shift_mode: int = 0):
    t = np.asarray(taps, np.int64)
    x = np.asarray(x, np.int16).reshape(-1, 2)
    Nt = t.size
    H = Nt // L
    h = np.zeros((H - 1, 2), np.int16) if history is None else np.asarray(history, np.int16).reshape(-1, 2)
    length = Nt
    while length > 0 and t[length - 1] == 0:
        length -= 1
    n_flush = length // L if flush else 0
    xx = np.concatenate([h, x, np.zeros((n_flush, 2), np.int16)]).astype(np.int64)
    n_tot = x.shape[0] + n_flush
    idx = (H - 1) + np.arange(n_tot)[:, None] - np.arange(H)[None, :]      # [j, i]
    tp = t.reshape(H, L)                                                    # [i, p] = c[p + i*L]
    acc = np.einsum("ip,jic->jpc", tp, xx[idx]).reshape(n_tot * L, 2)
    acc = acc.astype(np.uint64).astype(np.uint32).view(np.int32).astype(np.int64)
    sh = 15 - int(np.round(np.log2(L))) if shift_mode == 0 else 0
    out = _cl16a(acc >> sh)
    return out, xx[xx.shape[0] - (H - 1):].astype(np.int16) if H > 1 else np.zeros((0, 2), np.int16)


# ----------------------------------------------------------------------------------------------
# live objects: setCoeffs() / setCoefficients() on an object that has already filtered samples
# ----------------------------------------------------------------------------------------------
class NpDecimatorLive:
    """FilterDnsamplingFir as an object (dsptl_dnsampling_filters.h): `history` is the reference's member vector
    (oldest sample first, :218-219); setCoeffs() RESIZES it -- the first min(old, new) entries stay where they are,
    new entries are zero (:127) -- and resets leftShift (:133)."""

    def __init__(self, M: int, taps):
        self.M = M
        self.history = np.zeros((0, 2), np.int16)
        self.setCoeffs(taps)

    def setCoeffs(self, taps):
        t = as_taps(taps)
        assert t.size % self.M == 0, "dsptl_dnsampling_filters.h:122"
        h = np.zeros((t.size - 1, 2), np.int16)
        keep = min(h.shape[0], self.history.shape[0])
        h[:keep] = self.history[:keep]
        self.taps, self.history, self.left_shift = t, h, 0

    def reset(self):
        self.history[:] = 0

    def setLeftShiftBy2(self, s):
        self.left_shift = s

    def step(self, x):
        out, self.history = np_dec_step(self.taps, self.M, x, self.history, self.left_shift)
        return out


class NpUpsamplerLive:
    """FilterUpsamplingFir as an object (upsampling_filters.h): `buffer` is the reference's circular member vector,
    `top` its insertion index (:163-194).  setCoefficients() resizes the RAW buffer and leaves `top` alone (:118), so
    what the next step() sees as its history depends on where `top` stood."""

    def __init__(self, L: int, taps):
        self.L = L
        self.buffer = np.zeros((0, 2), np.int16)
        self.top = 0
        self.setCoefficients(taps)

    def setCoefficients(self, taps):
        t = as_taps(taps)
        assert t.size and t.size % self.L == 0, "upsampling_filters.h:110,113"
        H = t.size // self.L
        assert self.top < H, "the reference would write buffer[top] out of bounds"
        b = np.zeros((H, 2), np.int16)
        keep = min(H, self.buffer.shape[0])
        b[:keep] = self.buffer[:keep]
        self.taps, self.buffer = t, b
        self.length = t.size
        while self.length > 0 and t[self.length - 1] == 0:
            self.length -= 1

    def reset(self):
        self.buffer[:] = 0
        self.top = 0

    def getLength(self):
        return self.length

    def getImpLength(self):
        return self.taps.size

    def step(self, x, flush=False, shift_mode=0):
        x = np.asarray(x, np.int16).reshape(-1, 2)
        H = self.buffer.shape[0]
        # age order, oldest first, of the H - 1 samples in front of the next insertion: top+1 .. H-1, 0 .. top-1
        order = [(self.top + 1 + k) % H for k in range(H - 1)]
        out, _ = np_up_step(self.taps, self.L, x, self.buffer[order], flush, shift_mode)
        fed = np.concatenate([x, np.zeros((self.length // self.L if flush else 0, 2), np.int16)])
        for j in range(max(0, fed.shape[0] - H), fed.shape[0]):  # the last H insertions decide the buffer
            self.buffer[(self.top + j) % H] = fed[j]
        self.top = (self.top + fed.shape[0]) % H
        return out


# ----------------------------------------------------------------------------------------------
# tap design used by the tests: Hamming-windowed sinc, cutoff 1/ratio; bench.py's device arm uses
# the product's own srcdsp_b200/design.py, which tests/test_design.py keeps equal to these
# ----------------------------------------------------------------------------------------------
def design_lowpass_taps(ntaps: int, ratio: int, target_sum: int = 49152, pad_to: int | None = None) -> np.ndarray:
    """int32 taps with 32768 <= sum|c| <= 65535 (coeffScaling = 15, no int32 overflow for any
    int16 input).  pad_to appends zero taps (output-identical, SURVEY.md section 0 trap (i))."""
    n = np.arange(ntaps, dtype=np.float64) - (ntaps - 1) / 2.0
    h = np.sinc(n / ratio) * np.hamming(ntaps)
    c = np.round(h * (target_sum / np.sum(np.abs(h)))).astype(np.int32)
    s = int(np.sum(np.abs(c)))
    assert 32768 <= s <= 65535, s
    if pad_to is not None and pad_to > ntaps:
        c = np.concatenate([c, np.zeros(pad_to - ntaps, np.int32)])
    return c


def design_interp_taps(ntaps: int, L: int) -> np.ndarray:
    """Interpolation taps with prototype gain L in Q(15 - log2 L): every polyphase branch sums
    to about 32768 / L, i.e. unit pass-band gain after the reference's >> (15 - log2 L)."""
    n = np.arange(ntaps, dtype=np.float64) - (ntaps - 1) / 2.0
    h = np.sinc(n / L) * np.hamming(ntaps)
    h *= L / np.sum(h)
    return np.round(h * (32768.0 / L)).astype(np.int32)
