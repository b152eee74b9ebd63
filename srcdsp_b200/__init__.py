"""srcdsp_b200 -- B200-native (sm_100a) drop-in for the SrcDsp DDC hot path.

Host-side mirror of the reference's class API (same names and argument meaning):

    Mixer                 <- dsptl::Mixer<cs16, cs16, int16_t, N>            mixers.h:130-188
    FilterDnsamplingFir   <- dsptl::FilterDnsamplingFir<cs16,cs16,cs32,int32_t,M>
                                                       dsptl_dnsampling_filters.h:43-220
    FilterUpsamplingFir   <- dsptl::FilterUpsamplingFir<cs16,cs16,cs32,int32_t,L>
                                                       upsampling_filters.h:35-323
    Ddc                   <- the hand-written chain  mixer.step(); dec1.step(); dec2.step();

Every object is a BANK of `channels` independent streams (channels=1 is one reference object).
All compute happens in libsrcdsp_b200.so (hand-written CUDA behind the C ABI of
include/srcdsp_b200.h); this module only marshals buffers.  Buffers are either numpy int16
arrays [C, n, 2] / [n, 2] (host; staged through pinned-free async copies) or torch CUDA int16
tensors of the same shape (device resident, asynchronous on torch's current stream).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _capi
from ._capi import SrcDspError, check, lib
from .design import design_interp_taps, design_lowpass_taps

__all__ = ["DdcGroup", "Mixer", "FilterDnsamplingFir", "FilterDnsamplingFirFloat", "FilterFirFloat", "FilterFir", "FilterUpsamplingFir", "Ddc", "SrcDspError",
           "synth_fill", "launch_count", "device_count", "PinnedBuffer", "FifoWithTimeTrack", "saveBinarySamples",
           "readBinarySamples", "FixedPatternCorrelator", "design_lowpass_taps", "design_interp_taps"]


# ----------------------------------------------------------------------------------------------
# buffer marshalling
# ----------------------------------------------------------------------------------------------
def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class _Buf:
    """A [C, n, 2] int16 view of a numpy array or a torch CUDA tensor."""

    __slots__ = ("obj", "ptr", "C", "n", "stride", "device", "squeeze")

    def __init__(self, x, channels: int):
        self.obj = x
        shape = tuple(x.shape)
        if len(shape) == 2 and channels == 1:
            self.squeeze = True
            shape = (1,) + shape
        else:
            self.squeeze = False
        if len(shape) != 3 or shape[2] != 2 or shape[0] != channels:
            raise ValueError(f"expected int16 [{channels}, n, 2] (or [n, 2] for one channel), got {tuple(x.shape)}")
        self.C, self.n = shape[0], shape[1]
        if _is_torch(x):
            import torch
            if x.dtype != torch.int16 or not x.is_cuda:
                raise TypeError("torch buffers must be CUDA int16 tensors")
            st = x.stride()
            if st[-1] != 1 or st[-2] != 2:
                raise ValueError("samples must be contiguous interleaved I/Q")
            self.stride = (st[0] // 2) if not self.squeeze else self.n
            self.ptr = x.data_ptr()
            self.device = True
        else:
            if not isinstance(x, np.ndarray) or x.dtype != np.int16:
                raise TypeError("host buffers must be numpy int16 arrays")
            st = x.strides
            if st[-1] != 2 or st[-2] != 4:
                raise ValueError("samples must be contiguous interleaved I/Q")
            self.stride = (st[0] // 4) if not self.squeeze else self.n
            self.ptr = x.ctypes.data
            self.device = False


def _alloc_like(x, channels: int, n: int, squeeze: bool):
    shape = (n, 2) if squeeze else (channels, n, 2)
    if _is_torch(x):
        import torch
        return torch.empty(shape, dtype=torch.int16, device=x.device)
    return np.empty(shape, np.int16)


def _taps(t) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(t, dtype=np.int32).reshape(-1))


class _Handle:
    _destroy = None

    def __init__(self):
        self._h = C.c_void_p()
        self._stream = None

    def __del__(self):
        try:
            if self._h and self._destroy:
                getattr(lib(), self._destroy)(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def _bind_stream(self, buf: _Buf, setter: str):
        """Device buffers run on torch's current stream (so torch.cuda.Event timing sees it)."""
        if buf.device:
            import torch
            s = torch.cuda.current_stream().cuda_stream
            if s == 0:
                s = 1  # cudaStreamLegacy: the handle must stay ordered with torch's default stream
            if s != self._stream:
                check(getattr(lib(), setter)(self._h, C.c_void_p(s)))
                self._stream = s


# ----------------------------------------------------------------------------------------------
class Mixer(_Handle):
    """Table-lookup complex NCO mixer bank (mixers.h:130-188)."""

    _destroy = "srcdsp_mixer_destroy"

    def __init__(self, n_table: int = 4096, channels: int = 1, device: int = 0):
        super().__init__()
        self.channels, self.device, self.n_table = channels, device, n_table
        check(lib().srcdsp_mixer_create(C.byref(self._h), device, channels, n_table))

    def setFrequency(self, loFreq, ch: int = -1):
        """mixers.h:51-67.  A sequence sets one frequency per channel."""
        if np.ndim(loFreq) == 0:
            check(lib().srcdsp_mixer_set_frequency(self._h, ch, float(loFreq)))
        else:
            f = np.ascontiguousarray(loFreq, np.float32)
            if f.shape != (self.channels,):
                raise ValueError("need one frequency per channel")
            check(lib().srcdsp_mixer_set_frequencies(self._h, f.ctypes.data_as(C.POINTER(C.c_float))))

    def reset(self, loFreq: float = 0.0, ch: int = -1):
        check(lib().srcdsp_mixer_reset(self._h, ch, float(loFreq)))

    def adjustFrequency(self, adjustFreq: float = 0.0, ch: int = -1):
        check(lib().srcdsp_mixer_adjust_frequency(self._h, ch, float(adjustFreq)))

    def state(self, ch: int = 0):
        phi, freq, nom = C.c_int(), C.c_int(), C.c_float()
        check(lib().srcdsp_mixer_get_state(self._h, ch, C.byref(phi), C.byref(freq), C.byref(nom)))
        return phi.value, freq.value, nom.value

    def set_state(self, ch: int, phi: int, freq: int, nominal: float):
        check(lib().srcdsp_mixer_set_state(self._h, ch, phi, freq, float(nominal)))

    def step(self, x, out=None):
        """mixers.h:168-188.  out may be x (in place)."""
        bi = _Buf(x, self.channels)
        if out is None:
            out = _alloc_like(x, self.channels, bi.n, bi.squeeze)
        bo = _Buf(out, self.channels)
        if bo.n < bi.n:
            raise ValueError("out is smaller than in")
        self._bind_stream(bi, "srcdsp_mixer_set_stream")
        check(lib().srcdsp_mixer_step(self._h, bi.ptr, bi.stride, bo.ptr, bo.stride, bi.n))
        return out

    def sync(self):
        check(lib().srcdsp_mixer_sync(self._h))


# ----------------------------------------------------------------------------------------------
class FilterDnsamplingFir(_Handle):
    """Polyphase decimating FIR bank (dsptl_dnsampling_filters.h:43-220).

    obsolete=True selects the behaviour of dnsampling_filters.h (no `taps % M == 0` assert).
    """

    _destroy = "srcdsp_dec_destroy"

    def __init__(self, M: int, firCoeff: Optional[Sequence[int]] = None, channels: int = 1,
                 device: int = 0, obsolete: bool = False):
        super().__init__()
        self.M, self.channels, self.device, self.obsolete = M, channels, device, obsolete
        check(lib().srcdsp_dec_create(C.byref(self._h), device, channels, M))
        if firCoeff is not None:
            self.setCoeffs(firCoeff)

    def setCoeffs(self, firCoeff):
        t = _taps(firCoeff)
        check(lib().srcdsp_dec_set_coeffs(self._h, t.ctypes.data_as(C.POINTER(C.c_int32)), t.size,
                                          0 if self.obsolete else 1))
        self.ntaps = t.size

    def setLeftShiftBy2(self, leftShiftBy2: int):
        check(lib().srcdsp_dec_set_left_shift(self._h, int(leftShiftBy2)))

    def reset(self):
        check(lib().srcdsp_dec_reset(self._h))

    def set_kernel(self, kind: int):
        check(lib().srcdsp_dec_set_kernel(self._h, kind))

    @property
    def last_kernel(self) -> str:
        v = C.c_int()
        check(lib().srcdsp_dec_get_last_kernel(self._h, C.byref(v)))
        return {0: "none", 1: "dec_fir_kernel (IMAD)", 2: "dec_tc_kernel (tcgen05 int8, LDG producers)",
                3: "dec_tma_kernel (tcgen05 int8, TMA-fed)",
                4: "dec_band_kernel (tcgen05 int8, TMA-fed, band form)"}[v.value]

    @property
    def coeffScaling(self) -> int:
        v = C.c_int()
        check(lib().srcdsp_dec_get_coeff_scaling(self._h, C.byref(v)))
        return v.value

    def history(self, ch: int = 0) -> np.ndarray:
        n = C.c_size_t()
        check(lib().srcdsp_dec_get_state(self._h, ch, None, C.byref(n)))
        h = np.zeros((n.value, 2), np.int16)
        check(lib().srcdsp_dec_get_state(self._h, ch, h.ctypes.data, C.byref(n)))
        return h

    def set_history(self, ch: int, hist: np.ndarray):
        h = np.ascontiguousarray(hist, np.int16).reshape(-1, 2)
        check(lib().srcdsp_dec_set_state(self._h, ch, h.ctypes.data, h.shape[0]))

    def step(self, x, out=None):
        """dsptl_dnsampling_filters.h:172-220: len(out) * M == len(in)."""
        bi = _Buf(x, self.channels)
        if out is None:
            if bi.n % self.M:
                raise SrcDspError(_capi.E_SIZE, f"n_in ({bi.n}) must be a multiple of M ({self.M})")
            out = _alloc_like(x, self.channels, bi.n // self.M, bi.squeeze)
        bo = _Buf(out, self.channels)
        if bo.n * self.M != bi.n:
            raise SrcDspError(_capi.E_SIZE, "filteredSignal.size() * M != input.size() "
                                            "[dsptl_dnsampling_filters.h:181]")
        self._bind_stream(bi, "srcdsp_dec_set_stream")
        check(lib().srcdsp_dec_step(self._h, bi.ptr, bi.stride, bi.n, bo.ptr, bo.stride))
        return out

    def sync(self):
        check(lib().srcdsp_dec_sync(self._h))


class FilterFir(FilterDnsamplingFir):
    """Non-decimating FIR bank (reference filters.h:42-169): in age order
    it is the M = 1 decimator with `limitScale16(y, coeffScaling)`; setCoeffs clears the history."""

    def __init__(self, firCoeff: Optional[Sequence[int]] = None, channels: int = 1, device: int = 0):
        super().__init__(1, None, channels, device, obsolete=True)
        if firCoeff is not None:
            self.setCoeffs(firCoeff)

    def setCoeffs(self, firCoeff):
        super().setCoeffs(firCoeff)
        self.reset()  # filters.h:96


# ----------------------------------------------------------------------------------------------
class FilterUpsamplingFir(_Handle):
    """Polyphase interpolating FIR bank (upsampling_filters.h:35-323)."""

    _destroy = "srcdsp_up_destroy"

    def __init__(self, L: int, firCoeff: Optional[Sequence[int]] = None, channels: int = 1,
                 device: int = 0):
        super().__init__()
        self.L, self.channels, self.device = L, channels, device
        check(lib().srcdsp_up_create(C.byref(self._h), device, channels, L))
        if firCoeff is not None and len(firCoeff):
            self.setCoefficients(firCoeff)

    def setCoefficients(self, firCoeff):
        t = _taps(firCoeff)
        check(lib().srcdsp_up_set_coefficients(self._h, t.ctypes.data_as(C.POINTER(C.c_int32)), t.size))

    def reset(self):
        check(lib().srcdsp_up_reset(self._h))

    def _geti(self, name) -> int:
        v = C.c_int()
        check(getattr(lib(), name)(self._h, C.byref(v)))
        return v.value

    def getLength(self) -> int:
        return self._geti("srcdsp_up_get_length")

    def getImpLength(self) -> int:
        return self._geti("srcdsp_up_get_imp_length")

    def getUpsamplingRatio(self) -> int:
        return self._geti("srcdsp_up_get_ratio")

    def history(self, ch: int = 0) -> np.ndarray:
        n = C.c_size_t()
        check(lib().srcdsp_up_get_state(self._h, ch, None, C.byref(n)))
        h = np.zeros((n.value, 2), np.int16)
        check(lib().srcdsp_up_get_state(self._h, ch, h.ctypes.data, C.byref(n)))
        return h

    def set_history(self, ch: int, hist: np.ndarray):
        h = np.ascontiguousarray(hist, np.int16).reshape(-1, 2)
        check(lib().srcdsp_up_set_state(self._h, ch, h.ctypes.data, h.shape[0]))

    def step(self, x, out=None, flush: bool = False, iterator_overload: bool = False):
        """Vector overload (upsampling_filters.h:149-233, shift 15 - round(log2 L)) or, with
        iterator_overload=True, the iterator overload (:240-323, shift 0)."""
        bi = _Buf(x, self.channels)
        n_out = self.L * (bi.n + (self.getLength() // self.L if flush else 0))
        if out is None:
            out = _alloc_like(x, self.channels, n_out, bi.squeeze)
        bo = _Buf(out, self.channels)
        if (not flush and not iterator_overload and bo.n != n_out) or bo.n < n_out:
            raise SrcDspError(_capi.E_SIZE, "signal.size() * L != filteredSignal.size() "
                                            "[upsampling_filters.h:153]")
        self._bind_stream(bi, "srcdsp_up_set_stream")
        check(lib().srcdsp_up_step(self._h, bi.ptr, bi.stride, bi.n, bo.ptr, bo.stride, int(flush),
                                   1 if iterator_overload else 0))
        return out

    @property
    def last_kernel(self) -> str:
        v = C.c_int()
        check(lib().srcdsp_up_get_last_kernel(self._h, C.byref(v)))
        return {0: "none", 1: "up_fir_kernel", 2: "up_fir4_kernel", 3: "up_tc_kernel (tcgen05 int8)",
                4: "up_tc2_kernel (tcgen05 int8, taps on the N side)"}[v.value]

    def sync(self):
        check(lib().srcdsp_up_sync(self._h))


# ----------------------------------------------------------------------------------------------
class _BufF32:
    """A [C, n, 2] float32 view (interleaved I/Q == complex64) of a numpy array or a torch CUDA tensor."""

    __slots__ = ("obj", "ptr", "C", "n", "stride", "device", "squeeze")

    def __init__(self, x, channels: int):
        self.obj = x
        shape = tuple(x.shape)
        self.squeeze = len(shape) == 2 and channels == 1
        if self.squeeze:
            shape = (1,) + shape
        if len(shape) != 3 or shape[2] != 2 or shape[0] != channels:
            raise ValueError(f"expected float32 [{channels}, n, 2] (or [n, 2] for one channel), got {tuple(x.shape)}")
        self.C, self.n = shape[0], shape[1]
        if _is_torch(x):
            import torch
            if x.dtype != torch.float32 or not x.is_cuda:
                raise TypeError("torch buffers must be CUDA float32 tensors")
            st = x.stride()
            if st[-1] != 1 or st[-2] != 2:
                raise ValueError("samples must be contiguous interleaved I/Q")
            self.stride = (st[0] // 2) if not self.squeeze else self.n
            self.ptr, self.device = x.data_ptr(), True
        else:
            if not isinstance(x, np.ndarray) or x.dtype != np.float32:
                raise TypeError("host buffers must be numpy float32 arrays")
            st = x.strides
            if st[-1] != 4 or st[-2] != 8:
                raise ValueError("samples must be contiguous interleaved I/Q")
            self.stride = (st[0] // 8) if not self.squeeze else self.n
            self.ptr, self.device = x.ctypes.data, False


class FilterDnsamplingFirFloat(_Handle):
    """dsptl::FilterDnsamplingFir<complex<float>, complex<float>, complex<float>, float, M>
    (dsptl_dnsampling_filters.h:43-220), the float instantiation of the decimator: bit-exact with the compiled
    reference, including its int32 truncation / limitScale16 of the float sum (:214) and its integer abs() in
    coeffScaling (:128-132).  Buffers: float32 [C, n, 2] / [n, 2] numpy arrays or torch CUDA tensors."""

    _destroy = "srcdsp_decf_destroy"

    def __init__(self, M: int, firCoeff=None, channels: int = 1, device: int = 0, obsolete: bool = False):
        super().__init__()
        self.M, self.channels, self.device, self.obsolete = M, channels, device, obsolete
        check(lib().srcdsp_decf_create(C.byref(self._h), device, channels, M))
        if firCoeff is not None:
            self.setCoeffs(firCoeff)

    def setCoeffs(self, firCoeff):
        t = np.ascontiguousarray(np.asarray(firCoeff, dtype=np.float32).reshape(-1))
        check(lib().srcdsp_decf_set_coeffs(self._h, t.ctypes.data_as(C.POINTER(C.c_float)), t.size,
                                           0 if self.obsolete else 1))
        self.ntaps = t.size

    def setLeftShiftBy2(self, leftShiftBy2: int):
        check(lib().srcdsp_decf_set_left_shift(self._h, int(leftShiftBy2)))

    def reset(self):
        check(lib().srcdsp_decf_reset(self._h))

    @property
    def coeffScaling(self) -> int:
        v = C.c_uint()
        check(lib().srcdsp_decf_get_coeff_scaling(self._h, C.byref(v)))
        return v.value

    @property
    def last_kernel(self) -> str:
        v = C.c_int()
        check(lib().srcdsp_decf_get_last_kernel(self._h, C.byref(v)))
        return {1: "decf_fir_kernel (FP32 pipe, packed FFMA2 + FADD2 chains, output pair per thread)",
                2: "decf_quad_kernel (FP32 pipe, packed FFMA2 + FADD2 chains, four outputs per thread)"}.get(v.value, "none")

    def step(self, x, out=None):
        """dsptl_dnsampling_filters.h:172-220: out.size() * M == in.size()."""
        bi = _BufF32(x, self.channels)
        if bi.n % self.M:
            raise SrcDspError(_capi.E_SIZE, "filteredSignal.size() * M != input.size() [dsptl_dnsampling_filters.h:181]")
        n_out = bi.n // self.M
        if out is None:
            shape = (n_out, 2) if bi.squeeze else (self.channels, n_out, 2)
            if _is_torch(x):
                import torch
                out = torch.empty(shape, dtype=torch.float32, device=x.device)
            else:
                out = np.empty(shape, np.float32)
        bo = _BufF32(out, self.channels)
        if bo.n != n_out:
            raise SrcDspError(_capi.E_SIZE, "filteredSignal.size() * M != input.size() [dsptl_dnsampling_filters.h:181]")
        self._bind_stream(bi, "srcdsp_decf_set_stream")
        check(lib().srcdsp_decf_step(self._h, bi.ptr, bi.stride, bi.n, bo.ptr, bo.stride))
        return out

    def sync(self):
        check(lib().srcdsp_decf_sync(self._h))


class FilterFirFloat(FilterDnsamplingFirFloat):
    """FilterFir<complex<float>, complex<float>, complex<float>, float> (reference filters.h:42-169 with float types):
    in age order the M = 1 float decimator; setCoeffs clears the history (filters.h:96)."""

    def __init__(self, firCoeff=None, channels: int = 1, device: int = 0):
        super().__init__(1, None, channels, device, obsolete=True)
        if firCoeff is not None:
            self.setCoeffs(firCoeff)

    def setCoeffs(self, firCoeff):
        super().setCoeffs(firCoeff)
        self.reset()


# ----------------------------------------------------------------------------------------------
class Ddc(_Handle):
    """Fused chain mixer -> dec1 [-> dec2]; bit-identical to the three separate steps."""

    _destroy = "srcdsp_ddc_destroy"

    def __init__(self, mixer: Optional[Mixer], dec1: FilterDnsamplingFir,
                 dec2: Optional[FilterDnsamplingFir] = None):
        super().__init__()
        self.mixer, self.dec1, self.dec2 = mixer, dec1, dec2  # keep alive
        self.channels = dec1.channels
        self.M = dec1.M * (dec2.M if dec2 else 1)
        check(lib().srcdsp_ddc_create(C.byref(self._h), mixer._h if mixer else None, dec1._h,
                                      dec2._h if dec2 else None))

    def step(self, x, out=None):
        bi = _Buf(x, self.channels)
        if bi.n % self.M:
            raise SrcDspError(_capi.E_SIZE, f"n_in ({bi.n}) must be a multiple of {self.M}")
        if out is None:
            out = _alloc_like(x, self.channels, bi.n // self.M, bi.squeeze)
        bo = _Buf(out, self.channels)
        if bo.n * self.M != bi.n:
            raise SrcDspError(_capi.E_SIZE, "out.size() * M_total != in.size()")
        self._bind_stream(bi, "srcdsp_ddc_set_stream")
        check(lib().srcdsp_ddc_step(self._h, bi.ptr, bi.stride, bi.n, bo.ptr, bo.stride))
        return out

    def sync(self):
        check(lib().srcdsp_ddc_sync(self._h))


class DdcGroup(_Handle):
    """One chain (mixer -> dec1 [-> dec2]) over several GPUs of one box, driven from this process (srcdsp_group_*):
    mode "channels" = contiguous channel batches per device, mode "slices" = time slices of one long stream with a
    warm-up halo.  Buffers are HOST numpy arrays (views of PinnedBuffer for speed); every device copies its own part
    and writes its outputs to their final place in `out`."""

    _destroy = "srcdsp_group_destroy"

    def __init__(self, mode: str, devices: Sequence[int], channels: int, M1: int, taps1, M2: int = 0, taps2=None,
                 n_table: int = 0, obsolete: bool = True):
        super().__init__()
        dv = (C.c_int * len(devices))(*devices)
        self.mode, self.channels, self.M = mode, channels, M1 * (M2 if M2 else 1)
        check(lib().srcdsp_group_create(C.byref(self._h), {"channels": 0, "slices": 1}[mode], dv, len(devices), channels,
                                        n_table, M1, M2))
        t = _taps(taps1)
        check(lib().srcdsp_group_set_coeffs(self._h, 1, t.ctypes.data_as(C.POINTER(C.c_int32)), t.size, 0 if obsolete else 1))
        if M2:
            t = _taps(taps2)
            check(lib().srcdsp_group_set_coeffs(self._h, 2, t.ctypes.data_as(C.POINTER(C.c_int32)), t.size, 0 if obsolete else 1))

    def setFrequency(self, loFreq):
        f = np.ascontiguousarray(np.broadcast_to(np.asarray(loFreq, np.float32), (self.channels,)))
        check(lib().srcdsp_group_set_frequencies(self._h, f.ctypes.data_as(C.POINTER(C.c_float))))

    def reset(self):
        check(lib().srcdsp_group_reset(self._h))

    def layout(self):
        n, used = C.c_int(), C.c_int()
        check(lib().srcdsp_group_size(self._h, C.byref(n), C.byref(used)))
        out = []
        for i in range(n.value):
            d, c0, nc = C.c_int(), C.c_int(), C.c_int()
            check(lib().srcdsp_group_get_layout(self._h, i, C.byref(d), C.byref(c0), C.byref(nc)))
            out.append((d.value, c0.value, nc.value))
        return out, used.value

    def step(self, x, out=None):
        bi = _Buf(x, self.channels)
        if bi.device:
            raise TypeError("a DdcGroup takes host buffers")
        if bi.n % self.M:
            raise SrcDspError(_capi.E_SIZE, f"n_in ({bi.n}) must be a multiple of {self.M}")
        if out is None:
            out = _alloc_like(x, self.channels, bi.n // self.M, bi.squeeze)
        bo = _Buf(out, self.channels)
        if bo.n * self.M != bi.n:
            raise SrcDspError(_capi.E_SIZE, "out.size() * M_total != in.size()")
        check(lib().srcdsp_group_step(self._h, bi.ptr, bi.stride, bi.n, bo.ptr, bo.stride))
        return out


# ----------------------------------------------------------------------------------------------
def synth_fill(t, seed: int, ch0: int = 0, n0: int = 0, amp_shift: int = 0):
    """Fill a torch CUDA int16 tensor [C, n, 2] with the counter-based synthetic baseband."""
    import torch
    b = _Buf(t, t.shape[0] if t.dim() == 3 else 1)
    s = torch.cuda.current_stream().cuda_stream or 1
    check(lib().srcdsp_synth_fill(t.device.index or 0, C.c_void_p(s), b.ptr, b.stride, b.C, b.n,
                                  seed & 0xFFFFFFFF, ch0, n0, amp_shift))
    return t


def launch_count() -> int:
    return int(lib().srcdsp_launch_count())


def device_count() -> int:
    n = C.c_int()
    st = lib().srcdsp_device_count(C.byref(n))
    return n.value if st == 0 else 0


class PinnedBuffer:
    """Pinned host memory from the library (srcdsp_host_alloc) viewed as numpy int16 [C, n, 2]."""

    def __init__(self, channels: int, n: int):
        self._p = C.c_void_p()
        self.nbytes = channels * n * 4
        check(lib().srcdsp_host_alloc(C.byref(self._p), self.nbytes))
        buf = (C.c_int16 * (channels * n * 2)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=np.int16).reshape(channels, n, 2)

    def free(self):
        if self._p:
            self.array = None
            lib().srcdsp_host_free(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------
# the stages either side of the GPU path
# ----------------------------------------------------------------------------------------------
class FifoWithTimeTrack(_Handle):
    """dsptl::FifoWithTimeTrack<std::complex<int16_t>, N> (buffers.h:58-459): single-writer /
    single-reader ring with sample time stamps, in pinned host memory (srcdsp_fifo_*).  Elements are
    complex int16 samples, numpy int16 [n, 2]."""
    _destroy = "srcdsp_fifo_destroy"

    def __init__(self, capacity: int, samplingFrequency: float = 0.0):
        super().__init__()
        self.capacity = int(capacity)
        check(lib().srcdsp_fifo_create(C.byref(self._h), 4, self.capacity, float(samplingFrequency)))

    @property
    def pinned(self) -> bool:
        return bool(lib().srcdsp_fifo_is_pinned(self._h))

    def write(self, x: np.ndarray, seconds: int = 0, fracSeconds: float = 0.0):
        x = np.ascontiguousarray(x, dtype=np.int16).reshape(-1, 2)
        check(lib().srcdsp_fifo_write(self._h, x.ctypes.data, x.shape[0], int(seconds), float(fracSeconds)))

    def read(self, n: int, start: int):
        """Returns (error, start, samples): `error` is the reference's return value (True: range not
        available; samples is then None), `start` the possibly adjusted first time point."""
        out = np.empty((n, 2), np.int16)
        st, err = C.c_uint64(start), C.c_int()
        check(lib().srcdsp_fifo_read(self._h, out.ctypes.data, n, C.byref(st), C.byref(err)))
        return bool(err.value), st.value, (None if err.value else out)

    def readSegments(self, n: int, start: int):
        """Zero-copy read: (error, start, [views of the pinned ring]) -- at most two pieces."""
        st, err = C.c_uint64(start), C.c_int()
        p0, p1, n0, n1 = C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_size_t()
        check(lib().srcdsp_fifo_segments(self._h, n, C.byref(st), C.byref(p0), C.byref(n0), C.byref(p1), C.byref(n1),
                                         C.byref(err)))
        segs = []
        for p, k in ((p0, n0.value), (p1, n1.value)):
            if k:
                buf = (C.c_int16 * (2 * k)).from_address(p.value)
                segs.append(np.frombuffer(buf, dtype=np.int16).reshape(k, 2))
        return bool(err.value), st.value, segs

    def count(self) -> int:
        n = C.c_size_t()
        check(lib().srcdsp_fifo_count(self._h, C.byref(n)))
        return n.value

    def reset(self):
        check(lib().srcdsp_fifo_reset(self._h))

    def getAbsoluteTime(self, timePoint: int, fracTimePoint: float = 0.0):
        s, f = C.c_uint(), C.c_double()
        check(lib().srcdsp_fifo_get_absolute_time(self._h, int(timePoint), float(fracTimePoint), C.byref(s), C.byref(f)))
        return s.value, f.value

    def state(self):
        """(writePtr, timeStart, timeEnd, rolloverFlag): what dumpInfo prints."""
        wp, ts, te, ro = C.c_size_t(), C.c_uint64(), C.c_uint64(), C.c_int()
        check(lib().srcdsp_fifo_get_state(self._h, C.byref(wp), C.byref(ts), C.byref(te), C.byref(ro)))
        return wp.value, ts.value, te.value, bool(ro.value)

    def _set_time(self, timeStart: int, timeEnd: int):
        check(lib().srcdsp_fifo_set_time(self._h, int(timeStart), int(timeEnd)))


def saveBinarySamples(x: np.ndarray, path: str):
    """dsptl::saveBinarySamples (dsptl_files.h:101-109): raw interleaved I/Q, no header."""
    np.ascontiguousarray(x, dtype=np.int16).reshape(-1, 2).tofile(path)


def readBinarySamples(path: str, out: Optional[np.ndarray] = None, exact: bool = False) -> np.ndarray:
    """dsptl::readBinarySamples (dsptl_files.h:250-262) for complex int16.  Like the reference it APPENDS to
    `out` (its `out.empty()` is a no-op) and, unless exact=True, adds one element after the last complete
    sample: the reference's `while (is)` loop pushes once more with the values the failed read left behind
    (the last sample, with the bytes of a trailing partial sample on top; (0, 0) for an empty file)."""
    raw = np.fromfile(path, dtype=np.uint8)
    n = raw.size // 4
    x = raw[: 4 * n].view(np.int16).reshape(n, 2).copy()
    if not exact:
        last = (x[-1].copy() if n else np.zeros(2, np.int16)).view(np.uint8)
        rest = raw[4 * n:]
        last[: rest.size] = rest  # istream::read stores the bytes of a trailing partial sample it did get
        x = np.concatenate([x, last.view(np.int16)[None, :]])
    if out is not None and len(out):
        x = np.concatenate([np.asarray(out, np.int16).reshape(-1, 2), x])
    return x


class FixedPatternCorrelator(_Handle):
    """dsptl::FixedPatternCorrelator<int16_t, int32_t, N, S> (correlators.h:54-303): the stage behind the DDC.
    A bank of `channels` independent correlators sharing one pattern."""
    _destroy = "srcdsp_corr_destroy"

    def __init__(self, N: int = 32, S: int = 4, channels: int = 1, device: int = 0):
        super().__init__()
        self.N, self.S, self.channels = N, S, channels
        check(lib().srcdsp_corr_create(C.byref(self._h), device, channels, N, S))

    def setPattern(self, pattern, thresholdCoeff: float = 0.8):
        """pattern: N complex values as int32 [N, 2] (the replica of the wanted signal, not conjugated)."""
        p = np.ascontiguousarray(pattern, dtype=np.int32).reshape(self.N, 2)
        check(lib().srcdsp_corr_set_pattern(self._h, p.ctypes.data_as(C.POINTER(C.c_int32)), float(thresholdCoeff)))

    def reset(self):
        check(lib().srcdsp_corr_reset(self._h))

    def step(self, x):
        """x: [n, 2] / [C, n, 2] int16 (numpy or torch CUDA).  Returns (found, corrIndex) -- scalars for a single
        channel given a 2-D block, arrays [C] otherwise.  corrIndex is only meaningful where found."""
        b = _Buf(x, self.channels)
        self._bind_stream(b, "srcdsp_corr_set_stream")
        found = (C.c_int * self.channels)()
        idx = (C.c_int * self.channels)()
        check(lib().srcdsp_corr_step(self._h, b.ptr, b.stride, b.n, found, idx))
        f, i = np.array(found[:], dtype=bool), np.array(idx[:], dtype=np.int64)
        if self.channels == 1 and b.squeeze:
            return bool(f[0]), int(i[0])
        return f, i

    def getRefBitSamples(self, ch: int = 0) -> np.ndarray:
        out = np.zeros((self.N, 2), np.int16)
        check(lib().srcdsp_corr_get_ref_bit_samples(self._h, ch, out.ctypes.data))
        return out

    def getStatus(self, ch: int = 0) -> dict:
        e, c = (C.c_uint32 * 3)(), (C.c_uint32 * 3)()
        ce, cs, tf = C.c_uint32(), C.c_int(), C.c_double()
        check(lib().srcdsp_corr_get_status(self._h, ch, e, c, C.byref(ce), C.byref(cs), C.byref(tf)))
        return dict(energyValue=list(e), corrValue=list(c), coeffsEnergy=ce.value, coeffScaling=cs.value,
                    thresholdFactor=tf.value)
