"""ctypes binding of libsrcdsp_b200.so (the C ABI in include/srcdsp_b200.h).

The library is the product; this file only declares its prototypes.  There is no Python or
CPU fallback: if the shared library is missing, `lib()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsrcdsp_b200.so")

OK = 0
E_INVALID, E_SIZE, E_CUDA, E_NOMEM, E_STATE, E_NOGPU = -1, -2, -3, -4, -5, -6

_vp = C.c_void_p
_sz = C.c_size_t
_i16p = C.c_void_p  # buffers are passed as raw addresses (host or device)
_i32p = C.POINTER(C.c_int32)

# name -> (restype, argtypes); kept in one table so tests can check it against the header
PROTOTYPES = {
    "srcdsp_last_error": (C.c_char_p, []),
    "srcdsp_version": (C.c_int, []),
    "srcdsp_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "srcdsp_launch_count": (C.c_uint64, []),
    "srcdsp_host_alloc": (C.c_int, [C.POINTER(_vp), _sz]),
    "srcdsp_host_free": (C.c_int, [_vp]),
    "srcdsp_device_alloc": (C.c_int, [C.c_int, C.POINTER(_vp), _sz]),
    "srcdsp_device_free": (C.c_int, [C.c_int, _vp]),
    "srcdsp_memcpy": (C.c_int, [C.c_int, _vp, _vp, _sz]),
    "srcdsp_synth_fill": (C.c_int, [C.c_int, _vp, _i16p, _sz, C.c_int, _sz, C.c_uint32, C.c_uint32,
                                    C.c_uint64, C.c_int]),
    # mixer
    "srcdsp_mixer_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_uint]),
    "srcdsp_mixer_destroy": (C.c_int, [_vp]),
    "srcdsp_mixer_set_frequency": (C.c_int, [_vp, C.c_int, C.c_float]),
    "srcdsp_mixer_set_frequencies": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "srcdsp_mixer_reset": (C.c_int, [_vp, C.c_int, C.c_float]),
    "srcdsp_mixer_adjust_frequency": (C.c_int, [_vp, C.c_int, C.c_float]),
    "srcdsp_mixer_step": (C.c_int, [_vp, _i16p, _sz, _i16p, _sz, _sz]),
    "srcdsp_mixer_get_state": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                         C.POINTER(C.c_float)]),
    "srcdsp_mixer_set_state": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_float]),
    "srcdsp_mixer_set_stream": (C.c_int, [_vp, _vp]),
    "srcdsp_mixer_sync": (C.c_int, [_vp]),
    # decimator
    "srcdsp_dec_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int]),
    "srcdsp_dec_destroy": (C.c_int, [_vp]),
    "srcdsp_dec_set_coeffs": (C.c_int, [_vp, _i32p, C.c_int, C.c_int]),
    "srcdsp_dec_set_left_shift": (C.c_int, [_vp, C.c_int]),
    "srcdsp_dec_reset": (C.c_int, [_vp]),
    "srcdsp_dec_step": (C.c_int, [_vp, _i16p, _sz, _sz, _i16p, _sz]),
    "srcdsp_dec_get_coeff_scaling": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "srcdsp_dec_get_state": (C.c_int, [_vp, C.c_int, _i16p, C.POINTER(_sz)]),
    "srcdsp_dec_set_state": (C.c_int, [_vp, C.c_int, _i16p, _sz]),
    "srcdsp_dec_set_stream": (C.c_int, [_vp, _vp]),
    "srcdsp_dec_sync": (C.c_int, [_vp]),
    "srcdsp_dec_set_kernel": (C.c_int, [_vp, C.c_int]),
    "srcdsp_dec_get_last_kernel": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    # float decimator (the reference's complex<float> instantiation)
    "srcdsp_decf_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int]),
    "srcdsp_decf_destroy": (C.c_int, [_vp]),
    "srcdsp_decf_set_coeffs": (C.c_int, [_vp, C.POINTER(C.c_float), C.c_int, C.c_int]),
    "srcdsp_decf_set_left_shift": (C.c_int, [_vp, C.c_int]),
    "srcdsp_decf_reset": (C.c_int, [_vp]),
    "srcdsp_decf_step": (C.c_int, [_vp, _vp, _sz, _sz, _vp, _sz]),
    "srcdsp_decf_get_coeff_scaling": (C.c_int, [_vp, C.POINTER(C.c_uint)]),
    "srcdsp_decf_get_last_kernel": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "srcdsp_decf_get_state": (C.c_int, [_vp, C.c_int, _vp, C.POINTER(_sz)]),
    "srcdsp_decf_set_state": (C.c_int, [_vp, C.c_int, _vp, _sz]),
    "srcdsp_decf_set_stream": (C.c_int, [_vp, _vp]),
    "srcdsp_decf_sync": (C.c_int, [_vp]),
    # fused chain
    "srcdsp_ddc_create": (C.c_int, [C.POINTER(_vp), _vp, _vp, _vp]),
    "srcdsp_ddc_destroy": (C.c_int, [_vp]),
    "srcdsp_ddc_step": (C.c_int, [_vp, _i16p, _sz, _sz, _i16p, _sz]),
    "srcdsp_ddc_set_stream": (C.c_int, [_vp, _vp]),
    "srcdsp_ddc_sync": (C.c_int, [_vp]),
    # device group (one chain over several GPUs, one process)
    "srcdsp_group_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_uint, C.c_int, C.c_int]),
    "srcdsp_group_destroy": (C.c_int, [_vp]),
    "srcdsp_group_set_coeffs": (C.c_int, [_vp, C.c_int, _i32p, C.c_int, C.c_int]),
    "srcdsp_group_set_frequencies": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "srcdsp_group_reset": (C.c_int, [_vp]),
    "srcdsp_group_step": (C.c_int, [_vp, _i16p, _sz, _sz, _i16p, _sz]),
    "srcdsp_group_get_layout": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "srcdsp_group_size": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    # upsampler
    "srcdsp_up_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int]),
    "srcdsp_up_destroy": (C.c_int, [_vp]),
    "srcdsp_up_set_coefficients": (C.c_int, [_vp, _i32p, C.c_int]),
    "srcdsp_up_reset": (C.c_int, [_vp]),
    "srcdsp_up_step": (C.c_int, [_vp, _i16p, _sz, _sz, _i16p, _sz, C.c_int, C.c_int]),
    "srcdsp_up_get_length": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "srcdsp_up_get_imp_length": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "srcdsp_up_get_ratio": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "srcdsp_up_get_last_kernel": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "srcdsp_up_get_state": (C.c_int, [_vp, C.c_int, _i16p, C.POINTER(_sz)]),
    "srcdsp_up_set_state": (C.c_int, [_vp, C.c_int, _i16p, _sz]),
    "srcdsp_up_set_stream": (C.c_int, [_vp, _vp]),
    "srcdsp_up_sync": (C.c_int, [_vp]),
    # fifo (buffers.h)
    "srcdsp_fifo_create": (C.c_int, [C.POINTER(_vp), _sz, _sz, C.c_double]),
    "srcdsp_fifo_destroy": (C.c_int, [_vp]),
    "srcdsp_fifo_is_pinned": (C.c_int, [_vp]),
    "srcdsp_fifo_write": (C.c_int, [_vp, _vp, _sz, C.c_uint, C.c_double]),
    "srcdsp_fifo_read": (C.c_int, [_vp, _vp, _sz, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]),
    "srcdsp_fifo_segments": (C.c_int, [_vp, _sz, C.POINTER(C.c_uint64), C.POINTER(_vp), C.POINTER(_sz), C.POINTER(_vp),
                                       C.POINTER(_sz), C.POINTER(C.c_int)]),
    "srcdsp_fifo_count": (C.c_int, [_vp, C.POINTER(_sz)]),
    "srcdsp_fifo_reset": (C.c_int, [_vp]),
    "srcdsp_fifo_get_absolute_time": (C.c_int, [_vp, C.c_uint64, C.c_double, C.POINTER(C.c_uint), C.POINTER(C.c_double)]),
    "srcdsp_fifo_get_state": (C.c_int, [_vp, C.POINTER(_sz), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]),
    "srcdsp_fifo_set_time": (C.c_int, [_vp, C.c_uint64, C.c_uint64]),
    "srcdsp_fifo_storage": (_vp, [_vp]),
    # correlator (correlators.h)
    "srcdsp_corr_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int]),
    "srcdsp_corr_destroy": (C.c_int, [_vp]),
    "srcdsp_corr_set_pattern": (C.c_int, [_vp, _i32p, C.c_double]),
    "srcdsp_corr_reset": (C.c_int, [_vp]),
    "srcdsp_corr_step": (C.c_int, [_vp, _i16p, _sz, _sz, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "srcdsp_corr_get_ref_bit_samples": (C.c_int, [_vp, C.c_int, _i16p]),
    "srcdsp_corr_get_status": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                         C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "srcdsp_corr_set_stream": (C.c_int, [_vp, _vp]),
}

_lib = None


class SrcDspError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"srcdsp error {code}: {msg}")
        self.code = code


def lib() -> C.CDLL:
    """Load the C-ABI library.  Fails loudly: no CUDA library, no product."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m srcdsp_b200.build` "
                "(nvcc, sm_100a).  srcdsp_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int) -> None:
    if status != OK:
        raise SrcDspError(status, lib().srcdsp_last_error().decode(errors="replace"))
