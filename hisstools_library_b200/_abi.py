"""ctypes binding of libhisstools_b200.so -- exactly the symbols include/hisstools_b200.h declares.

There is no fallback: if the CUDA library is missing it is built with nvcc, and if that is not
possible importing this module raises.  Compute entry points return HB_ERR_CUDA (-1) on a machine
without a usable device and the wrappers turn that into HissError.
"""
import ctypes as C
import os

from . import build as _build

HB_F32, HB_F64 = 0, 1
HB_OK, HB_ERR_CUDA, HB_ERR_BAD_ARG, HB_ERR_UNSUPPORTED, HB_ERR_NO_IR, HB_ERR_BUSY = 0, -1, -2, -3, -4, -5

UP = C.c_size_t        # uintptr_t
IP = C.c_ssize_t       # intptr_t
V = C.c_void_p
U32 = C.c_uint32

# name -> (restype, argtypes); mirrors include/hisstools_b200.h one to one
SIGNATURES = {
    "hb_last_error": (C.c_char_p, []),
    "hb_launch_count": (C.c_uint64, []),
    "hb_version": (C.c_char_p, []),
    "hb_fft_setup_create": (C.c_int, [C.POINTER(V), C.c_int, UP, C.c_int]),
    "hb_fft_setup_destroy": (None, [V]),
    "hb_fft": (C.c_int, [V, V, V, UP]),
    "hb_ifft": (C.c_int, [V, V, V, UP]),
    "hb_rfft": (C.c_int, [V, V, V, UP]),
    "hb_rifft": (C.c_int, [V, V, V, UP]),
    "hb_rfft_real": (C.c_int, [V, V, C.c_int, V, V, UP, UP]),
    "hb_rifft_real": (C.c_int, [V, V, V, V, UP]),
    "hb_rfft_real_batched_dev": (C.c_int, [V, V, V, V, UP, UP, UP, UP, V]),
    "hb_rifft_real_batched_dev": (C.c_int, [V, V, V, V, UP, UP, UP, UP, V]),
    "hb_conv_create": (C.c_int, [C.POINTER(V), C.c_int, U32, U32, U32, UP, UP, UP, UP, C.c_int]),
    "hb_conv_destroy": (None, [V]),
    "hb_conv_set_fft_size": (C.c_int, [V, UP]),
    "hb_conv_set_length": (C.c_int, [V, UP]),
    "hb_conv_set_offset": (C.c_int, [V, UP]),
    "hb_conv_set_reset_offset": (C.c_int, [V, IP]),
    "hb_conv_set_ir": (C.c_int, [V, U32, U32, U32, V, C.c_int, UP]),
    "hb_conv_set_ir_live": (C.c_int, [V, U32, U32, U32, V, C.c_int, UP]),
    "hb_conv_reset_pair": (C.c_int, [V, U32, U32, U32]),
    "hb_conv_set_ir_dev": (C.c_int, [V, U32, U32, U32, V, UP]),
    "hb_conv_resize": (C.c_int, [V, UP]),
    "hb_conv_reset": (C.c_int, [V]),
    "hb_conv_partitions": (UP, [V]),
    "hb_conv_max_length": (UP, [V]),
    "hb_conv_fft_size": (UP, [V]),
    "hb_conv_process": (C.c_int, [V, C.POINTER(V), C.POINTER(V), UP, C.c_int]),
    "hb_conv_process_dev": (C.c_int, [V, V, UP, V, UP, UP, C.c_int, V]),
    "hb_conv_shard_export": (C.c_int, [V, U32, U32, V]),
    "hb_conv_shard_attach": (C.c_int, [V, V]),
    "hb_conv_process_shard_dev": (C.c_int, [V, V, UP, V, UP, UP, C.c_int, V]),
    "hb_conv_shard_attach_local": (C.c_int, [V, C.POINTER(V)]),
    "hb_conv_shard_status": (C.c_int, [V, C.POINTER(U32)]),
    "hb_conv_join": (C.c_int, [V, V]),
    "hb_conv_set_tuning": (C.c_int, [V, C.c_int, C.c_int]),
    "hb_conv_set_host_pipeline": (C.c_int, [V, C.c_int]),
    "hb_conv_bytes_per_hop": (C.c_uint64, [V]),
    "hb_matrix_create": (C.c_int, [C.POINTER(V), C.c_int, U32, U32, U32, UP, C.c_int, U32, U32, U32, U32, C.c_int]),
    "hb_matrix_create_latency": (C.c_int, [C.POINTER(V), C.c_int, U32, U32, U32, UP, C.c_int, C.c_int]),
    "hb_matrix_create_multi": (C.c_int, [C.POINTER(V), C.c_int, U32, U32, U32, UP, C.c_int, U32, U32, U32, U32, C.POINTER(C.c_int), U32]),
    "hb_matrix_create_latency_multi": (C.c_int, [C.POINTER(V), C.c_int, U32, U32, U32, UP, C.c_int, C.POINTER(C.c_int), U32]),
    "hb_matrix_shards": (U32, [V]),
    "hb_matrix_shard": (V, [V, U32]),
    "hb_matrix_exchange": (C.c_int, [V]),
    "hb_matrix_destroy": (None, [V]),
    "hb_matrix_set_reset_offset": (C.c_int, [V, IP]),
    "hb_matrix_resize": (C.c_int, [V, U32, U32, U32, UP]),
    "hb_matrix_set": (C.c_int, [V, U32, U32, U32, V, C.c_int, UP, C.c_int]),
    "hb_matrix_reset": (C.c_int, [V]),
    "hb_matrix_reset_pair": (C.c_int, [V, U32, U32, U32]),
    "hb_matrix_set_hop_overlap": (C.c_int, [V, C.c_int]),
    "hb_matrix_process": (C.c_int, [V, C.POINTER(V), C.POINTER(V), UP, C.c_int]),
    "hb_matrix_process_dev": (C.c_int, [V, V, UP, V, UP, UP, C.c_int, V]),
    "hb_matrix_parts": (U32, [V]),
    "hb_matrix_part": (V, [V, U32]),
    "hb_matrix_head_taps": (U32, [V]),
    "hb_spectral_create": (C.c_int, [C.POINTER(V), C.c_int, UP, C.c_int]),
    "hb_spectral_destroy": (None, [V]),
    "hb_spectral_set_max_fft_size": (C.c_int, [V, UP]),
    "hb_spectral_max_fft_size": (UP, [V]),
    "hb_spectral_convolved_size": (UP, [V, UP, UP, C.c_int]),
    "hb_spectral_convolve": (C.c_int, [V, V, V, UP, V, UP, C.c_int, C.POINTER(UP)]),
    "hb_spectral_correlate": (C.c_int, [V, V, V, UP, V, UP, C.c_int, C.POINTER(UP)]),
    "hb_spectral_change_phase": (C.c_int, [V, V, V, UP, C.c_double, C.c_double, C.POINTER(UP)]),
    "hb_spectral_convolve_complex": (C.c_int, [V, V, V, V, UP, V, UP, V, UP, V, UP, C.c_int, C.POINTER(UP)]),
    "hb_spectral_correlate_complex": (C.c_int, [V, V, V, V, UP, V, UP, V, UP, V, UP, C.c_int, C.POINTER(UP)]),
    "hb_audio_probe": (C.c_int, [C.c_char_p, V]),
    "hb_audio_read": (C.c_int, [C.c_char_p, U32, U32, C.c_int32, V, C.c_int, C.c_int]),
    "hb_audio_read_raw": (C.c_int, [C.c_char_p, U32, U32, V]),
    "hb_audio_writer_open": (C.c_int, [C.POINTER(V), C.c_char_p, C.c_int, C.c_int, U32, C.c_double, C.c_int]),
    "hb_audio_writer_write": (C.c_int, [V, V, C.c_int, U32, C.c_int32]),
    "hb_audio_writer_write_raw": (C.c_int, [V, V, U32]),
    "hb_audio_writer_seek": (C.c_int, [V, U32]),
    "hb_audio_writer_position": (U32, [V]),
    "hb_audio_writer_info": (C.c_int, [V, V, C.POINTER(C.c_int)]),
    "hb_audio_writer_close": (None, [V]),
    "hb_audio_decode_dev": (C.c_int, [V, V, C.c_uint64, C.c_int32, V, C.c_uint64, C.c_int, C.c_int, V]),
    "hb_conv_set_ir_file": (C.c_int, [V, U32, U32, U32, C.c_char_p, U32, C.c_int]),
    "hb_conv_dtype": (C.c_int, [V]),
    "hb_conv_set_profiling": (C.c_int, [V, C.c_int]),
    "hb_conv_get_profile": (C.c_int, [V, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "hb_conv_set_multi_hop": (C.c_int, [V, C.c_int]),
    "hb_conv_set_fft_path": (C.c_int, [V, C.c_int]),
    "hb_conv_fft_path": (C.c_int, [V]),
    "hb_conv_set_trace": (C.c_int, [V, C.c_int]),
    "hb_conv_get_trace": (C.c_int, [V, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "hb_conv_set_schedule": (C.c_int, [V, C.c_int]),
    "hb_conv_set_tail_streams": (C.c_int, [V, C.c_int]),
    "hb_conv_set_hop_overlap": (C.c_int, [V, C.c_int]),
    "hb_conv_tail_streams": (C.c_int, [V]),
    "hb_conv_schedule": (C.c_int, [V]),
    "hb_conv_bytes_per_launch": (C.c_uint64, [V]),
}


class HissError(RuntimeError):
    """A negative hb_status from the library (CUDA failure, bad argument, unsupported size)."""

    def __init__(self, code, message):
        super().__init__("hisstools_b200 error %d: %s" % (code, message))
        self.code = code


_lib = None


def lib():
    """The loaded library (built on first use when missing or stale)."""
    global _lib
    if _lib is None:
        # HISSTOOLS_B200_LIB: an explicit (pre-built) library instead of the in-tree one
        path = os.environ.get("HISSTOOLS_B200_LIB") or _build.build()
        handle = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError = header and library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error():
    return lib().hb_last_error().decode("utf-8", "replace")


def check(code):
    """Pass reference ConvolveError codes (>= 0) through; raise on library failures (< 0, except NO_IR / BUSY)."""
    if code < 0 and code not in (HB_ERR_NO_IR, HB_ERR_BUSY):
        raise HissError(code, last_error())
    return code
