"""spectral_processor -- one-shot FFT convolution of two real buffers on the B200.

Mirror of the convolution part of the reference's spectral_processor<T> (SpectralProcessor.hpp:11-683):
same class name, `EdgeMode` values, `convolve(output, in1, in2, mode)`, `convolved_size`,
`set_max_fft_size` / `max_fft_size`.  correlate / change_phase are outside the convolution path
(SURVEY 8f-4).  The transforms, the per-bin product and the edge-mode arrangement run as CUDA kernels
behind hb_spectral_* of include/hisstools_b200.h.
"""
import ctypes as C
import enum

import numpy as np

from . import _abi


class EdgeMode(enum.IntEnum):
    """spectral_processor::EdgeMode (SpectralProcessor.hpp:22)"""
    Linear = 0
    Wrap = 1
    WrapCentre = 2
    Fold = 3
    FoldRepeat = 4


class spectral_processor:
    EdgeMode = EdgeMode

    def __init__(self, max_fft_size=32768, dtype=np.float32, device=0):
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.float32, np.float64):
            raise TypeError("dtype must be float32 or float64")
        self._h = C.c_void_p()
        _abi.check(_abi.lib().hb_spectral_create(C.byref(self._h), _abi.HB_F64 if self.dtype == np.float64 else _abi.HB_F32,
                                                 int(max_fft_size), int(device)))

    def close(self):
        if self._h:
            _abi.lib().hb_spectral_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_max_fft_size(self, size):
        _abi.check(_abi.lib().hb_spectral_set_max_fft_size(self._h, int(size)))

    def max_fft_size(self):
        return int(_abi.lib().hb_spectral_max_fft_size(self._h))

    def convolved_size(self, size1, size2, mode):
        return int(_abi.lib().hb_spectral_convolved_size(self._h, int(size1), int(size2), int(mode)))

    def convolve(self, output, in1, in2, mode=EdgeMode.Linear):
        """output[:convolved_size] = in1 * in2 under `mode`; returns the number of samples written
        (0: nothing done -- empty input or FFT above the maximum, as the reference)."""
        in1 = np.ascontiguousarray(in1, self.dtype)
        in2 = np.ascontiguousarray(in2, self.dtype)
        need = self.convolved_size(in1.size, in2.size, mode)
        if output.dtype != self.dtype or not output.flags["C_CONTIGUOUS"] or output.size < need:
            raise ValueError("output must be a contiguous %s array of at least %d samples" % (self.dtype, need))
        written = C.c_size_t(0)
        _abi.check(_abi.lib().hb_spectral_convolve(self._h, output.ctypes.data_as(C.c_void_p), in1.ctypes.data_as(C.c_void_p), in1.size,
                                                   in2.ctypes.data_as(C.c_void_p), in2.size, int(mode), C.byref(written)))
        return int(written.value)
