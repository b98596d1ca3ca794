"""spectral_processor -- one-shot FFT convolution and correlation on the B200.

Mirror of the reference's spectral_processor<T> (SpectralProcessor.hpp:11-683): same class name, `EdgeMode`
values, `convolve` / `correlate` for real inputs (output, in1, in2, mode) and for complex inputs
(r_out, i_out, r_in1, i_in1, r_in2, i_in2, mode), `convolved_size` / `correlated_size`,
`set_max_fft_size` / `max_fft_size`, `change_phase`.  The transforms, the
per-bin product and the edge-mode arrangement run as CUDA kernels behind hb_spectral_* of
include/hisstools_b200.h.
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

    correlated_size = convolved_size        # SpectralProcessor.hpp:215-218

    def _check_out(self, out, need):
        if out.dtype != self.dtype or not out.flags["C_CONTIGUOUS"] or out.size < need:
            raise ValueError("output must be a contiguous %s array of at least %d samples" % (self.dtype, need))

    def _real(self, fn, output, in1, in2, mode):
        in1 = np.ascontiguousarray(in1, self.dtype)
        in2 = np.ascontiguousarray(in2, self.dtype)
        self._check_out(output, self.convolved_size(in1.size, in2.size, mode))
        written = C.c_size_t(0)
        _abi.check(fn(self._h, output.ctypes.data_as(C.c_void_p), in1.ctypes.data_as(C.c_void_p), in1.size,
                      in2.ctypes.data_as(C.c_void_p), in2.size, int(mode), C.byref(written)))
        return int(written.value)

    def _complex(self, fn, r_out, i_out, planes, mode):
        planes = [np.zeros(0, self.dtype) if p is None else np.ascontiguousarray(p, self.dtype) for p in planes]
        need = self.convolved_size(max(planes[0].size, planes[1].size), max(planes[2].size, planes[3].size), mode)
        self._check_out(r_out, need)
        self._check_out(i_out, need)
        written = C.c_size_t(0)
        args = []
        for p in planes:
            args += [p.ctypes.data_as(C.c_void_p) if p.size else None, p.size]
        _abi.check(fn(self._h, r_out.ctypes.data_as(C.c_void_p), i_out.ctypes.data_as(C.c_void_p), *args, int(mode), C.byref(written)))
        return int(written.value)

    def convolve(self, *args):
        """convolve(output, in1, in2, mode) for real inputs or convolve(r_out, i_out, r_in1, i_in1, r_in2, i_in2, mode) for
        complex ones (a missing plane: None), SpectralProcessor.hpp:164-172.  Returns the number of samples written
        (0: nothing done -- empty input or FFT above the maximum, as the reference)."""
        if len(args) in (3, 4):
            return self._real(_abi.lib().hb_spectral_convolve, args[0], args[1], args[2], args[3] if len(args) == 4 else EdgeMode.Linear)
        return self._complex(_abi.lib().hb_spectral_convolve_complex, args[0], args[1], args[2:6], args[6])

    def correlate(self, *args):
        """correlate(output, in1, in2, mode) / correlate(r_out, i_out, r_in1, i_in1, r_in2, i_in2, mode), SpectralProcessor.hpp:176-184."""
        if len(args) in (3, 4):
            return self._real(_abi.lib().hb_spectral_correlate, args[0], args[1], args[2], args[3] if len(args) == 4 else EdgeMode.Linear)
        return self._complex(_abi.lib().hb_spectral_correlate_complex, args[0], args[1], args[2:6], args[6])

    def change_phase(self, output, input, size, phase, time_multiplier=1.0):
        """change_phase (SpectralProcessor.hpp:186-208): output[:fft_size] = the input with its phase moved towards minimum
        (phase 0), linear (0.5) or maximum (1) phase; returns the number of samples written (the FFT size)."""
        x = np.ascontiguousarray(input, self.dtype)
        if output.dtype != self.dtype or not output.flags["C_CONTIGUOUS"]:
            raise ValueError("output must be a contiguous %s array" % self.dtype)
        written = C.c_size_t(0)
        _abi.check(_abi.lib().hb_spectral_change_phase(self._h, output.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), int(size),
                                                       float(phase), float(time_multiplier), C.byref(written)))
        return int(written.value)
