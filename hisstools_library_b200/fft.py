"""hisstools_fft / ifft / rfft / rifft family on the B200.

Mirror of the reference FFT interface (HISSTools_FFT/HISSTools_FFT.h:87-369): same names, argument
order and conventions -- split-complex planes, forward kernel exp(-j theta), the real forward
transform returns 2*DFT with DC in realp[0] and Nyquist in imagp[0], nothing is scaled
(rifft(rfft(x)) = 2N x).  Planes are numpy arrays transformed IN PLACE, as the reference transforms
the caller's memory.  Every transform runs as a CUDA kernel behind the C ABI; zip / unzip are pure
re-orderings of host memory (Core:1185-1287) and stay on the host.
"""
import ctypes as C

import numpy as np

from . import _abi


class Split:
    """FFT_SPLIT_COMPLEX_F / _D (HISSTools_FFT.h:26-34): two planes of one dtype."""

    def __init__(self, realp, imagp):
        if realp.dtype != imagp.dtype or realp.dtype not in (np.float32, np.float64):
            raise TypeError("planes must both be float32 or both float64")
        if not (realp.flags["C_CONTIGUOUS"] and imagp.flags["C_CONTIGUOUS"]):
            raise ValueError("planes must be contiguous")
        self.realp = realp
        self.imagp = imagp

    @classmethod
    def zeros(cls, n, dtype=np.float32):
        return cls(np.zeros(n, dtype), np.zeros(n, dtype))

    @property
    def dtype(self):
        return self.realp.dtype


class Setup:
    """FFT_SETUP_F / FFT_SETUP_D (HISSTools_FFT.h:57,63): an opaque twiddle-table owner."""

    def __init__(self, max_fft_log_2, dtype=np.float32, device=0):
        self.dtype = np.dtype(dtype)
        self._h = C.c_void_p()
        code = _abi.lib().hb_fft_setup_create(C.byref(self._h), _abi.HB_F64 if self.dtype == np.float64 else _abi.HB_F32,
                                              int(max_fft_log_2), int(device))
        _abi.check(code)
        self.max_fft_log_2 = int(max_fft_log_2)

    def destroy(self):
        if self._h:
            _abi.lib().hb_fft_setup_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def hisstools_create_setup(max_fft_log_2, dtype=np.float32, device=0):
    """hisstools_create_setup(FFT_SETUP_F/D *setup, max_fft_log_2) -- HISSTools_FFT.h:87,98."""
    return Setup(max_fft_log_2, dtype, device)


def hisstools_destroy_setup(setup):
    """HISSTools_FFT.h:108,118."""
    setup.destroy()


def _planes(setup, split, n):
    if split.dtype != setup.dtype:
        raise TypeError("setup is %s but the planes are %s" % (setup.dtype, split.dtype))
    if split.realp.size < n or split.imagp.size < n:
        raise ValueError("planes hold fewer than %d points" % n)
    return split.realp.ctypes.data_as(C.c_void_p), split.imagp.ctypes.data_as(C.c_void_p)


def hisstools_fft(setup, split, log2n):
    """In-place complex FFT of 2^log2n points (HISSTools_FFT.h:130,142)."""
    re, im = _planes(setup, split, 1 << log2n)
    _abi.check(_abi.lib().hb_fft(setup._h, re, im, log2n))


def hisstools_ifft(setup, split, log2n):
    """In-place unscaled inverse complex FFT (HISSTools_FFT.h:220,232)."""
    re, im = _planes(setup, split, 1 << log2n)
    _abi.check(_abi.lib().hb_ifft(setup._h, re, im, log2n))


def hisstools_rfft(setup, *args):
    """hisstools_rfft(setup, split, log2n): in place on unzipped planes (HISSTools_FFT.h:154,166), or
    hisstools_rfft(setup, input, split, in_length, log2n): out of place from a real array with zero
    padding (HISSTools_FFT.h:180,194,208; float input with a double setup is the :208 overload)."""
    if len(args) == 2:
        split, log2n = args
        if log2n < 1:
            return
        re, im = _planes(setup, split, 1 << (log2n - 1))
        _abi.check(_abi.lib().hb_rfft(setup._h, re, im, log2n))
        return
    inp, split, in_length, log2n = args
    inp = np.ascontiguousarray(inp)
    if inp.dtype not in (np.float32, np.float64):
        raise TypeError("input must be float32 or float64")
    if inp.size < min(in_length, 1 << log2n):
        raise ValueError("input shorter than in_length")
    re, im = _planes(setup, split, 1 << (log2n - 1))
    _abi.check(_abi.lib().hb_rfft_real(setup._h, inp.ctypes.data_as(C.c_void_p), _abi.HB_F64 if inp.dtype == np.float64 else _abi.HB_F32,
                                       re, im, int(in_length), log2n))


def hisstools_rifft(setup, split, *args):
    """hisstools_rifft(setup, split, log2n): in place (HISSTools_FFT.h:244,256), or
    hisstools_rifft(setup, split, output, log2n): also zips the result into `output` (:269,282)."""
    if len(args) == 1:
        (log2n,) = args
        if log2n < 1:
            return
        re, im = _planes(setup, split, 1 << (log2n - 1))
        _abi.check(_abi.lib().hb_rifft(setup._h, re, im, log2n))
        return
    output, log2n = args
    if output.dtype != setup.dtype or not output.flags["C_CONTIGUOUS"] or output.size < (1 << log2n):
        raise ValueError("output must be a contiguous array of the setup's dtype with 2^log2n elements")
    re, im = _planes(setup, split, 1 << (log2n - 1))
    _abi.check(_abi.lib().hb_rifft_real(setup._h, re, im, output.ctypes.data_as(C.c_void_p), log2n))


def hisstools_unzip(inp, split, log2n):
    """Even samples -> realp, odd -> imagp (HISSTools_FFT.h:333,345; Core:1185-1224)."""
    half = (1 << log2n) >> 1
    inp = np.asarray(inp)
    split.realp[:half] = inp[0:2 * half:2]
    split.imagp[:half] = inp[1:2 * half:2]


def hisstools_zip(split, output, log2n):
    """Interleave the planes back into a real array (HISSTools_FFT.h:357,369; Core:1228-1254)."""
    half = (1 << log2n) >> 1
    output[0:2 * half:2] = split.realp[:half]
    output[1:2 * half:2] = split.imagp[:half]


def hisstools_unzip_zero(inp, split, in_length, log2n):
    """Unzip with zero padding; in_length is clamped to the FFT size and an odd last sample lands in
    realp (HISSTools_FFT.h:295,308,321; Core:1258-1287)."""
    n = 1 << log2n
    half = n >> 1
    inp = np.asarray(inp)
    length = min(int(in_length), n)
    pairs = length >> 1
    split.realp[:half] = 0
    split.imagp[:half] = 0
    split.realp[:pairs] = inp[0:2 * pairs:2]
    split.imagp[:pairs] = inp[1:2 * pairs:2]
    if length & 1:
        split.realp[pairs] = inp[length - 1]
