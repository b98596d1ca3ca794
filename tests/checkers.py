"""ctypes windows onto the two CHECKERS used by the test-suite (test infrastructure only):

* ``ref``    -- oracle/_ref/libhisstools_ref.so: the unmodified reference compiled in place
               (oracle/Makefile, oracle/ref_shim.cpp).  Present in this container and shipped
               prebuilt to the GPU box; tests that need it skip when it is absent.
* ``oracle`` -- oracle/libhiss_oracle.so: our plain-C restatement (oracle/hiss_oracle.c).

Nothing in the product package imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)
SZ = C.c_size_t


def fptr(a):
    """numpy array -> typed ctypes pointer (array must stay alive)."""
    if a.dtype == np.float32:
        return a.ctypes.data_as(c_f32p)
    if a.dtype == np.float64:
        return a.ctypes.data_as(c_f64p)
    raise TypeError(a.dtype)


def rel_rms(got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    den = np.sqrt(np.sum(want * want))
    num = np.sqrt(np.sum((got - want) ** 2))
    return num / den if den > 0 else num


def build_checkers():
    """(Re)build the checkers; `make ref` keeps prebuilt files when /root/reference is absent."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True, stdout=subprocess.DEVNULL)


def _load(path):
    return C.CDLL(path) if os.path.exists(path) else None


_oracle = None
_ref = None
_ref_spec = None


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "libhiss_oracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)
        lib = C.CDLL(path)
        for suf, P in (("_f32", c_f32p), ("_f64", c_f64p)):
            def sig(name, res, *args):
                fn = getattr(lib, name + suf)
                fn.restype = res
                fn.argtypes = list(args)
            V = C.c_void_p
            sig("orc_fft_setup_create", V, C.c_uint)
            sig("orc_fft_setup_destroy", None, V)
            for n in ("orc_fft", "orc_ifft", "orc_rfft", "orc_rifft"):
                sig(n, None, V, P, P, C.c_uint)
            sig("orc_unzip", None, P, P, P, C.c_uint)
            sig("orc_zip", None, P, P, P, C.c_uint)
            sig("orc_unzip_zero", None, P, P, P, SZ, C.c_uint)
            sig("orc_rfft_real", None, V, P, P, P, SZ, C.c_uint)
            sig("orc_rifft_real", None, V, P, P, P, C.c_uint)
            sig("orc_pconv_create", V, SZ, SZ, SZ, SZ)
            sig("orc_pconv_destroy", None, V)
            sig("orc_pconv_set_fft_size", C.c_int, V, SZ)
            sig("orc_pconv_set_length", C.c_int, V, SZ)
            sig("orc_pconv_set_offset", None, V, SZ)
            sig("orc_pconv_set_reset_offset", None, V, C.c_long)
            sig("orc_pconv_set", C.c_int, V, P, SZ)
            sig("orc_pconv_reset", None, V)
            sig("orc_pconv_process", C.c_int, V, P, P, SZ)
            sig("orc_tdconv_create", V, SZ, SZ)
            sig("orc_tdconv_destroy", None, V)
            sig("orc_tdconv_set", C.c_int, V, P, SZ)
            sig("orc_tdconv_process", C.c_int, V, P, P, SZ)
            sig("orc_mono_create", V, SZ, C.c_int, SZ, SZ, SZ, SZ)
            sig("orc_mono_destroy", None, V)
            sig("orc_mono_resize", C.c_int, V, SZ)
            sig("orc_mono_set", C.c_int, V, P, SZ, C.c_int)
            sig("orc_mono_reset", None, V)
            sig("orc_mono_process", None, V, P, P, SZ, C.c_int)
            sig("orc_spectral_convolve", SZ, P, P, SZ, P, SZ, C.c_int, SZ)
            sig("orc_spectral_binary", SZ, P, P, SZ, P, SZ, C.c_int, C.c_int, SZ)
            sig("orc_spectral_change_phase", SZ, P, P, SZ, C.c_double, C.c_double)
            sig("orc_spectral_binary_complex", SZ, P, P, P, SZ, P, SZ, P, SZ, P, SZ, C.c_int, C.c_int, SZ)
        _oracle = lib
    return _oracle


def ref():
    """The compiled reference, or None when oracle/_ref is absent."""
    global _ref
    if _ref is None:
        lib = _load(os.path.join(ORACLE_DIR, "_ref", "libhisstools_ref.so"))
        if lib is None:
            return None
        V = C.c_void_p
        UP = C.c_size_t
        IP = C.c_ssize_t
        PP32 = C.POINTER(c_f32p)
        PP64 = C.POINTER(c_f64p)

        def sig(name, res, *args):
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = list(args)

        for suf, P in (("_f32", c_f32p), ("_f64", c_f64p)):
            sig("ref_fft_setup" + suf, V, UP)
            sig("ref_fft_setup_free" + suf, None, V)
            for n in ("fft", "ifft", "rfft", "rifft"):
                sig("ref_%s%s" % (n, suf), None, V, P, P, UP)
            sig("ref_rfft_real" + suf, None, V, P, P, P, UP, UP)
            sig("ref_rifft_real" + suf, None, V, P, P, P, UP)
            sig("ref_unzip" + suf, None, P, P, P, UP)
            sig("ref_zip" + suf, None, P, P, P, UP)
            sig("ref_unzip_zero" + suf, None, P, P, P, UP, UP)
            sig("ref_restated_create" + suf, V, UP)
            sig("ref_restated_destroy" + suf, None, V)
            sig("ref_restated_set" + suf, None, V, P, UP)
            sig("ref_restated_process" + suf, C.c_int, V, P, P, UP)
        sig("ref_rfft_real_f32_f64", None, V, c_f32p, c_f64p, c_f64p, UP, UP)
        sig("ref_unzip_zero_f32_f64", None, c_f32p, c_f64p, c_f64p, UP, UP)
        sig("ref_pconv_create", V, UP, UP, UP, UP)
        sig("ref_pconv_destroy", None, V)
        sig("ref_pconv_set_fft_size", C.c_int, V, UP)
        sig("ref_pconv_set_length", C.c_int, V, UP)
        sig("ref_pconv_set_offset", None, V, UP)
        sig("ref_pconv_set_reset_offset", None, V, IP)
        sig("ref_pconv_set", C.c_int, V, c_f32p, UP)
        sig("ref_pconv_reset", None, V)
        sig("ref_pconv_process", C.c_int, V, c_f32p, c_f32p, UP)
        sig("ref_mono_create_latency", V, UP, C.c_int)
        sig("ref_mono_create_custom", V, UP, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32)
        sig("ref_mono_destroy", None, V)
        sig("ref_mono_set_reset_offset", None, V, IP)
        sig("ref_mono_resize", C.c_int, V, UP)
        sig("ref_mono_set", C.c_int, V, c_f32p, UP, C.c_int)
        sig("ref_mono_reset", C.c_int, V)
        sig("ref_mono_process", None, V, c_f32p, c_f32p, c_f32p, UP, C.c_int)
        sig("ref_n2m_create", V, C.c_uint32, UP, C.c_int)
        sig("ref_n2m_destroy", None, V)
        sig("ref_n2m_resize", C.c_int, V, C.c_uint32, UP)
        sig("ref_n2m_set", C.c_int, V, C.c_uint32, c_f32p, UP, C.c_int)
        sig("ref_n2m_reset", C.c_int, V, C.c_uint32)
        sig("ref_n2m_process", None, V, PP32, c_f32p, c_f32p, SZ, SZ)
        sig("ref_conv_create", V, C.c_uint32, C.c_uint32, C.c_int)
        sig("ref_conv_create_parallel", V, C.c_uint32, C.c_int)
        sig("ref_conv_destroy", None, V)
        sig("ref_conv_clear", None, V, C.c_int)
        sig("ref_conv_clear_chan", None, V, C.c_uint32, C.c_uint32, C.c_int)
        sig("ref_conv_reset", None, V)
        sig("ref_conv_reset_chan", C.c_int, V, C.c_uint32, C.c_uint32)
        sig("ref_conv_resize", C.c_int, V, C.c_uint32, C.c_uint32, UP)
        sig("ref_conv_set_f32", C.c_int, V, C.c_uint32, C.c_uint32, c_f32p, UP, C.c_int)
        sig("ref_conv_set_f64", C.c_int, V, C.c_uint32, C.c_uint32, c_f64p, UP, C.c_int)
        sig("ref_conv_process_f32", None, V, PP32, PP32, SZ, SZ, SZ)
        sig("ref_conv_process_f64", None, V, PP64, PP64, SZ, SZ, SZ)
        sig("ref_matrix_create", V, C.c_uint32, C.c_uint32, UP, C.c_uint32, C.c_int)
        sig("ref_matrix_destroy", None, V)
        sig("ref_matrix_set", C.c_int, V, C.c_uint32, C.c_uint32, c_f32p, UP)
        sig("ref_matrix_process", None, V, PP32, PP32, SZ)
        sig("ref_matrix_reset_pair", C.c_int, V, C.c_uint32, C.c_uint32)
        sig("ref_matrix_set_pool_mt", C.c_int, V, PP32, C.c_uint32, UP, C.c_int)
        sig("ref_matrix_process_mt", None, V, PP32, PP32, SZ, C.c_int)
        sig("ref_matrix_time", C.c_double, V, PP32, PP32, SZ, C.c_int, C.c_int, C.c_int)
        sig("ref_hardware_threads", C.c_int)
        sig("ref_restated_time_f64", C.c_double, C.POINTER(V), C.c_int, PP64, PP64, SZ, C.c_int, C.c_int, C.c_int)
        _ref = lib
    return _ref


def ref_spectral():
    global _ref_spec
    if _ref_spec is None:
        lib = _load(os.path.join(ORACLE_DIR, "_ref", "libhisstools_ref_spectral.so"))
        if lib is None:
            return None
        lib.ref_spectral_convolve_f32.restype = SZ
        lib.ref_spectral_convolve_f32.argtypes = [c_f32p, c_f32p, SZ, c_f32p, SZ, C.c_int, SZ]
        lib.ref_spectral_convolve_f64.restype = SZ
        lib.ref_spectral_convolve_f64.argtypes = [c_f64p, c_f64p, SZ, c_f64p, SZ, C.c_int, SZ]
        for suf, P in (("_f32", c_f32p), ("_f64", c_f64p)):
            fn = getattr(lib, "ref_spectral_binary" + suf)
            fn.restype = SZ
            fn.argtypes = [P, P, SZ, P, SZ, C.c_int, C.c_int, SZ]
            fn = getattr(lib, "ref_spectral_binary_complex" + suf)
            fn.restype = SZ
            fn.argtypes = [P, P, P, SZ, P, SZ, P, SZ, P, SZ, C.c_int, C.c_int, SZ]
            fn = getattr(lib, "ref_spectral_change_phase" + suf)
            fn.restype = SZ
            fn.argtypes = [P, P, SZ, C.c_double, C.c_double, SZ]
        _ref_spec = lib
    return _ref_spec


class RefAudioInfo(C.Structure):
    """ref_audio_info of oracle/ref_audio_shim.cpp"""
    _fields_ = [("file_type", C.c_int32), ("pcm_format", C.c_int32), ("header_big_endian", C.c_int32), ("audio_big_endian", C.c_int32),
                ("channels", C.c_uint32), ("frames", C.c_uint32), ("sampling_rate", C.c_double), ("error_flags", C.c_int32), ("is_open", C.c_int32)]


_ref_audio = None


def ref_audio():
    """The reference's audio-file reader / writer (oracle/_ref/libhisstools_ref_audio.so), or None."""
    global _ref_audio
    if _ref_audio is None:
        lib = _load(os.path.join(ORACLE_DIR, "_ref", "libhisstools_ref_audio.so"))
        if lib is None:
            return None
        lib.ref_audio_write.restype = C.c_int
        lib.ref_audio_write.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, c_f64p, C.c_uint32]
        lib.ref_audio_probe.restype = None
        lib.ref_audio_probe.argtypes = [C.c_char_p, C.POINTER(RefAudioInfo)]
        lib.ref_audio_read_f32.restype = C.c_int
        lib.ref_audio_read_f32.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, c_f32p]
        lib.ref_audio_read_f64.restype = C.c_int
        lib.ref_audio_read_f64.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, c_f64p]
        if hasattr(lib, "ref_oaudio_open"):
            lib.ref_oaudio_open.restype = C.c_void_p
            lib.ref_oaudio_open.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
            lib.ref_oaudio_write_f64.restype = None
            lib.ref_oaudio_write_f64.argtypes = [C.c_void_p, c_f64p, C.c_uint32, C.c_int]
            lib.ref_oaudio_write_f32.restype = None
            lib.ref_oaudio_write_f32.argtypes = [C.c_void_p, c_f32p, C.c_uint32, C.c_int]
            lib.ref_oaudio_write_raw.restype = None
            lib.ref_oaudio_write_raw.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32]
            lib.ref_oaudio_seek.restype = None
            lib.ref_oaudio_seek.argtypes = [C.c_void_p, C.c_uint32]
            for name in ("ref_oaudio_position", "ref_oaudio_frames"):
                getattr(lib, name).restype = C.c_uint32
                getattr(lib, name).argtypes = [C.c_void_p]
            for name in ("ref_oaudio_flags", "ref_oaudio_is_open", "ref_oaudio_file_type"):
                getattr(lib, name).restype = C.c_int
                getattr(lib, name).argtypes = [C.c_void_p]
            lib.ref_oaudio_close.restype = None
            lib.ref_oaudio_close.argtypes = [C.c_void_p]
            lib.ref_audio_read_raw.restype = C.c_int
            lib.ref_audio_read_raw.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_void_p]
        _ref_audio = lib
    return _ref_audio


def planar_ptrs(arr2d):
    """2-D C-contiguous float array -> (ctypes array of row pointers)."""
    assert arr2d.flags["C_CONTIGUOUS"]
    P = c_f32p if arr2d.dtype == np.float32 else c_f64p
    rows = (P * arr2d.shape[0])()
    for i in range(arr2d.shape[0]):
        rows[i] = arr2d[i].ctypes.data_as(P)
    return rows


SUF = {np.dtype(np.float32): "_f32", np.dtype(np.float64): "_f64"}


# ---- small python conveniences over the checkers ------------------------------------------------

def ref_pconv_run(fft_size, ir, x, block, max_len=None, offset=0, length=0, reset_offset=0):
    """Stream x through the reference PartitionedConvolve in `block`-sample calls."""
    lib = ref()
    ir = np.ascontiguousarray(ir, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    h = lib.ref_pconv_create(fft_size, max_len if max_len is not None else len(ir), offset, length)
    lib.ref_pconv_set_reset_offset(h, reset_offset)
    err = lib.ref_pconv_set(h, fptr(ir), len(ir))
    y = np.zeros_like(x)
    pos = 0
    while pos < len(x):
        n = min(block, len(x) - pos)
        lib.ref_pconv_process(h, fptr(x[pos:]), fptr(y[pos:]), n)
        pos += n
    lib.ref_pconv_destroy(h)
    return y, err


def oracle_pconv_run(fft_size, ir, x, block, max_len=None, offset=0, length=0, reset_offset=0, dtype=np.float32):
    lib = oracle()
    suf = SUF[np.dtype(dtype)]
    ir = np.ascontiguousarray(ir, dtype)
    x = np.ascontiguousarray(x, dtype)
    g = lambda n: getattr(lib, n + suf)
    h = g("orc_pconv_create")(fft_size, max_len if max_len is not None else len(ir), offset, length)
    g("orc_pconv_set_reset_offset")(h, reset_offset)
    err = g("orc_pconv_set")(h, fptr(ir), len(ir))
    y = np.zeros_like(x)
    pos = 0
    while pos < len(x):
        n = min(block, len(x) - pos)
        g("orc_pconv_process")(h, fptr(x[pos:]), fptr(y[pos:]), n)
        pos += n
    g("orc_pconv_destroy")(h)
    return y, err


def direct_convolve_delayed(ir, x, delay):
    """float64 ground truth: out[n] = sum_k h[k] x[n - delay - k]."""
    full = np.convolve(np.asarray(x, np.float64), np.asarray(ir, np.float64))
    out = np.zeros(len(x))
    if delay < len(x):
        out[delay:] = full[: len(x) - delay]
    return out


def direct_convolve_delayed_fft(ir, x, delay):
    """the same ground truth for long inputs: float64 FFT convolution (numpy), exact to ~1e-15 relative."""
    x = np.asarray(x, np.float64)
    ir = np.asarray(ir, np.float64)
    n = 1
    while n < len(x) + len(ir):
        n *= 2
    full = np.fft.irfft(np.fft.rfft(x, n) * np.fft.rfft(ir, n), n)[: len(x) + len(ir) - 1]
    out = np.zeros(len(x))
    if delay < len(x):
        out[delay:] = full[: len(x) - delay]
    return out


def synth_audio(n, channel=0):
    """white noise uniform[-1,1), seed 1000+channel."""
    return np.random.default_rng(1000 + channel).uniform(-1.0, 1.0, n).astype(np.float32)


def synth_ir(length, pair=0):
    """N(0,1) * exp(-6.9 k / L), seed 2000+pair."""
    rng = np.random.default_rng(2000 + pair)
    k = np.arange(length, dtype=np.float64)
    return (rng.standard_normal(length) * np.exp(-6.9 * k / max(length, 1))).astype(np.float32)


def ref_matrix_run(irs, xs, fft_size, block=None, threads=None):
    """The reference's uniform-partition matrix (rows of MonoConvolve(maxLen, false, fft) summed as
    NToMonoConvolve.cpp:35-43 does) on irs[rows][ins][taps] and xs[ins][n], streamed in `block`-sample calls
    (default: one hop) with the rows dealt to host threads.  Returns y[rows][n] (float32)."""
    lib = ref()
    irs = np.ascontiguousarray(irs, np.float32)
    xs = np.ascontiguousarray(xs, np.float32)
    rows, ins, taps = irs.shape
    n = xs.shape[1]
    block = block or fft_size // 2
    m = lib.ref_matrix_create(ins, rows, taps, fft_size, 0)
    if not m:
        raise MemoryError("reference matrix allocation failed")
    try:
        for o in range(rows):
            for i in range(ins):
                lib.ref_matrix_set(m, i, o, fptr(irs[o, i]), taps)
        y = np.zeros((rows, n), np.float32)
        use = threads or min(rows, os.cpu_count() or 1)
        P32 = c_f32p
        for pos in range(0, n, block):
            nb = min(block, n - pos)
            xp = (P32 * ins)(*[xs[i, pos:].ctypes.data_as(P32) for i in range(ins)])
            yp = (P32 * rows)(*[y[o, pos:].ctypes.data_as(P32) for o in range(rows)])
            lib.ref_matrix_process_mt(m, xp, yp, nb, use)
    finally:
        lib.ref_matrix_destroy(m)
    return y


def ref_restated_run_f64(ir, x, fft_size, block=None):
    """The double-precision oracle of SURVEY 8c (PartitionedConvolve.cpp:173-426 restated over the reference's own
    FFT_SETUP_D transforms, oracle/ref_shim.cpp PConvRestated<double>) on one channel."""
    lib = ref()
    ir = np.ascontiguousarray(ir, np.float64)
    x = np.ascontiguousarray(x, np.float64)
    block = block or fft_size // 2
    h = lib.ref_restated_create_f64(fft_size)
    lib.ref_restated_set_f64(h, fptr(ir), len(ir))
    y = np.zeros_like(x)
    for pos in range(0, len(x), block):
        lib.ref_restated_process_f64(h, fptr(x[pos:]), fptr(y[pos:]), min(block, len(x) - pos))
    lib.ref_restated_destroy_f64(h)
    return y
