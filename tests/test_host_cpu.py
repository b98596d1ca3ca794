"""CPU-side checks of the product's host logic (no GPU needed):
 * the C-ABI library builds, loads and exports every symbol include/hisstools_b200.h declares, and the
   ctypes table matches the header one to one;
 * without a CUDA device every compute entry point fails loudly (no CPU fallback exists);
 * the index arithmetic of the shared-memory FFT and the stream-K unit decomposition, by running the
   very same __host__ __device__ functions on the CPU (tests/host_emul) against the oracle;
 * the partition map of MonoConvolve::setPartitions (MonoConvolve.cpp:203-258).
"""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import checkers as ck

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_gpu():
    import torch
    return not torch.cuda.is_available()


def _header_functions():
    text = open(os.path.join(ROOT, "include", "hisstools_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from hisstools_library_b200 import _abi
    lib = _abi.lib()
    names = _header_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "declared in the header but not exported: " + n
    assert sorted(_abi.SIGNATURES) == names, set(_abi.SIGNATURES) ^ set(names)
    assert b"sm_100a" in lib.hb_version()


def test_library_holds_sm100a_code_only():
    from hisstools_library_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(not _no_gpu(), reason="a CUDA device is present")
def test_no_device_means_loud_failure_not_fallback():
    import hisstools_library_b200 as hb
    from hisstools_library_b200 import _abi
    with pytest.raises(hb.HissError) as e:
        hb.PartitionedConvolve(1024, 4096, 0, 0)
    assert e.value.code == _abi.HB_ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(hb.HissError):
        hb.hisstools_create_setup(10)
    with pytest.raises(hb.HissError):
        hb.Convolver(2, 2, hb.kLatencyShort)
    # invalid partition sizes are a host-side error, as in the reference (MonoConvolve.cpp:212,229)
    with pytest.raises(RuntimeError):
        hb.MonoConvolve(1000, False, 1024, 256)


def test_partition_scheme_matches_reference_map():
    from hisstools_library_b200 import partition_scheme
    # Zero = TD[0,128) + 256:[128,512) + 1024:[512,2048) + 4096:[2048,8192) + 16384:[8192,inf)
    sizes, head, fixed, tail = partition_scheme(True, 256, 1024, 4096, 16384)
    assert head == 128 and fixed == [(256, 128, 384), (1024, 512, 1536), (4096, 2048, 6144)] and tail == (16384, 8192)
    sizes, head, fixed, tail = partition_scheme(False, 256, 1024, 4096, 16384)
    assert head == 0 and fixed == [(256, 0, 384), (1024, 384, 1536), (4096, 1920, 6144)] and tail == (16384, 8064)
    sizes, head, fixed, tail = partition_scheme(False, 1024, 4096, 16384, 0)
    assert fixed == [(1024, 0, 1536), (4096, 1536, 6144)] and tail == (16384, 7680)
    assert partition_scheme(False, 8192) == ([8192], 0, [], (8192, 0))
    for bad in [(False, 16), (False, 1024, 1024), (False, 1 << 21), (False, 0)]:
        with pytest.raises(RuntimeError):
            partition_scheme(*bad)


# ---- CPU emulation of the device index arithmetic -------------------------------------------------

@pytest.fixture(scope="module")
def emul():
    d = tempfile.mkdtemp(prefix="hb_emul_")
    so = os.path.join(d, "libemul.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-o", so, os.path.join(ROOT, "tests", "host_emul", "emul_fft.cpp")], check=True)
    lib = C.CDLL(so)
    for suf, P in (("_f32", ck.c_f32p), ("_f64", ck.c_f64p)):
        getattr(lib, "emul_cfft" + suf).argtypes = [P, P, C.c_int, C.c_int, C.c_uint32]
        getattr(lib, "emul_rfft" + suf).argtypes = [P, P, C.c_int, C.c_int, C.c_int, C.c_uint32]
    for n in ("emul_unit_begin", "emul_unit_owner"):
        getattr(lib, n).restype = C.c_uint64
        getattr(lib, n).argtypes = [C.c_uint64] * 3
    return lib


def _threads(log2m, ept):
    t = max((1 << log2m) // ept, 32)
    return (t + 31) & ~31


@pytest.mark.parametrize("suf,tol", [("_f32", 1e-5), ("_f64", 1e-12)])
def test_emulated_device_fft_matches_oracle(emul, suf, tol):
    dt = np.float32 if suf == "_f32" else np.float64
    lib = ck.oracle()
    rng = np.random.default_rng(3)
    for log2n in range(1, 15):
        ept = 16 if (1 << log2n) // 8 > 1024 else 8
        # complex transform of 2^log2n points
        re, im = rng.uniform(-1, 1, 1 << log2n).astype(dt), rng.uniform(-1, 1, 1 << log2n).astype(dt)
        wr, wi = re.copy(), im.copy()
        s = getattr(lib, "orc_fft_setup_create" + suf)(max(log2n, 4))
        getattr(lib, "orc_fft" + suf)(s, ck.fptr(wr), ck.fptr(wi), log2n)
        getattr(emul, "emul_cfft" + suf)(ck.fptr(re), ck.fptr(im), log2n, ept, _threads(log2n, ept))
        assert ck.rel_rms(np.stack([re, im]), np.stack([wr, wi])) <= tol, ("fft", log2n)
        # real transforms of 2^log2n points (planes of half that)
        if log2n >= 2:
            half = 1 << (log2n - 1)
            for inverse, op in ((0, "orc_rfft"), (1, "orc_rifft")):
                re, im = rng.uniform(-1, 1, half).astype(dt), rng.uniform(-1, 1, half).astype(dt)
                wr, wi = re.copy(), im.copy()
                getattr(lib, op + suf)(s, ck.fptr(wr), ck.fptr(wi), log2n)
                e2 = 16 if half // 8 > 1024 else 8
                getattr(emul, "emul_rfft" + suf)(ck.fptr(re), ck.fptr(im), log2n, inverse, e2, _threads(log2n - 1, e2))
                assert ck.rel_rms(np.stack([re, im]), np.stack([wr, wi])) <= tol, (op, log2n)
        getattr(lib, "orc_fft_setup_destroy" + suf)(s)


def test_stream_k_unit_ranges_partition_the_work(emul):
    """every unit belongs to exactly one CTA, ranges are contiguous, and unit_owner inverts unit_begin."""
    for U, G in [(262144, 148), (64, 64), (64, 148), (4096, 148), (1, 1), (513, 148), (8192, 296), (7, 3)]:
        G = min(G, U)
        begins = [emul.emul_unit_begin(g, U, G) for g in range(G + 1)]
        assert begins[0] == 0 and begins[-1] == U
        assert all(b1 >= b0 for b0, b1 in zip(begins, begins[1:]))
        sizes = [b1 - b0 for b0, b1 in zip(begins, begins[1:])]
        assert max(sizes) - min(sizes) <= 1
        probe = sorted(set([0, U - 1] + [b for b in begins[:-1]] + [max(b - 1, 0) for b in begins[1:]]))
        for u in probe:
            g = emul.emul_unit_owner(u, U, G)
            assert begins[g] <= u < begins[g + 1], (U, G, u, g)
