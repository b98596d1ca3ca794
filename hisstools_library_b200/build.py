"""Build libhisstools_b200.so (hand-written sm_100a CUDA + the extern "C" boundary) in-tree with nvcc.

The library is compiled for sm_100a only; nvcc cross-compiles without a GPU.  The built file is
git-ignored but travels to the GPU box with the working tree.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBNAME = "libhisstools_b200.so"
SOURCES = ["hb_fft.cu", "hb_conv.cu", "hb_matrix.cu", "hb_spectral.cu", "hb_audio.cu"]
HEADERS = ["hb_common.cuh", "hb_fft_core.cuh", "hb_fft_block.cuh", "hb_fft_big.cuh", "hb_conv_kernels.cuh", "hb_conv_big.cuh",
           os.path.join("..", "..", "include", "hisstools_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "1886"]


def lib_path():
    return os.path.join(LIBDIR, LIBNAME)


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build " + LIBNAME)
    return exe


def needs_build():
    out = lib_path()
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".cu", ".h"))]       # every source and header in csrc/
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile the library if it is missing or older than its sources; returns its path."""
    if not force and not needs_build():
        return lib_path()
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", lib_path()] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building " + LIBNAME)
    if verbose:
        sys.stderr.write(res.stderr)
    return lib_path()


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
