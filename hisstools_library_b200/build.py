"""Build libhisstools_b200.so (hand-written sm_100a CUDA + the extern "C" boundary) in-tree with nvcc.

The library is compiled for sm_100a only; nvcc cross-compiles without a GPU.  The built file is
git-ignored but travels to the GPU box with the working tree.

Every translation unit is compiled to an object file of its own (in parallel) and the objects are linked
into a temporary file that replaces the library atomically, all under a file lock: several ranks importing
the package at once (torchrun, mp.spawn) with a missing or stale library neither run nvcc on the same output
nor load a half-written file.
"""
import concurrent.futures
import fcntl
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIBNAME = "libhisstools_b200.so"
SOURCES = ["hb_fft.cu", "hb_conv.cu", "hb_conv_mh.cu", "hb_matrix.cu", "hb_spectral.cu", "hb_audio.cu", "hb_audio_out.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-diag-suppress", "1886"]


def lib_path():
    return os.path.join(LIBDIR, LIBNAME)


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build " + LIBNAME)
    return exe


def _deps():
    """every source and header the library is made of (a change in any header rebuilds every object)"""
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".cu", ".h"))]
    deps += [os.path.join(HERE, "..", "include", "hisstools_b200.h"), os.path.abspath(__file__)]
    return deps


def needs_build():
    out = lib_path()
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in _deps())


def _compile(src, obj, verbose):
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, src)]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    return src, res


def build(force=False, verbose=False):
    """Compile the library if it is missing or older than its sources; returns its path."""
    if not force and not needs_build():
        return lib_path()
    os.makedirs(OBJDIR, exist_ok=True)
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():              # another process built it while this one waited
                return lib_path()
            headers_t = max(os.path.getmtime(d) for d in _deps() if not d.endswith(".cu"))
            todo = []
            for src in SOURCES:
                obj = os.path.join(OBJDIR, src[:-3] + ".o")
                stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(headers_t, os.path.getmtime(os.path.join(CSRC, src)))
                if stale:
                    todo.append((src, obj))
            with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as pool:
                for src, res in pool.map(lambda a: _compile(a[0], a[1], verbose), todo):
                    if res.returncode != 0:
                        sys.stderr.write(res.stdout + res.stderr)
                        raise RuntimeError("nvcc failed compiling " + src)
                    if verbose:
                        sys.stderr.write(res.stderr)
            tmp = lib_path() + ".tmp.%d" % os.getpid()
            objs = [os.path.join(OBJDIR, s[:-3] + ".o") for s in SOURCES]
            res = subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp] + objs,
                                 capture_output=True, text=True)
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                raise RuntimeError("nvcc failed linking " + LIBNAME)
            os.replace(tmp, lib_path())
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return lib_path()


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
