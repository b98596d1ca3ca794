"""Host-side mirror of the reference convolver classes over the CUDA engine.

Same class names, method names, argument meaning and error codes as
HIRT_Multichannel_Convolution/{PartitionedConvolve,MonoConvolve,NToMonoConvolve,Convolver}.h; every
`process` runs on the GPU through the C ABI (include/hisstools_b200.h) -- there is no CPU path.

Where the reference builds an N x M matrix out of N*M independent MonoConvolve objects
(Convolver.cpp:5-22, NToMonoConvolve.cpp:4-9), this mirror gives every part of the partition scheme
ONE engine that holds the whole matrix: inputs are transformed once, the sum over inputs and
partitions happens in the frequency domain and each output is inverse-transformed once.

Superset over the reference: `dtype=np.float64` selects a true double engine (the
reference classes are float-only, PartitionedConvolve.h:38-41), NToMonoConvolve / Convolver accept
the custom partition sizes MonoConvolve accepts (MonoConvolve.h:31), and `process_device` takes
device-resident buffers.
"""
import ctypes as C

import numpy as np

from . import _abi
from .errors import ConvolveError, LatencyMode

_ERR = ConvolveError


def _hb_dtype(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return _abi.HB_F32
    if dtype == np.float64:
        return _abi.HB_F64
    raise TypeError("dtype must be float32 or float64")


class _Engine:
    """One hb_conv handle: `groups` banks of an ins x outs matrix, uniform partitions."""

    def __init__(self, dtype, groups, ins, outs, max_fft, max_length, offset, length, device=0, borrowed=None):
        self.dtype = np.dtype(dtype)
        self.groups, self.ins, self.outs = int(groups), int(ins), int(outs)
        self.device = device
        self._owned = borrowed is None
        if borrowed is not None:
            self._h = C.c_void_p(borrowed)          # a part of an hb_matrix: the matrix owns it
            self.ctor_error = 0
            return
        self._h = C.c_void_p()
        code = _abi.lib().hb_conv_create(C.byref(self._h), _hb_dtype(dtype), self.groups, self.ins, self.outs,
                                         int(max_fft), int(max_length), int(offset), int(length), int(device))
        self.ctor_error = _abi.check(code)

    def close(self):
        if self._h and self._owned:
            _abi.lib().hb_conv_destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # setters: reference ConvolveError codes come back as ints
    def set_fft_size(self, n):
        return _abi.check(_abi.lib().hb_conv_set_fft_size(self._h, int(n)))

    def set_length(self, n):
        return _abi.check(_abi.lib().hb_conv_set_length(self._h, int(n)))

    def set_offset(self, n):
        return _abi.check(_abi.lib().hb_conv_set_offset(self._h, int(n)))

    def set_reset_offset(self, n):
        return _abi.check(_abi.lib().hb_conv_set_reset_offset(self._h, int(n)))

    def set_ir(self, group, i, o, ir, length=None):
        if ir is None or (length is not None and length == 0):
            return _abi.check(_abi.lib().hb_conv_set_ir(self._h, group, i, o, None, _abi.HB_F32, 0))
        ir = np.ascontiguousarray(ir)
        if ir.dtype not in (np.float32, np.float64):
            ir = ir.astype(self.dtype)
        n = ir.size if length is None else int(length)
        if n > ir.size:
            raise ValueError("length exceeds the impulse response array")
        return _abi.check(_abi.lib().hb_conv_set_ir(self._h, group, i, o, ir.ctypes.data_as(C.c_void_p), _hb_dtype(ir.dtype), n))

    def set_ir_live(self, group, i, o, ir, length=None):
        """set_ir on a running engine that restarts this pair only (hb_conv_set_ir_live)"""
        if ir is None or (length is not None and length == 0):
            return _abi.check(_abi.lib().hb_conv_set_ir_live(self._h, group, i, o, None, _abi.HB_F32, 0))
        ir = np.ascontiguousarray(ir)
        if ir.dtype not in (np.float32, np.float64):
            ir = ir.astype(self.dtype)
        n = ir.size if length is None else int(length)
        return _abi.check(_abi.lib().hb_conv_set_ir_live(self._h, group, i, o, ir.ctypes.data_as(C.c_void_p), _hb_dtype(ir.dtype), n))

    def reset_pair(self, group, i, o):
        return _abi.check(_abi.lib().hb_conv_reset_pair(self._h, group, i, o))

    def set_ir_device(self, group, i, o, data_ptr, length):
        return _abi.check(_abi.lib().hb_conv_set_ir_dev(self._h, group, i, o, C.c_void_p(data_ptr), int(length)))

    def resize(self, max_length):
        return _abi.check(_abi.lib().hb_conv_resize(self._h, int(max_length)))

    def reset(self):
        return _abi.check(_abi.lib().hb_conv_reset(self._h))

    @property
    def partitions(self):
        return int(_abi.lib().hb_conv_partitions(self._h))

    @property
    def max_length(self):
        return int(_abi.lib().hb_conv_max_length(self._h))

    @property
    def fft_size(self):
        return int(_abi.lib().hb_conv_fft_size(self._h))

    @property
    def bytes_per_hop(self):
        return int(_abi.lib().hb_conv_bytes_per_hop(self._h))

    def set_tuning(self, ctas_per_sm=0, variant=1):
        return _abi.check(_abi.lib().hb_conv_set_tuning(self._h, int(ctas_per_sm), int(variant)))

    def set_host_pipeline(self, pipelined=True):
        return _abi.check(_abi.lib().hb_conv_set_host_pipeline(self._h, 1 if pipelined else 0))

    # fused multi-GPU exchange (hb_conv_shard_*)
    def shard_export(self, world, rank):
        buf = C.create_string_buffer(64)
        _abi.check(_abi.lib().hb_conv_shard_export(self._h, int(world), int(rank), buf))
        return buf.raw

    def shard_attach(self, handles):
        blob = b"".join(handles)
        _abi.check(_abi.lib().hb_conv_shard_attach(self._h, C.c_char_p(blob)))

    def process_shard_device(self, in_ptr, in_ld, out_ptr, out_ld, n, accumulate=False, stream=0):
        return _abi.check(_abi.lib().hb_conv_process_shard_dev(self._h, C.c_void_p(in_ptr), int(in_ld), C.c_void_p(out_ptr), int(out_ld),
                                                               int(n), 1 if accumulate else 0, C.c_void_p(stream)))

    def shard_status(self):
        """bit mask of the ranks whose blocks the owner-side sum stopped waiting for (0 = none was ever late)"""
        late = C.c_uint32()
        _abi.check(_abi.lib().hb_conv_shard_status(self._h, C.byref(late)))
        return int(late.value)

    def join(self, stream=0):
        """make `stream` wait for the tail the overlapped schedule launched ahead (hb_conv_join)"""
        return _abi.check(_abi.lib().hb_conv_join(self._h, C.c_void_p(stream)))

    def set_profiling(self, enable=True):
        return _abi.check(_abi.lib().hb_conv_set_profiling(self._h, 1 if enable else 0))

    def get_profile(self):
        """({forward, cmac, wait, inverse, tail} summed ms, hops) since profiling was enabled; `cmac` is the whole
        multiply-accumulate in the serial schedule and the head (partition 0) in the overlapped one."""
        ms, h = (C.c_double * 5)(), C.c_uint64()
        _abi.check(_abi.lib().hb_conv_get_profile(self._h, ms, C.byref(h)))
        return dict(zip(("forward", "cmac", "wait", "inverse", "tail"), [float(v) for v in ms])), int(h.value)

    def set_fft_path(self, path=0):
        """0 automatic, 1 one CTA per transform, 2 cluster of 8 CTAs, 3 four-step (hb_conv_set_fft_path)."""
        return _abi.check(_abi.lib().hb_conv_set_fft_path(self._h, int(path)))

    @property
    def fft_path(self):
        return _abi.lib().hb_conv_fft_path(self._h)

    def set_multi_hop(self, enable=True):
        return _abi.check(_abi.lib().hb_conv_set_multi_hop(self._h, 1 if enable else 0))

    def set_trace(self, enable=True):
        return _abi.check(_abi.lib().hb_conv_set_trace(self._h, 1 if enable else 0))

    def get_trace(self):
        """(stamps[16 hops][5 kinds][entry/exit][256 CTAs] in ns (0 = not run), hops processed so far)."""
        buf, hop = (C.c_uint64 * (16 * 5 * 2 * 256))(), C.c_uint64()
        _abi.check(_abi.lib().hb_conv_get_trace(self._h, buf, C.byref(hop)))
        return np.frombuffer(buf, dtype=np.uint64).reshape(16, 5, 2, 256).copy(), int(hop.value)

    def set_schedule(self, overlapped=True):
        """True: overlapped (tail of the next hop computed ahead on a second stream), False: serial, None: automatic
        (the fused single-launch hop where it applies, else overlapped / serial by size)."""
        return _abi.check(_abi.lib().hb_conv_set_schedule(self._h, 2 if overlapped is None else (1 if overlapped else 0)))

    def set_hop_overlap(self, mode=1):
        """fused hops: 0 never overlap consecutive calls, 1 on the engine's own stream, 2 on any stream (hb_conv_set_hop_overlap)"""
        return _abi.check(_abi.lib().hb_conv_set_hop_overlap(self._h, int(mode)))

    def set_tail_streams(self, streams=1):
        """overlapped schedule: 2 = tails of consecutive hops on alternating streams (hb_conv_set_tail_streams)"""
        return _abi.check(_abi.lib().hb_conv_set_tail_streams(self._h, int(streams)))

    @property
    def tail_streams(self):
        return _abi.lib().hb_conv_tail_streams(self._h)

    @property
    def schedule(self):
        return ("serial", "overlapped", "fused")[_abi.lib().hb_conv_schedule(self._h)]

    @property
    def bytes_per_launch(self):
        return int(_abi.lib().hb_conv_bytes_per_launch(self._h))

    def process(self, in_rows, out_rows, n, accumulate=False):
        """in_rows / out_rows: lists of contiguous 1-D arrays of the engine dtype (>= n samples); None = silent input / unwanted output."""
        ip = (C.c_void_p * len(in_rows))(*[None if r is None else r.ctypes.data for r in in_rows])
        op = (C.c_void_p * len(out_rows))(*[None if r is None else r.ctypes.data for r in out_rows])
        return _abi.check(_abi.lib().hb_conv_process(self._h, ip, op, int(n), 1 if accumulate else 0))

    @staticmethod
    def row_pointers(rows):
        """the `const void *const *` a host-pointer call takes, built once for rows that are reused (numpy's .ctypes costs about a
        microsecond per row -- as much as the call itself for a small engine)"""
        return (C.c_void_p * len(rows))(*[None if r is None else r.ctypes.data for r in rows])

    def process_pointers(self, in_ptrs, out_ptrs, n, accumulate=False):
        """process() on arrays made by row_pointers()"""
        return _abi.check(_abi.lib().hb_conv_process(self._h, in_ptrs, out_ptrs, int(n), 1 if accumulate else 0))

    def process_device(self, in_ptr, in_ld, out_ptr, out_ld, n, accumulate=False, stream=0):
        return _abi.check(_abi.lib().hb_conv_process_dev(self._h, C.c_void_p(in_ptr), int(in_ld), C.c_void_p(out_ptr), int(out_ld),
                                                         int(n), 1 if accumulate else 0, C.c_void_p(stream)))


def _rows(arrays, count, n, dtype, writable=False):
    """Normalise `arrays` (2-D array or sequence of 1-D arrays) to `count` contiguous rows of dtype.
    Returns (rows, writeback) where writeback lists (temp, target) pairs to copy after processing."""
    rows, back = [], []
    for k in range(count):
        a = arrays[k]
        if not isinstance(a, np.ndarray):
            raise TypeError("audio rows must be numpy arrays")
        if a.ndim != 1 or a.size < n:
            raise ValueError("audio row %d must be 1-D with at least %d samples" % (k, n))
        if a.dtype == dtype and a.flags["C_CONTIGUOUS"]:
            rows.append(a)
        else:
            t = np.ascontiguousarray(a[:n], dtype=dtype)
            rows.append(t)
            if writable:
                back.append((t, a))
    return rows, back


# ---------------------------------------------------------------------------------------------------
class PartitionedConvolve:
    """HISSTools::PartitionedConvolve (PartitionedConvolve.h:23-41): uniform partitioned convolution of
    one channel; output is the linear convolution delayed by FFTSize/2 samples."""

    def __init__(self, maxFFTSize=16384, maxLength=131072, offset=0, length=0, dtype=np.float32, device=0):
        self._e = _Engine(dtype, 1, 1, 1, maxFFTSize, maxLength, offset, length, device)
        self.dtype = self._e.dtype

    def setFFTSize(self, FFTSize):
        return _ERR(self._e.set_fft_size(FFTSize))

    def setLength(self, length):
        return _ERR(self._e.set_length(length))

    def setOffset(self, offset):
        self._e.set_offset(offset)

    def setResetOffset(self, offset=-1):
        self._e.set_reset_offset(offset)

    def set(self, input, length=None):
        return _ERR(self._e.set_ir(0, 0, 0, input, length))

    def reset(self):
        self._e.reset()

    def process(self, in_, out, numSamples):
        """Returns False (and leaves `out` untouched) when no IR is loaded (PartitionedConvolve.cpp:262-263)."""
        n = int(numSamples)
        rin, _ = _rows([in_], 1, n, self.dtype)
        rout, back = _rows([out], 1, n, self.dtype, writable=True)
        code = self._e.process(rin, rout, n)
        if code == _abi.HB_ERR_NO_IR:
            return False
        for t, a in back:
            a[:n] = t
        return True

    @property
    def engine(self):
        return self._e


# ---------------------------------------------------------------------------------------------------
def partition_scheme(zeroLatency, A, B=0, C=0, D=0):
    """The part map of MonoConvolve::setPartitions (MonoConvolve.cpp:203-258).

    Returns (sizes, head_taps, fixed, tail) with fixed = [(fft, offset, taps)...] and
    tail = (fft, offset); head_taps is the length of the time-domain head (0 without zero latency)."""
    sizes = []
    prev = 0
    for s in (A, B, C, D):
        if (1 << 5) <= s <= (1 << 20) and s > prev:
            sizes.append(int(s))
            prev = s
        elif s:
            raise RuntimeError("invalid FFT size or order")
    if not sizes:
        raise RuntimeError("no valid FFT sizes given")
    n = len(sizes)
    offset = sizes[0] >> 1 if zeroLatency else 0
    head = offset
    fixed = []

    def part(size, nxt):
        nonlocal offset
        taps = (nxt - size) >> 1
        fixed.append((size, offset, taps))
        offset += taps

    if n == 4:
        part(sizes[0], sizes[1])
    if n > 2:
        part(sizes[n - 3], sizes[n - 2])
    if n > 1:
        part(sizes[n - 2], sizes[n - 1])
    return sizes, head, fixed, (sizes[-1], offset)


_LATENCY_SIZES = {
    LatencyMode.kLatencyZero: (True, 256, 1024, 4096, 16384),       # MonoConvolve.cpp:28
    LatencyMode.kLatencyShort: (False, 256, 1024, 4096, 16384),     # :29
    LatencyMode.kLatencyMedium: (False, 1024, 4096, 16384, 0),      # :30
}


class _Matrix:
    """groups x (ins x outs) convolution matrix with one partition scheme: one hb_matrix handle, the
    shared machinery behind MonoConvolve, NToMonoConvolve and Convolver (host logic in csrc/hb_matrix.cu)."""

    def __init__(self, groups, ins, outs, maxLength, scheme, dtype, device, devices=None):
        zero, A, B, C_, D = scheme
        self.dtype = np.dtype(dtype)
        self.groups, self.ins, self.outs = groups, ins, outs
        self.sizes, _, _, _ = partition_scheme(zero, A, B, C_, D)       # raises RuntimeError like the reference
        self._h = C.c_void_p()
        self.devices = None if devices is None else [int(d) for d in devices]
        if self.devices is not None:
            # one matrix dealt to several GPUs of this process (hb_matrix_create_multi)
            arr = (C.c_int * len(self.devices))(*self.devices)
            code = _abi.lib().hb_matrix_create_multi(C.byref(self._h), _hb_dtype(dtype), groups, ins, outs, int(maxLength),
                                                     1 if zero else 0, int(A), int(B), int(C_), int(D), arr, len(self.devices))
        else:
            code = _abi.lib().hb_matrix_create(C.byref(self._h), _hb_dtype(dtype), groups, ins, outs, int(maxLength),
                                               1 if zero else 0, int(A), int(B), int(C_), int(D), int(device))
        _abi.check(code)
        lib = _abi.lib()
        self.head_taps = int(lib.hb_matrix_head_taps(self._h))
        self.engines = [_Engine(dtype, groups, ins, outs, 0, 0, 0, 0, device, borrowed=lib.hb_matrix_part(self._h, k))
                        for k in range(lib.hb_matrix_parts(self._h))]
        self.tail = self.engines[-1]

    @property
    def exchange(self):
        """how the sum over inputs crosses devices: none, peer-reads (owner-side kernel) or fused (inverse-FFT epilogue)"""
        return ("none", "peer-reads", "fused")[_abi.lib().hb_matrix_exchange(self._h)]

    def shard_engines(self, part=-1):
        """one borrowed _Engine per device: part `part` (default: the tail) of every shard of a multi-device matrix"""
        lib = _abi.lib()
        out = []
        for d in range(lib.hb_matrix_shards(self._h)):
            sh = C.c_void_p(lib.hb_matrix_shard(self._h, d))
            k = lib.hb_matrix_parts(sh) - 1 if part < 0 else part
            dev = self.devices[d] if self.devices is not None else 0
            out.append(_Engine(self.dtype, 1, 1, 1, 0, 0, 0, 0, dev, borrowed=lib.hb_matrix_part(sh, k)))
        return out

    def setResetOffset(self, offset=-1):
        _abi.check(_abi.lib().hb_matrix_set_reset_offset(self._h, int(offset)))

    def resize(self, g, i, o, length):
        """MonoConvolve::resize (MonoConvolve.cpp:100-110): drops the pair's IR, sets its allocation."""
        return _ERR(_abi.check(_abi.lib().hb_matrix_resize(self._h, g, i, o, int(length))))

    def set(self, g, i, o, ir, length, requestResize):
        """MonoConvolve::set (MonoConvolve.cpp:118-140)."""
        if ir is None:
            return _ERR(_abi.check(_abi.lib().hb_matrix_set(self._h, g, i, o, None, _abi.HB_F32, 0, 1 if requestResize else 0)))
        ir = np.ascontiguousarray(ir)
        if ir.dtype not in (np.float32, np.float64):
            ir = ir.astype(self.dtype)
        if int(length) > ir.size:
            raise ValueError("length exceeds the impulse response array")
        return _ERR(_abi.check(_abi.lib().hb_matrix_set(self._h, g, i, o, ir.ctypes.data_as(C.c_void_p), _hb_dtype(ir.dtype),
                                                        int(length), 1 if requestResize else 0)))

    def reset(self):
        _abi.check(_abi.lib().hb_matrix_reset(self._h))
        return _ERR.CONVOLVE_ERR_NONE

    def reset_pair(self, g, i, o):
        """Convolver::reset(inChan, outChan) (Convolver.cpp:88-97): that pair restarts from silence, the others keep running"""
        return _ERR(_abi.check(_abi.lib().hb_matrix_reset_pair(self._h, g, i, o)))

    def process(self, in_rows, out_rows, n, accumulate):
        """Sum of all parts into out_rows (None rows: silent input / unwanted output); True when written."""
        ip = (C.c_void_p * len(in_rows))(*[None if r is None else r.ctypes.data for r in in_rows])
        op = (C.c_void_p * len(out_rows))(*[None if r is None else r.ctypes.data for r in out_rows])
        return _abi.check(_abi.lib().hb_matrix_process(self._h, ip, op, int(n), 1 if accumulate else 0)) == _abi.HB_OK

    def set_hop_overlap(self, mode=1):
        """hb_matrix_set_hop_overlap: fused parts may run consecutive device calls side by side (0 never, 1 own stream, 2 any)"""
        return _abi.check(_abi.lib().hb_matrix_set_hop_overlap(self._h, int(mode)))

    def process_device(self, in_ptr, in_ld, out_ptr, out_ld, n, accumulate=False, stream=0):
        code = _abi.lib().hb_matrix_process_dev(self._h, C.c_void_p(in_ptr), int(in_ld), C.c_void_p(out_ptr), int(out_ld),
                                                int(n), 1 if accumulate else 0, C.c_void_p(stream))
        return _abi.check(code) == _abi.HB_OK

    def close(self):
        if self._h:
            for e in self.engines:
                e.close()
            _abi.lib().hb_matrix_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _scheme(args):
    """(latency) or (zeroLatency, A[, B, C, D]) -> the 5-tuple setPartitions takes."""
    if len(args) == 1 and isinstance(args[0], (LatencyMode, int)) and not isinstance(args[0], bool):
        return _LATENCY_SIZES[LatencyMode(args[0])]
    if not args:
        raise TypeError("a LatencyMode or (zeroLatency, A, B=0, C=0, D=0) is required")
    zero = bool(args[0])
    sizes = list(args[1:]) + [0] * (5 - len(args))
    if len(sizes) != 4:
        raise TypeError("expected (zeroLatency, A, B=0, C=0, D=0)")
    return (zero,) + tuple(int(s) for s in sizes)


class MonoConvolve:
    """HISSTools::MonoConvolve (MonoConvolve.h:30-48).
    MonoConvolve(maxLength, latency) or MonoConvolve(maxLength, zeroLatency, A, B=0, C=0, D=0)."""

    def __init__(self, maxLength, *scheme, dtype=np.float32, device=0):
        self._m = _Matrix(1, 1, 1, maxLength, _scheme(scheme), dtype, device)
        self.dtype = self._m.dtype

    def setResetOffset(self, offset=-1):
        self._m.setResetOffset(offset)

    def resize(self, length):
        return self._m.resize(0, 0, 0, length)

    def set(self, input, length=None, requestResize=False):
        if length is None:
            length = 0 if input is None else len(input)
        return self._m.set(0, 0, 0, input, length, requestResize)

    def reset(self):
        return self._m.reset()

    def process(self, in_, temp, out, numSamples, accumulate=False):
        """`temp` is accepted for signature compatibility (MonoConvolve.h:46) and not used."""
        n = int(numSamples)
        rin, _ = _rows([in_], 1, n, self.dtype)
        rout, back = _rows([out], 1, n, self.dtype, writable=True)
        if self._m.process(rin, rout, n, accumulate):
            for t, a in back:
                a[:n] = t

    @property
    def matrix(self):
        return self._m


class NToMonoConvolve:
    """HISSTools::NToMonoConvolve (NToMonoConvolve.h:18-24): N inputs summed into one output."""

    def __init__(self, inChans, maxLength, *scheme, dtype=np.float32, device=0):
        self.mNumInChans = int(inChans)
        self._m = _Matrix(1, self.mNumInChans, 1, maxLength, _scheme(scheme), dtype, device)
        self.dtype = self._m.dtype

    def _chan(self, inChan):
        return inChan < self.mNumInChans

    def resize(self, inChan, impulse_length):
        if not self._chan(inChan):
            return _ERR.CONVOLVE_ERR_IN_CHAN_OUT_OF_RANGE
        return self._m.resize(0, inChan, 0, impulse_length)

    def set(self, inChan, input, impulse_length=None, resize=False):
        if not self._chan(inChan):
            return _ERR.CONVOLVE_ERR_IN_CHAN_OUT_OF_RANGE
        if impulse_length is None:
            impulse_length = 0 if input is None else len(input)
        return self._m.set(0, inChan, 0, input, impulse_length, resize)

    def reset(self, inChan=None):
        if inChan is None:
            return self._m.reset()
        if not self._chan(inChan):
            return _ERR.CONVOLVE_ERR_IN_CHAN_OUT_OF_RANGE
        return self._m.reset_pair(0, inChan, 0)             # that input's convolver only (NToMonoConvolve.cpp:28-33)

    def setResetOffset(self, offset=-1):
        self._m.setResetOffset(offset)

    def process(self, ins, out, temp, numSamples, activeInChans):
        """out = sum over the first activeInChans inputs (NToMonoConvolve.cpp:35-43); inputs beyond
        activeInChans are fed silence."""
        n = int(numSamples)
        active = min(int(activeInChans), self.mNumInChans)
        rin, _ = _rows(ins, active, n, self.dtype)
        rin = rin + [None] * (self.mNumInChans - active)
        rout, back = _rows([out], 1, n, self.dtype, writable=True)
        rout[0][:n] = 0
        self._m.process(rin, rout, n, True)
        for t, a in back:
            a[:n] = t

    @property
    def matrix(self):
        return self._m


class Convolver:
    """HISSTools::Convolver (Convolver.h:25-50).
    Convolver(numIns, numOuts, latency) -- N x M matrix; Convolver(numIO, latency) -- numIO parallel
    channels.  Custom partitions: pass (zeroLatency, A, B, C, D) instead of a LatencyMode.
    The reference constructs every pair with room for 16384 taps (Convolver.cpp:18,35); use
    set(..., resize=True) for longer IRs, or the maxLength keyword to pre-allocate."""

    def __init__(self, *args, dtype=np.float32, device=0, maxLength=16384, devices=None):
        """devices=[...]: the matrix is dealt to these GPUs of this process (input channels of an N x M matrix, banks of a
        parallel convolver); everything else -- set, resize, process with all rows -- stays as on one device."""
        args = list(args)
        if len(args) >= 2 and isinstance(args[1], (int, np.integer)) and not isinstance(args[1], (LatencyMode, bool)):
            self.mN2M = True
            self.mNumIns = max(int(args[0]), 1)
            self.mNumOuts = int(args[1])
            scheme = args[2:]
            self._m = _Matrix(1, self.mNumIns, self.mNumOuts, maxLength, _scheme(scheme), dtype, device, devices)
        else:
            self.mN2M = False
            self.mNumIns = self.mNumOuts = max(int(args[0]), 1)
            scheme = args[1:]
            self._m = _Matrix(self.mNumIns, 1, 1, maxLength, _scheme(scheme), dtype, device, devices)
        self.dtype = self._m.dtype

    def _pair(self, inChan, outChan):
        """(group, in, out) of the engine, or an error code (Convolver.cpp:86-124)."""
        if not self.mN2M:
            inChan -= outChan                 # parallel mode: callers pass the same channel twice
        if outChan >= self.mNumOuts or outChan < 0:
            return None, _ERR.CONVOLVE_ERR_OUT_CHAN_OUT_OF_RANGE
        limit = self.mNumIns if self.mN2M else 1
        if inChan < 0 or inChan >= limit:
            return None, _ERR.CONVOLVE_ERR_IN_CHAN_OUT_OF_RANGE
        return ((0, inChan, outChan) if self.mN2M else (outChan, 0, 0)), _ERR.CONVOLVE_ERR_NONE

    def clear(self, *args):
        """clear(resize) or clear(inChan, outChan, resize) (Convolver.cpp:51-71)."""
        if len(args) == 1:
            if self.mN2M:
                for o in range(self.mNumOuts):
                    for i in range(self.mNumIns):
                        self.set(i, o, None, 0, args[0])
            else:
                for o in range(self.mNumOuts):
                    self.set(o, o, None, 0, args[0])
            return
        inChan, outChan, resize = args
        self.set(inChan, outChan, None, 0, resize)

    def reset(self, inChan=None, outChan=None):
        if inChan is None:
            self._m.reset()
            return None
        pair, err = self._pair(inChan, outChan)
        if err:
            return err
        return self._m.reset_pair(pair[0], pair[1], pair[2])

    def resize(self, inChan, outChan, length):
        pair, err = self._pair(inChan, outChan)
        if err:
            # the reference reports a bad OUTPUT channel as IN_CHAN here (Convolver.cpp:109)
            return _ERR.CONVOLVE_ERR_IN_CHAN_OUT_OF_RANGE
        return self._m.resize(pair[0], pair[1], pair[2], length)

    def set(self, inChan, outChan, input, length=None, resize=False):
        """float or double IRs (Convolver.cpp:114-134); a double IR is rounded to the engine dtype."""
        pair, err = self._pair(inChan, outChan)
        if err:
            return err
        if length is None:
            length = 0 if input is None else len(input)
        return self._m.set(pair[0], pair[1], pair[2], input, length, resize)

    def setResetOffset(self, offset=-1):
        self._m.setResetOffset(offset)

    def process(self, ins, outs, numIns, numOuts, numSamples):
        """float or double audio rows (Convolver.cpp:138-183).  Outputs below numOuts are always
        written (silence when nothing is loaded); inputs beyond numIns are fed silence."""
        n = int(numSamples)
        numIns = min(int(numIns), self.mNumIns)
        numOuts = min(int(numOuts), self.mNumOuts)
        rin, _ = _rows(ins, numIns, n, self.dtype)
        rin = rin + [None] * (self.mNumIns - numIns)
        rout, back = _rows(outs, numOuts, n, self.dtype, writable=True)
        for r in rout:
            r[:n] = 0
        rout = rout + [None] * (self.mNumOuts - numOuts)
        self._m.process(rin, rout, n, True)
        for t, a in back:
            a[:n] = t

    def process_device(self, in_ptr, in_ld, out_ptr, out_ld, numSamples, stream=0):
        """Device-resident rows of the engine dtype: in [numIns][in_ld], out [numOuts][out_ld]."""
        return self._m.process_device(in_ptr, in_ld, out_ptr, out_ld, numSamples, False, stream)

    def set_hop_overlap(self, mode=1):
        """may consecutive process_device calls of one block run side by side on the GPU?  0 never, 1 on the object's own stream
        (stream=0), 2 on any stream -- the rows of a call are complete when it is made (hb_matrix_set_hop_overlap)"""
        return self._m.set_hop_overlap(mode)

    @property
    def matrix(self):
        return self._m
