"""N x M Convolver sharded over the GPUs of one node: one process per GPU (torch.distributed).

The reference's matrix convolver sums, for every output, the convolutions of all inputs
(NToMonoConvolve.cpp:39-42).  Pairs are independent, so the INPUT channels are dealt to the ranks:
rank r holds the IR spectra of inputs [in_lo, in_hi) against ALL outputs (1/world of the spectra,
1/world of the HBM traffic), forward-transforms only its own inputs, and produces a partial block for
every output.  The one exchange step of the path is the sum of those partial blocks -- a
reduce-scatter over NVLink (NCCL) that leaves rank r owning outputs [out_lo, out_hi).  Parallel
(non-matrix) convolvers have no exchange at all: their channels are dealt out as independent banks.

This module is host logic only; the arithmetic is the CUDA engine behind the C ABI (_Matrix ->
hb_matrix).  `engine_factory` exists so that the CPU test-suite can exercise the sharding plan and
the collective under gloo with a stand-in engine; the default is the CUDA engine and nothing else.
"""
import numpy as np

from .convolve import _Matrix, _scheme
from .errors import ConvolveError

_ERR = ConvolveError


class ShardPlan:
    """Which rank owns which input channels and which output channels after the reduce-scatter."""

    def __init__(self, num_ins, num_outs, world, rank):
        if num_ins % world or num_outs % world:
            raise ValueError("numIns (%d) and numOuts (%d) must be multiples of the world size %d" % (num_ins, num_outs, world))
        self.num_ins, self.num_outs, self.world, self.rank = num_ins, num_outs, world, rank
        self.local_ins = num_ins // world
        self.local_outs = num_outs // world
        self.in_lo, self.in_hi = rank * self.local_ins, (rank + 1) * self.local_ins
        self.out_lo, self.out_hi = rank * self.local_outs, (rank + 1) * self.local_outs

    def input_owner(self, in_chan):
        return in_chan // self.local_ins

    def output_owner(self, out_chan):
        return out_chan // self.local_outs


class _CudaMatrixEngine:
    """default engine: the hb_matrix handle on this rank's GPU, fed with torch CUDA tensors"""

    def __init__(self, ins, outs, max_length, scheme, dtype, device):
        self.m = _Matrix(1, ins, outs, max_length, scheme, dtype, device)

    def set(self, i, o, ir, length, resize):
        return self.m.set(0, i, o, ir, length, resize)

    def set_reset_offset(self, offset):
        self.m.setResetOffset(offset)

    def reset(self):
        self.m.reset()

    def process_tensor(self, x, y, n, stream):
        """x [ins, >=n], y [outs, >=n] contiguous-row CUDA tensors; y is overwritten.  True when written."""
        return self.m.process_device(x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), n, False, stream)

    # the exchange fused into the inverse-FFT epilogue needs ONE uniform engine (no head, one FFT size)
    def can_fuse(self):
        return self.m.head_taps == 0 and len(self.m.engines) == 1

    def shard_export(self, world, rank):
        return self.m.tail.shard_export(world, rank)

    def shard_attach(self, handles):
        self.m.tail.shard_attach(handles)

    def process_shard_tensor(self, x, y_shard, n, stream):
        """x [local ins, >=n]; y_shard [outs/world, >=n] receives this rank's complete output rows."""
        from . import _abi
        return self.m.tail.process_shard_device(x.data_ptr(), x.stride(0), y_shard.data_ptr(), y_shard.stride(0), n, False, stream) == _abi.HB_OK

    def close(self):
        self.m.close()


class ShardedConvolver:
    """HISSTools::Convolver in N x M mode (Convolver.h:25-50) with the input channels sharded over the
    process group.  Construct it on every rank with the same arguments.

    set(inChan, outChan, ...) takes GLOBAL channel indices on every rank; only the owner stores the IR.
    process_device(x_local, y_shard, n): x_local holds this rank's input rows [local_ins, n], y_shard
    receives this rank's output rows [local_outs, n] (the full sum over all inputs)."""

    def __init__(self, numIns, numOuts, *scheme, maxLength=16384, dtype=np.float32, device=None, group=None, engine_factory=None,
                 exchange="auto"):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.plan = ShardPlan(int(numIns), int(numOuts), self.world, self.rank)
        self.dtype = np.dtype(dtype)
        if device is None:
            import torch
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        factory = engine_factory or _CudaMatrixEngine
        self.engine = factory(self.plan.local_ins, self.plan.num_outs, maxLength, _scheme(scheme), dtype, device)
        self._partial = None
        self.backend = dist.get_backend(group) if dist.is_initialized() else None
        # exchange: "fused" = partial blocks stored straight into the owner's memory by the inverse-FFT kernel
        # (peer-mapped inboxes over NVLink, hb_conv_shard_*); "nccl" = reduce-scatter of a local partial buffer;
        # "auto" = fused when the engine and the topology allow it
        self.exchange = "nccl"
        if exchange not in ("auto", "fused", "nccl"):
            raise ValueError("exchange must be auto, fused or nccl")
        if exchange != "nccl" and self.world > 1 and self.backend == "nccl" and getattr(self.engine, "can_fuse", lambda: False)():
            ok = 1
            try:
                handle = self.engine.shard_export(self.world, self.rank)
            except Exception:
                handle, ok = b"", 0
            handles = [None] * self.world
            dist.all_gather_object(handles, (ok, handle), group=group)
            if all(h[0] for h in handles):
                try:
                    self.engine.shard_attach([h[1] for h in handles])
                except Exception:
                    ok = 0
            else:
                ok = 0
            flags = [None] * self.world
            dist.all_gather_object(flags, ok, group=group)
            if all(flags):
                self.exchange = "fused"
            elif exchange == "fused":
                raise RuntimeError("fused exchange requested but peer memory could not be mapped on every rank")
        elif exchange == "fused":
            raise RuntimeError("fused exchange needs world > 1, the nccl backend and a single uniform partition size")

    def setResetOffset(self, offset=-1):
        self.engine.set_reset_offset(offset)

    def reset(self):
        self.engine.reset()

    def set(self, inChan, outChan, input, length=None, resize=False):
        """Range errors as Convolver::set (Convolver.cpp:114-124) on every rank; the owner of inChan loads the IR."""
        if outChan >= self.plan.num_outs or outChan < 0:
            return _ERR.CONVOLVE_ERR_OUT_CHAN_OUT_OF_RANGE
        if inChan >= self.plan.num_ins or inChan < 0:
            return _ERR.CONVOLVE_ERR_IN_CHAN_OUT_OF_RANGE
        if self.plan.input_owner(inChan) != self.rank:
            return _ERR.CONVOLVE_ERR_NONE
        if length is None:
            length = 0 if input is None else len(input)
        return self.engine.set(inChan - self.plan.in_lo, outChan, input, length, resize)

    def _partial_like(self, y_shard, n):
        import torch
        if self._partial is None or self._partial.shape[1] < n or self._partial.device != y_shard.device or self._partial.dtype != y_shard.dtype:
            self._partial = torch.zeros(self.plan.num_outs, n, dtype=y_shard.dtype, device=y_shard.device)
        return self._partial

    def process_device(self, x_local, y_shard, n, stream=0):
        """One block: local partial outputs for every output channel, then the sum over ranks.
        Returns True when y_shard was written (some rank had an IR loaded)."""
        dist = self.dist
        if not stream and hasattr(x_local, "is_cuda") and x_local.is_cuda:
            # 0 would mean "the engine's own stream" to the C ABI; the collective and the copies below run on torch's
            # current stream, so the engine's kernels must be enqueued there as well
            import torch
            stream = torch.cuda.current_stream(x_local.device).cuda_stream
        if self.exchange == "fused":
            return self.engine.process_shard_tensor(x_local, y_shard, n, stream)
        part = self._partial_like(y_shard, n)
        pv = part[:, :n] if part.shape[1] != n else part
        wrote = self.engine.process_tensor(x_local, part, n, stream)
        if self.world == 1:
            if wrote:
                y_shard[:, :n].copy_(pv)
            return wrote
        if not wrote:
            pv.zero_()                                      # a rank with nothing loaded contributes silence
        if self.backend == "nccl" and part.shape[1] == n and y_shard.is_contiguous() and y_shard.shape[1] == n:
            dist.reduce_scatter_tensor(y_shard, part, op=dist.ReduceOp.SUM, group=self.group)
        else:
            # gloo has no reduce-scatter: all-reduce, keep this rank's rows
            buf = pv.contiguous()
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            y_shard[:, :n].copy_(buf[self.plan.out_lo:self.plan.out_hi])
        return True

    def close(self):
        self.engine.close()
