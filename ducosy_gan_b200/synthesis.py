"""Volume-level dual-HU synthesis: the arithmetic of the reference's ``generate()`` + ``synthesis()``
(generate.py:89-102 and generate.py:213-237) without the per-slice DICOM round trip.

    raw stored values [S,H,W] int16
        -> HU window (soft-tissue, lung)         preprocess.py:72-84      (fused into the stem im2col kernel)
        -> soft-tissue Generator, lung Generator  model.py:92-115          (two CUDA streams, batched slices)
        -> de-window + complementary composite    preprocess.py:96-111, generate.py:218-237   (one kernel)
        -> merged stored values [S,H,W] int16
        -> (optional) z / unsharp volume smoothing generate.py:254-263                        (postprocess.py, on device)

Slices are independent, so a volume shards across ranks by contiguous slice ranges with no collective
(``shard_range``).  DICOM I/O stays with the caller (out of scope).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ops
from .modules.model import Generator

SOFT_HU = (-150.0, 250.0)   # modules/argmanager.py:121-136
LUNG_HU = (-1000.0, -150.0)  # modules/argmanager.py:138-152


def chunk_size(num_slices: int, batch_slices: int) -> int:
    """Slices per generator call for a volume (or shard) of ``num_slices``: the whole thing at once while it is at most 1.5x
    ``batch_slices`` (a 37/38-slice shard of a 300-slice volume over 8 GPUs runs as ONE batch instead of 30 + a thin tail),
    otherwise equal chunks of at most ``batch_slices`` (33 -> 17 + 16, 300 -> 10 x 30)."""
    if num_slices <= 0:
        return 1
    if 2 * num_slices <= 3 * batch_slices:
        return num_slices
    n_chunks = -(-num_slices // batch_slices)
    return -(-num_slices // n_chunks)


def shard_range(num_slices: int, rank: int, world_size: int):
    """Contiguous slice range [lo, hi) of ``rank``; ranges are disjoint, ordered and cover [0, S)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return num_slices * rank // world_size, num_slices * (rank + 1) // world_size


class DualHUSynthesizer:
    """Both A2B generators + composite for whole volumes on one GPU.

    soft_model / lung_model: ``ducosy_gan_b200.modules.model.Generator(input_channels=1)`` instances holding the
    checkpoints (as generate.py:29-49 builds them).  ``batch_slices`` slices go through each generator at once
    (30 is where the throughput curve flattens on a B200: 974 slices/s vs 900 at 15 and 830 at 10; about 6.4 GB of workspace
    per generator).  ``precision``: operand mode of both generators (``modules.model.inference_operand_dtype``).
    """

    def __init__(self, soft_model: Generator, lung_model: Generator, soft_hu=SOFT_HU, lung_hu=LUNG_HU,
                 batch_slices: int = 30, device=None, precision: str | None = None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.soft_model, self.lung_model = soft_model, lung_model
        if precision is not None:          # "fp16" (default) | "bf16" | "fp16x2" (the <= 1 HU split-operand arm, ~3x slower)
            soft_model.precision = lung_model.precision = precision
        self.soft_hu, self.lung_hu = tuple(map(float, soft_hu)), tuple(map(float, lung_hu))
        self.batch_slices = int(batch_slices)
        self._streams = None
        self._copy_streams = None
        self._bufs = {}

    # ------------------------------------------------------------------ helpers
    def _engines(self):
        es = self.soft_model._engine(self.device)
        el = self.lung_model._engine(self.device)
        es.sync_weights(self.soft_model._ordered_params())
        el.sync_weights(self.lung_model._ordered_params())
        return es, el

    def _buffers(self, B, H, W):
        """fp32 generator outputs of one chunk; a pair that is large enough is reused (shards / volumes of other sizes)."""
        have = self._bufs.get((H, W))
        if have is None or have[0].shape[0] < B:
            self._bufs.clear()
            mk = lambda: torch.empty((B, 1, H, W), dtype=torch.float32, device=self.device)
            have = self._bufs[(H, W)] = (mk(), mk())
        return have[0][:B], have[1][:B]

    def launches_per_chunk(self):
        import ctypes as C
        from . import _lib
        es, el = self._engines()
        lib = _lib.load()
        return lib.ducosy_generator_num_launches(C.byref(es.cfg)) + lib.ducosy_generator_num_launches(C.byref(el.cfg)) + 1

    # ------------------------------------------------------------------ device-resident volume
    def synthesize_device(self, raw_px: torch.Tensor, slope=1.0, intercept=-1024.0, out: torch.Tensor | None = None,
                          postprocess: bool = False, group=None, before_chunk=None, after_chunk=None):
        """raw_px: int16 [S,H,W] on this GPU -> merged int16 [S,H,W] (same device).  Asynchronous.
        ``postprocess=True`` appends the volume smoothing of generate.py:254-263 on the device (``postprocess.py``); when
        the volume is sharded over the ranks of ``group`` this is where the path has its one exchange (z halo)."""
        if raw_px.dtype != torch.int16 or raw_px.dim() != 3 or not raw_px.is_cuda:
            raise RuntimeError("synthesize_device expects an int16 [S,H,W] CUDA tensor of stored pixel values")
        raw_px = raw_px.contiguous()
        S, H, W = raw_px.shape
        if out is None:
            out = torch.empty_like(raw_px)
        if S == 0:
            return out
        es, el = self._engines()
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream()
            if self._streams is None:
                one = os.environ.get("DUCOSY_SINGLE_STREAM", "0") == "1"   # experiment switch: serialise the two generators
                self._streams = (cur, cur) if one else (torch.cuda.Stream(), torch.cuda.Stream())
            s_soft, s_lung = self._streams
            B = chunk_size(S, self.batch_slices)
            ys, yl = self._buffers(B, H, W)
            # make sure both engine workspaces exist before the side streams use them
            es.workspace(B, H, W)
            el.workspace(B, H, W)
            for lo in range(0, S, B):
                # a ragged tail (n < B) goes through the same two-stream pipeline: it reuses the engines' workspaces (laid
                # out from the sizes of the call) and the leading n slices of ys / yl -- nothing is drained or reallocated
                n = min(B, S - lo)
                chunk = raw_px[lo:lo + n]
                if before_chunk is not None:
                    before_chunk(lo, n)      # e.g. wait for the host->device copy of these slices (synthesize_volume)
                s_soft.wait_stream(cur)
                s_lung.wait_stream(cur)
                with torch.cuda.stream(s_soft):
                    es.forward_hu(chunk, slope, intercept, *self.soft_hu, out=ys[:n])
                with torch.cuda.stream(s_lung):
                    el.forward_hu(chunk, slope, intercept, *self.lung_hu, out=yl[:n])
                cur.wait_stream(s_soft)
                cur.wait_stream(s_lung)
                ops.dewindow_composite(chunk, ys[:n], yl[:n], slope, intercept, self.soft_hu, self.lung_hu, out=out[lo:lo + n])
                if after_chunk is not None:
                    after_chunk(lo, n)
            if postprocess:
                from .postprocess import postprocess_volume_sharded
                out.copy_(postprocess_volume_sharded(out, group))
        return out

    # ------------------------------------------------------------------ host volume (the end-to-end call)
    def synthesize_volume(self, raw_px, slope=1.0, intercept=-1024.0, out_host: torch.Tensor | None = None,
                          postprocess: bool = False, group=None):
        """raw_px: int16 [S,H,W] numpy array or (ideally pinned) CPU tensor -> merged int16 CPU tensor.
        Host->device and device->host copies are part of this call (the reference's .to(device)/.cpu())."""
        if isinstance(raw_px, np.ndarray):
            raw_px = torch.from_numpy(np.ascontiguousarray(raw_px))
        if raw_px.is_cuda:
            return self.synthesize_device(raw_px, slope, intercept, postprocess=postprocess, group=group)
        if raw_px.dtype != torch.int16 or raw_px.dim() != 3:
            raise RuntimeError("synthesize_volume expects int16 [S,H,W] stored pixel values")
        S, H, W = raw_px.shape
        with torch.cuda.device(self.device):
            if out_host is None:
                out_host = torch.empty(raw_px.shape, dtype=torch.int16, pin_memory=True)
            cur = torch.cuda.current_stream()
            if self._copy_streams is None:
                self._copy_streams = (torch.cuda.Stream(), torch.cuda.Stream())
            s_in, s_out = self._copy_streams
            dev_in = torch.empty((S, H, W), dtype=torch.int16, device=self.device)
            dev_out = torch.empty_like(dev_in)
            s_in.wait_stream(cur)
            s_out.wait_stream(cur)
            # chunk-wise pipeline: the copy engines move slices k+1 in and k-1 out while the SMs work on chunk k
            B = chunk_size(S, self.batch_slices)
            ready = {}
            with torch.cuda.stream(s_in):
                for lo in range(0, S, B):
                    dev_in[lo:lo + B].copy_(raw_px[lo:lo + B], non_blocking=True)
                    ready[lo] = torch.cuda.Event()
                    ready[lo].record(s_in)

            def before_chunk(lo, n):
                cur.wait_event(ready[lo])

            def after_chunk(lo, n):
                if postprocess:
                    return                     # the smoothing needs the whole volume: one copy at the end
                done = torch.cuda.Event()
                done.record(cur)
                s_out.wait_event(done)
                with torch.cuda.stream(s_out):
                    out_host[lo:lo + n].copy_(dev_out[lo:lo + n], non_blocking=True)

            if S:
                self.synthesize_device(dev_in, slope, intercept, out=dev_out, postprocess=postprocess, group=group,
                                       before_chunk=before_chunk, after_chunk=after_chunk)
            if postprocess:
                out_host.copy_(dev_out, non_blocking=True)
            cur.wait_stream(s_out)
            cur.synchronize()
        return out_host
