"""Host-side owner of one `dafne_ctx`: weights, workspace, and the calls into the C ABI.

One engine per (process, GPU). torch is used for device memory and streams only; every arithmetic step of the hot
path happens inside libdafne_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _capi
from .spec import ModelSpec

DET = _capi.DET_STRIDE


class DetectionWire:
    """The fixed-shape result record of one batch in ONE device buffer: [N][cap][DET] float32 detections immediately
    followed by [N] int32 counts (typed int32: the payload is never arithmetic data). `dets` / `counts` are views.
    It is what a rank contributes to the all-gather of a multi-GPU step (dafne_b200/distributed.py::gather_wire), and
    a caller that keeps one across steps has no per-step allocation: the kernels write every element (rows past the
    last detection are zero-filled by the finalize kernel)."""

    def __init__(self, N: int, cap: int, device):
        self.N, self.cap = int(N), int(cap)
        self.buf = torch.empty(self.N * self.cap * DET + self.N, dtype=torch.int32, device=device)

    @property
    def dets(self) -> torch.Tensor:
        return self.buf[: self.N * self.cap * DET].view(torch.float32).view(self.N, self.cap, DET)

    @property
    def counts(self) -> torch.Tensor:
        return self.buf[self.N * self.cap * DET:]


def _i32_array(values: Sequence[int]):
    arr = (C.c_int32 * len(values))(*[int(v) for v in values])
    return arr


class DafneEngine:
    def __init__(self, spec: ModelSpec, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise _capi.DafneError("dafne_b200 needs a CUDA device: the CUDA kernels are the only execution path")
        self.spec = spec
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.lib = _capi.lib()
        handle = C.c_void_p()
        cspec = spec.to_c()
        _capi.check(self.lib.dafne_ctx_create(C.byref(cspec), self.device.index or 0, C.byref(handle)),
                    "dafne_ctx_create")
        self._ctx = handle
        self._ws: Optional[torch.Tensor] = None
        self._shape: Optional[Tuple[int, int, int]] = None
        self._weights_ready = False
        self._in_flight: Dict[int, tuple] = {}  # ticket -> host tensors of a pipelined batch (kept alive until _end)

    def close(self) -> None:
        if getattr(self, "_ctx", None):
            self.lib.dafne_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------------------------------- weights
    def load_state_dict(self, state_dict: Dict[str, torch.Tensor]) -> None:
        """Upload tensors by detectron2 name (reference layouts, fp32); extra keys are ignored with a record."""
        from .weights import state_dict_shapes

        shapes = state_dict_shapes(self.spec)
        missing = [k for k in shapes if k not in state_dict]
        if missing:
            raise KeyError(f"state dict lacks {len(missing)} tensors, e.g. {missing[:3]}")
        names, ptrs, shp, keep = [], [], [], []
        with torch.cuda.device(self.device):
            for k, shape in shapes.items():
                t = state_dict[k]
                if tuple(t.shape) != tuple(shape):
                    raise ValueError(f"{k}: shape {tuple(t.shape)} != expected {tuple(shape)}")
                t = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
                keep.append(t)
                names.append(k.encode())
                ptrs.append(t.data_ptr())
                s4 = list(shape) + [1] * (4 - len(shape))
                shp.extend(s4)
            n = len(names)
            c_names = (C.c_char_p * n)(*names)
            c_ptrs = (C.c_void_p * n)(*ptrs)
            c_shp = (C.c_int64 * (4 * n))(*shp)
            stream = _capi.stream_ptr()
            _capi.check(self.lib.dafne_load_weights(self._ctx, n, c_names, c_ptrs, c_shp, stream), "dafne_load_weights")
            _capi.check(self.lib.dafne_weights_finalize(self._ctx, stream), "dafne_weights_finalize")
            torch.cuda.current_stream().synchronize()  # the staging tensors in `keep` may now be released
        self._weights_ready = True

    # ---------------------------------------------------------------------------------------- workspace
    def bind(self, N: int, H: int, W: int) -> None:
        if self._shape == (N, H, W):
            return
        if not self._weights_ready:
            raise _capi.DafneError("load_state_dict() must be called before running the model")
        if self._in_flight:
            raise _capi.DafneError("bind(): pipelined batches are still in flight (call detect_host_end first)")
        need = C.c_size_t()
        _capi.check(self.lib.dafne_workspace_bytes(self._ctx, N, H, W, C.byref(need)), "dafne_workspace_bytes")
        if self._ws is None or self._ws.numel() < need.value + 1024:
            if self._ws is not None:
                # the context's private copy streams are invisible to torch's caching allocator: nothing may still
                # be running on the old workspace when its block goes back to the pool
                with torch.cuda.device(self.device):
                    torch.cuda.synchronize()
            self._ws = None
            self._ws = torch.empty(need.value + 1024, dtype=torch.uint8, device=self.device)
        base = (self._ws.data_ptr() + 1023) // 1024 * 1024
        _capi.check(self.lib.dafne_bind_workspace(self._ctx, N, H, W, base, need.value), "dafne_bind_workspace")
        self._shape = (N, H, W)
        self.workspace_bytes = need.value

    # ---------------------------------------------------------------------------------------- hot path
    def forward_dense(self, images: torch.Tensor, image_sizes: Sequence[Tuple[int, int]]) -> None:
        N, _, H, W = images.shape
        self.bind(N, H, W)
        dtype = {torch.uint8: 0, torch.float32: 1}[images.dtype]
        sizes = _i32_array([v for hw in image_sizes for v in hw])
        _capi.check(self.lib.dafne_forward_dense(self._ctx, images.data_ptr(), dtype, sizes, _capi.stream_ptr()),
                    "dafne_forward_dense")

    def head_outputs(self, level: int) -> Dict[str, torch.Tensor]:
        """Copies of the head outputs of the last forward_dense for one level, in the reference's NCHW form."""
        N = self._shape[0]
        out = {}
        for which, (name, nvalid) in enumerate((("logits", self.spec.num_classes), ("ctr_delta", 9), ("center", 2))):
            p, ld, h, w = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
            _capi.check(self.lib.dafne_head_output(self._ctx, level, which, C.byref(p), C.byref(ld), C.byref(h),
                                                   C.byref(w)), "dafne_head_output")
            n_el = N * h.value * w.value * ld.value
            t = torch.as_tensor(_DeviceFloats(p.value, n_el), device=self.device)  # zero-copy view of the workspace
            t = t.view(N, h.value, w.value, ld.value)[..., :nvalid].permute(0, 3, 1, 2).contiguous()
            out[name] = t
        return out

    def keep_activations(self, keep: bool = True) -> None:
        """Per-layer parity support: disable activation-memory reuse (call before the first forward)."""
        _capi.check(self.lib.dafne_debug_keep_activations(self._ctx, int(keep)), "dafne_debug_keep_activations")
        self._shape = None

    def activation(self, name: str) -> torch.Tensor:
        """NCHW float32 copy of a named intermediate of the last forward (needs keep_activations)."""
        p, n, h, w, c = C.c_void_p(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _capi.check(self.lib.dafne_debug_activation(self._ctx, name.encode(), C.byref(p), C.byref(n), C.byref(h),
                                                    C.byref(w), C.byref(c)), "dafne_debug_activation")
        n_el = n.value * h.value * w.value * c.value
        t = torch.as_tensor(_DeviceHalfs(p.value, n_el), device=self.device)
        return t.view(n.value, h.value, w.value, c.value).permute(0, 3, 1, 2).float().contiguous()

    def postprocess(self, image_sizes, output_sizes=None, do_postprocess=True, capacity: Optional[int] = None,
                    out: Optional[DetectionWire] = None):
        """-> (dets [N, cap, 20] fp32, counts [N] int32) on the device. `out`: a DetectionWire to write into (no
        allocation in the call); otherwise fresh tensors are returned."""
        N = self._shape[0]
        cap = capacity or (self.spec.post_nms_topk + 64)
        if out is None:
            out = DetectionWire(N, cap, self.device)  # torch.empty: the kernels write every element
        elif (out.N, out.cap) != (N, cap) or out.buf.device != self.device:
            raise ValueError(f"DetectionWire is [{out.N}, {out.cap}] on {out.buf.device}, need [{N}, {cap}] on {self.device}")
        dets, counts = out.dets, out.counts
        sizes = _i32_array([v for hw in image_sizes for v in hw])
        osz = _i32_array([v for hw in (output_sizes or image_sizes) for v in hw])
        _capi.check(self.lib.dafne_postprocess(self._ctx, sizes, osz, int(do_postprocess), dets.data_ptr(),
                                               counts.data_ptr(), cap, _capi.stream_ptr()), "dafne_postprocess")
        return dets, counts

    def detect(self, images: torch.Tensor, image_sizes, output_sizes=None, do_postprocess=True,
               capacity: Optional[int] = None, out: Optional[DetectionWire] = None):
        """Device tensors in, device tensors out: dets [N, cap, 20] fp32, counts [N] int32. No host sync."""
        self.forward_dense(images, image_sizes)
        return self.postprocess(image_sizes, output_sizes, do_postprocess, capacity, out)

    def capture(self, images: torch.Tensor, image_sizes, output_sizes=None, do_postprocess=True,
                capacity: Optional[int] = None, out: Optional[DetectionWire] = None):
        """Capture the whole step as one CUDA graph bound to THESE tensors (`images` is read, `out` written on every
        `replay()`; the caller refreshes the contents of `images` in place). Returns (dets, counts) views of `out`."""
        N, _, H, W = images.shape
        self.bind(N, H, W)
        cap = capacity or (self.spec.post_nms_topk + 64)
        out = out if out is not None else DetectionWire(N, cap, self.device)
        dtype = {torch.uint8: 0, torch.float32: 1}[images.dtype]
        sizes = _i32_array([v for hw in image_sizes for v in hw])
        osz = _i32_array([v for hw in (output_sizes or image_sizes) for v in hw])
        # stream capture is not allowed on the legacy default stream: capture on a side stream ordered after the
        # current one (the replay may go to any stream)
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            _capi.check(self.lib.dafne_graph_capture(self._ctx, images.data_ptr(), dtype, sizes, osz,
                                                     int(do_postprocess), out.dets.data_ptr(), out.counts.data_ptr(),
                                                     cap, _capi.stream_ptr()), "dafne_graph_capture")
        cur.wait_stream(side)
        self._graph_refs = (images, out)  # the graph holds raw pointers into these
        return out.dets, out.counts

    def replay(self) -> None:
        """One step = one cudaGraphLaunch on the current stream."""
        _capi.check(self.lib.dafne_graph_launch(self._ctx, _capi.stream_ptr()), "dafne_graph_launch")

    def _check_host_buffers(self, host_images, host_dets, host_counts, N, cap):
        """The C side copies N * cap * 20 floats packed at row stride `cap`, asynchronously: shapes must match
        exactly, the tensors must be contiguous, and pinned (a pageable buffer would make the copies synchronous and
        the pipelined form meaningless)."""
        if tuple(host_dets.shape) != (N, cap, DET) or host_dets.dtype != torch.float32 or not host_dets.is_contiguous():
            raise ValueError(f"host_dets must be a contiguous float32 [{N}, {cap}, {DET}] tensor, got "
                             f"{tuple(host_dets.shape)} {host_dets.dtype}")
        if tuple(host_counts.shape) != (N,) or host_counts.dtype != torch.int32 or not host_counts.is_contiguous():
            raise ValueError(f"host_counts must be a contiguous int32 [{N}] tensor, got {tuple(host_counts.shape)}")
        for name, t in (("host_images", host_images), ("host_dets", host_dets), ("host_counts", host_counts)):
            if t.is_cuda or not t.is_contiguous():
                raise ValueError(f"{name} must be a contiguous HOST tensor")
        for name, t in (("host_dets", host_dets), ("host_counts", host_counts)):
            if not t.is_pinned():
                raise ValueError(f"{name} must be pinned host memory (tensor.pin_memory())")

    def detect_host(self, host_images: torch.Tensor, image_sizes, output_sizes=None, host_dets=None,
                    host_counts=None, capacity: Optional[int] = None):
        """The reference-facing call with HOST buffers (pinned for full bandwidth): H2D + detect + D2H + sync."""
        N, _, H, W = host_images.shape
        self.bind(N, H, W)
        cap = min(capacity or (self.spec.post_nms_topk + 64), 2048)
        if host_dets is None:
            host_dets = torch.empty(N, cap, DET, dtype=torch.float32).pin_memory()
        if host_counts is None:
            host_counts = torch.empty(N, dtype=torch.int32).pin_memory()
        self._check_host_buffers(host_images, host_dets, host_counts, N, cap)
        dtype = {torch.uint8: 0, torch.float32: 1}[host_images.dtype]
        sizes = _i32_array([v for hw in image_sizes for v in hw])
        osz = _i32_array([v for hw in (output_sizes or image_sizes) for v in hw])
        _capi.check(self.lib.dafne_detect_host(self._ctx, host_images.data_ptr(), dtype, sizes, osz,
                                               host_dets.data_ptr(), host_counts.data_ptr(), cap,
                                               _capi.stream_ptr()), "dafne_detect_host")
        return host_dets, host_counts

    def detect_host_begin(self, host_images: torch.Tensor, image_sizes, output_sizes, host_dets: torch.Tensor,
                          host_counts: torch.Tensor, capacity: Optional[int] = None) -> int:
        """Pipelined detect_host: enqueue H2D (copy stream) + detect + D2H, return a ticket without synchronising.
        Up to two batches in flight; `host_dets` / `host_counts` (pinned) are filled when detect_host_end returns.
        The engine keeps the host tensors referenced until then."""
        N, _, H, W = host_images.shape
        self.bind(N, H, W)
        cap = min(capacity or (self.spec.post_nms_topk + 64), 2048)
        self._check_host_buffers(host_images, host_dets, host_counts, N, cap)
        dtype = {torch.uint8: 0, torch.float32: 1}[host_images.dtype]
        sizes = _i32_array([v for hw in image_sizes for v in hw])
        osz = _i32_array([v for hw in (output_sizes or image_sizes) for v in hw])
        ticket = C.c_int(-1)
        _capi.check(self.lib.dafne_detect_host_begin(self._ctx, host_images.data_ptr(), dtype, sizes, osz,
                                                     host_dets.data_ptr(), host_counts.data_ptr(), cap,
                                                     _capi.stream_ptr(), C.byref(ticket)), "dafne_detect_host_begin")
        self._in_flight[ticket.value] = (host_images, host_dets, host_counts)
        return ticket.value

    def detect_host_end(self, ticket: int) -> None:
        _capi.check(self.lib.dafne_detect_host_end(self._ctx, int(ticket)), "dafne_detect_host_end")
        self._in_flight.pop(int(ticket), None)

    def slot_wire(self, ticket: int) -> torch.Tensor:
        """Zero-copy int32 view of the DEVICE result record of a pipelined batch (detections, then counts -- the layout
        of DetectionWire.buf), valid in stream order after detect_host_begin(ticket): what a multi-GPU step
        all-gathers, with no host bounce."""
        p, nbytes, cap = C.c_void_p(), C.c_size_t(), C.c_int()
        _capi.check(self.lib.dafne_host_slot_wire(self._ctx, int(ticket), C.byref(p), C.byref(nbytes), C.byref(cap)),
                    "dafne_host_slot_wire")
        return torch.as_tensor(_DeviceInts(p.value, nbytes.value // 4), device=self.device)

    def postprocess_external(self, logits: List[torch.Tensor], reg: List[torch.Tensor], ctr: List[torch.Tensor],
                             image_sizes, output_sizes=None, do_postprocess=True, capacity: Optional[int] = None):
        """Post-process head outputs given in the reference's own NCHW fp32 form (the bit-exact parity gate)."""
        N = logits[0].shape[0]
        level_hw = [v for t in logits for v in t.shape[2:]]
        keep = []
        lp, rp, cp = [], [], []
        for lg, rg, ct in zip(logits, reg, ctr):
            # NCHW -> NHWC views materialised: pure data movement, no arithmetic
            a = lg.to(self.device, torch.float32).permute(0, 2, 3, 1).contiguous()
            b = rg.to(self.device, torch.float32).permute(0, 2, 3, 1).contiguous()
            c = ct.to(self.device, torch.float32).permute(0, 2, 3, 1).contiguous()
            keep += [a, b, c]
            lp.append(a.data_ptr())
            rp.append(b.data_ptr())
            cp.append(c.data_ptr())
        hw = _i32_array(level_hw)
        need = C.c_size_t()
        _capi.check(self.lib.dafne_postprocess_scratch_bytes(self._ctx, N, hw, C.byref(need)), "scratch_bytes")
        scratch = torch.empty(need.value + 1024, dtype=torch.uint8, device=self.device)
        sbase = (scratch.data_ptr() + 1023) // 1024 * 1024
        cap = capacity or (self.spec.post_nms_topk + 64)
        wire = DetectionWire(N, cap, self.device)
        dets, counts = wire.dets, wire.counts
        sizes = _i32_array([v for s in image_sizes for v in s])
        osz = _i32_array([v for s in (output_sizes or image_sizes) for v in s])
        L = len(logits)
        _capi.check(
            self.lib.dafne_postprocess_external(
                self._ctx, N, hw, (C.c_void_p * L)(*lp), (C.c_void_p * L)(*rp), (C.c_void_p * L)(*cp), sizes, osz,
                int(do_postprocess), dets.data_ptr(), counts.data_ptr(), cap, sbase, need.value, _capi.stream_ptr()),
            "dafne_postprocess_external")
        torch.cuda.current_stream().synchronize()
        return dets, counts

    def profile_forward(self, images: torch.Tensor, image_sizes) -> List[dict]:
        """One dense forward with a CUDA event after every launch; returns per-launch records (ms, flops, ...)."""
        _capi.check(self.lib.dafne_set_profiling(self._ctx, 1), "dafne_set_profiling")
        try:
            self.forward_dense(images, image_sizes)
            torch.cuda.current_stream().synchronize()
        finally:
            _capi.check(self.lib.dafne_set_profiling(self._ctx, 0), "dafne_set_profiling")
        cnt = C.c_int()
        _capi.check(self.lib.dafne_get_profile(self._ctx, None, 0, C.byref(cnt)), "dafne_get_profile")
        arr = (_capi.OpProfileC * cnt.value)()
        _capi.check(self.lib.dafne_get_profile(self._ctx, C.cast(arr, C.c_void_p), cnt.value, C.byref(cnt)),
                    "dafne_get_profile")
        return [
            dict(name=o.name.decode(), ms=o.ms, kind=o.kind, block_n=o.block_n, ksize=o.ksize, stride=o.stride,
                 cin=o.cin, cout=o.cout, hout=o.hout, wout=o.wout, flops=o.flops, bytes=o.bytes)
            for o in arr
        ]

    def post_counts(self) -> List[dict]:
        """Diagnostic of the last post-processing (synchronises): per image the candidates per level, boxes entering
        NMS and boxes kept by NMS before the post-NMS top-k."""
        N = self._shape[0]
        arr = (C.c_int32 * (8 * N))()
        _capi.check(self.lib.dafne_debug_post_counts(self._ctx, arr, _capi.stream_ptr()), "dafne_debug_post_counts")
        return [dict(candidates=list(arr[8 * n:8 * n + 5]), nms_in=arr[8 * n + 5], nms_kept=arr[8 * n + 6],
                     capacity=arr[8 * n + 7]) for n in range(N)]

    def nms_stats(self) -> dict:
        """Work counters of the rotated NMS of the last post-processing (synchronises), summed over the batch."""
        arr = (C.c_uint64 * 8)()
        _capi.check(self.lib.dafne_debug_nms_stats(self._ctx, arr, _capi.stream_ptr()), "dafne_debug_nms_stats")
        return dict(diag_pairs=arr[0], diag_clipped_pairs=arr[1], bcast_pairs=arr[3], bcast_clipped_pairs=arr[4])

    def stats(self, reset: bool = False) -> Tuple[int, float]:
        launches, flops = C.c_int64(), C.c_double()
        _capi.check(self.lib.dafne_stats(self._ctx, C.byref(launches), C.byref(flops), int(reset)), "dafne_stats")
        return launches.value, flops.value


class _DeviceFloats:
    """`__cuda_array_interface__` holder: lets torch view caller-owned device memory without copying."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class _DeviceInts:
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}


class _DeviceHalfs:
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f2", "data": (ptr, False), "version": 2}
