"""Multi-GPU inference: images shard by batch across ranks (one process per GPU), no data-path collective during the
forward or the NMS (every image is independent), then ONE all-gather of the fixed-shape detections.

Replaces the reference's pickled `comm.gather(self._predictions, dst=0)` (dafne/evaluation/dafne_evaluator.py:61-64)
and detectron2's InferenceSampler sharding (tools/plain_train_net.py:280-313).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous shard of `n_items` for `rank` (detectron2 InferenceSampler semantics: ceil-sized leading shards)."""
    per = (n_items + world - 1) // world
    begin = min(rank * per, n_items)
    return range(begin, min(begin + per, n_items))


def pack_wire(dets: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    """[B, cap, D] float32 detections + [B] int32 counts -> one [B, cap * D + 1] int32 buffer (the detections as raw
    bits: the wire is never typed as arithmetic data), so the exchange is ONE collective."""
    b = dets.shape[0]
    wire = torch.empty(b, dets[0].numel() + 1, dtype=torch.int32, device=dets.device)
    wire[:, :-1] = dets.contiguous().view(torch.int32).reshape(b, -1)
    wire[:, -1] = counts.to(torch.int32)
    return wire


def unpack_wire(wire: torch.Tensor, det_shape) -> Tuple[torch.Tensor, torch.Tensor]:
    n = wire.shape[0]
    dets = wire[:, :-1].contiguous().view(torch.float32).reshape((n,) + tuple(det_shape))
    counts = wire[:, -1].contiguous()
    return dets, counts


def gather_detections(dets: torch.Tensor, counts: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather [B_local, cap, 20] detections and [B_local] counts -> ([world*B_local, cap, 20], [world*B_local]),
    rank-major: exactly one all-gather, issued on the current stream (NCCL) right after the last kernel.
    Generic form (packs a wire buffer); a caller that holds its results in a DetectionWire uses gather_wire()."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dets, counts
    world = dist.get_world_size(group)
    wire = pack_wire(dets, counts)
    if dist.get_backend(group) == "nccl":
        out = torch.empty(world * wire.shape[0], wire.shape[1], dtype=wire.dtype, device=wire.device)
        dist.all_gather_into_tensor(out, wire, group=group)
    else:
        parts = [torch.empty_like(wire) for _ in range(world)]
        dist.all_gather(parts, wire, group=group)
        out = torch.cat(parts, 0)
    return unpack_wire(out, dets.shape[1:])


def gather_wire(wire: torch.Tensor, n_local: int, cap: int, det_stride: int, out: torch.Tensor = None, group=None):
    """The path's one exchange with no packing and no allocation: `wire` is a rank's result record (int32
    [n_local * cap * det_stride + n_local]: detections, then counts -- engine.DetectionWire.buf or
    DafneEngine.slot_wire(ticket)), `out` an int32 [world, len(wire)] buffer. ONE all_gather_into_tensor on the current
    stream; returns views (dets [world, n_local, cap, det_stride] float32, counts [world, n_local] int32) of `out`."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    nd = n_local * cap * det_stride
    assert wire.dtype == torch.int32 and wire.numel() == nd + n_local
    if out is None:
        out = torch.empty(world, wire.numel(), dtype=torch.int32, device=wire.device)
    if world == 1:
        out[0].copy_(wire)
    elif dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, wire, group=group)
    else:
        parts = [torch.empty_like(wire) for _ in range(world)]
        dist.all_gather(parts, wire, group=group)
        out.copy_(torch.stack(parts, 0))
    dets = out[:, :nd].view(torch.float32).view(world, n_local, cap, det_stride)
    counts = out[:, nd:]
    return dets, counts


def pad_shard(items: Sequence, per_rank: int, filler):
    """Every rank must contribute the same shape: pad a short (last) shard with `filler` items."""
    items = list(items)
    return items + [filler] * (per_rank - len(items)), len(items)


def detect_sharded(engine, images: torch.Tensor, image_sizes: Sequence[Tuple[int, int]], output_sizes=None,
                   capacity=None, group=None):
    """`images` is the GLOBAL batch (same on every rank, or only this rank's slice is ever touched); each rank runs its
    contiguous shard and all ranks receive every image's detections. Returns (dets, counts) for the global batch."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = images.shape[0]
    per = (n + world - 1) // world
    r = shard_range(n, rank, world)
    local = images[r.start:r.stop]
    sizes = list(image_sizes[r.start:r.stop])
    osz = list(output_sizes[r.start:r.stop]) if output_sizes is not None else None
    if len(r) < per:  # pad the last shard with copies of a valid image; its rows are dropped after the gather
        pad = per - len(r)
        filler = images[:1] if len(r) == 0 else local[-1:]
        local = torch.cat([local] + [filler] * pad, 0)
        fs = image_sizes[0] if len(r) == 0 else sizes[-1]
        sizes = sizes + [fs] * pad
        if osz is not None:
            osz = osz + [(output_sizes[0] if len(r) == 0 else osz[-1])] * pad
    dets, counts = engine.detect(local.contiguous(), sizes, osz, True, capacity)
    dets, counts = gather_detections(dets, counts, group)
    if world * per == n:
        return dets[:n], counts[:n]
    return _drop_padding(dets, counts, n, per, world)


def _drop_padding(dets, counts, n, per, world):
    keep = []
    for rk in range(world):
        r = shard_range(n, rk, world)
        keep.extend(range(rk * per, rk * per + len(r)))
    idx = torch.tensor(keep, device=dets.device, dtype=torch.long)
    return dets[idx], counts[idx]
