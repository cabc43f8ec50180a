"""world_size-2 gloo test (CPU) of the N>1 host logic: sharding, padding of a ragged last shard, the all-gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dafne_b200.distributed import detect_sharded, gather_detections, pack_wire, shard_range, unpack_wire


class FakeEngine:
    """Stands in for DafneEngine.detect: row i of the output encodes the image's first pixel, so order is checkable."""

    def detect(self, images, sizes, osz, do_pp, capacity):
        n = images.shape[0]
        dets = torch.zeros(n, 4, 20)
        counts = torch.zeros(n, dtype=torch.int32)
        for i in range(n):
            tag = float(images[i, 0, 0, 0])
            dets[i, 0, :] = tag
            counts[i] = int(tag) % 3 + 1
            assert sizes[i][0] > 0
        return dets, counts


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_images, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        images = torch.arange(n_images, dtype=torch.float32).view(n_images, 1, 1, 1).expand(n_images, 3, 2, 2).contiguous()
        sizes = [(2, 2)] * n_images
        dets, counts = detect_sharded(FakeEngine(), images, sizes, None, 4)
        out_q.put(("sharded", rank, dets[:, 0, 0].tolist(), counts.tolist()))
        d2, c2 = gather_detections(torch.full((2, 4, 20), float(rank)), torch.full((2,), rank, dtype=torch.int32))
        out_q.put(("gather", rank, d2[:, 0, 0].tolist(), c2.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [4, 5])
def test_sharded_detect_two_ranks_gloo(n_images):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_images, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(4)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    first = [r for r in results if r[0] == "sharded"]
    assert len(first) == 2
    for _, _, tags, counts in first:  # every rank sees every image, in global order, padding rows dropped
        assert tags == [float(i) for i in range(n_images)]
        assert counts == [i % 3 + 1 for i in range(n_images)]
    second = [r for r in results if r[0] == "gather"]
    assert len(second) == 2
    for _, _, tags, counts in second:
        assert tags == [0.0, 0.0, 1.0, 1.0] and counts == [0, 0, 1, 1]


def test_shard_range():
    assert [list(shard_range(5, r, 2)) for r in range(2)] == [[0, 1, 2], [3, 4]]
    assert [list(shard_range(2, r, 4)) for r in range(4)] == [[0], [1], [], []]
    assert sum(len(shard_range(128, r, 8)) for r in range(8)) == 128


def test_wire_format_round_trips_counts_bit_exactly():
    g = torch.Generator().manual_seed(0)
    dets = torch.randn(3, 5, 20, generator=g)
    counts = torch.tensor([0, 5, 2_000_000_000], dtype=torch.int32)  # as raw bits: includes NaN-looking patterns
    wire = pack_wire(dets, counts)
    assert wire.shape == (3, 101) and wire.dtype == torch.float32
    d2, c2 = unpack_wire(wire.clone(), dets.shape[1:])
    assert torch.equal(d2, dets) and torch.equal(c2, counts)

