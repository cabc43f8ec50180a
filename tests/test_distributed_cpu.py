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
        from dafne_b200.distributed import gather_wire

        wire = torch.cat([torch.full((2 * 4 * 20,), float(rank)).view(torch.int32), torch.full((2,), rank + 7, dtype=torch.int32)])
        d3, c3 = gather_wire(wire, 2, 4, 20)
        out_q.put(("wire", rank, d3[:, :, 0, 0].reshape(-1).tolist(), c3.reshape(-1).tolist()))
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
    results = [q.get(timeout=120) for _ in range(6)]
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
    third = [r for r in results if r[0] == "wire"]
    assert len(third) == 2
    for _, _, tags, counts in third:  # one collective on the engine's record layout, rank-major
        assert tags == [0.0, 0.0, 1.0, 1.0] and counts == [7, 7, 8, 8]


def test_shard_range():
    assert [list(shard_range(5, r, 2)) for r in range(2)] == [[0, 1, 2], [3, 4]]
    assert [list(shard_range(2, r, 4)) for r in range(4)] == [[0], [1], [], []]
    assert sum(len(shard_range(128, r, 8)) for r in range(8)) == 128


def test_wire_format_round_trips_bit_exactly():
    """The wire is typed int32 (raw bits): NaN payloads, denormals and -0.0 in the detections survive unchanged."""
    g = torch.Generator().manual_seed(0)
    dets = torch.randn(3, 5, 20, generator=g)
    dets[0, 0, :4] = torch.tensor([float("nan"), -0.0, 1e-42, float("inf")])
    counts = torch.tensor([0, 5, 2_000_000_000], dtype=torch.int32)
    wire = pack_wire(dets, counts)
    assert wire.shape == (3, 101) and wire.dtype == torch.int32
    d2, c2 = unpack_wire(wire.clone(), dets.shape[1:])
    assert torch.equal(d2.view(torch.int32), dets.view(torch.int32)) and torch.equal(c2, counts)


def test_gather_wire_single_process_views():
    """gather_wire on the engine's record layout (detections then counts in one int32 buffer), world size 1."""
    from dafne_b200.distributed import gather_wire

    n, cap, det = 2, 3, 20
    dets = torch.arange(n * cap * det, dtype=torch.float32).view(n, cap, det)
    counts = torch.tensor([3, 1], dtype=torch.int32)
    wire = torch.cat([dets.view(torch.int32).reshape(-1), counts])
    d, c = gather_wire(wire, n, cap, det)
    assert d.shape == (1, n, cap, det) and c.shape == (1, n)
    assert torch.equal(d[0], dets) and torch.equal(c[0], counts)
