"""N>1 host logic on CPU: world_size-2 gloo processes shard 1024 streams with no data-path collective and
aggregate throughput as total frames / max time over ranks (the bench.py contract)."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from infercam_onnx_b200 import streams


def test_sharding_is_a_partition():
    for world in (1, 2, 4, 8):
        parts = [streams.shard_streams(1024, r, world) for r in range(world)]
        flat = sorted(s for p in parts for s in p)
        assert flat == list(range(1024))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
        assert all(streams.owner(s, world) == r for r, p in enumerate(parts) for s in p)
    assert [len(b) for b in streams.batches(list(range(600)), 256)] == [256, 256, 88]
    assert streams.stream_id("cam0") == streams.stream_id("cam0") != streams.stream_id("cam1")
    with pytest.raises(ValueError):
        streams.shard_streams(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = streams.shard_streams(1024, rank, world)
    # every rank drives ITS streams through its own batcher (the C++ queue / router with an injected backend: no GPU here)
    import numpy as np
    from infercam_onnx_b200.batcher import StreamBatcher
    b = StreamBatcher(backend=lambda dev, frames: [np.float32([[0, 0, 1, 1, float(f[0, 0, 0])]]) for f in frames],
                      max_batch=64, max_delay=0.001, capacity=4096, workers=2, max_frame_bytes=64)
    for s in mine:
        f = np.zeros((2, 2, 3), np.uint8)
        f[0, 0, 0] = s % 251
        assert b.try_submit(s, f, tag=s)
    b.flush()
    res = b.poll(2048)
    assert sorted(r["tag"] for r in res) == mine and all(r["dets"][0, 4] == r["stream"] % 251 for r in res)
    b.close()
    seconds = 1.0 + rank  # rank 1 is the slow one
    fps = streams.aggregate_throughput(dist, len(mine), seconds)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    dist.barrier()
    q.put((rank, fps, sorted(s for g in gathered for s in g) == list(range(1024))))
    dist.destroy_process_group()


def test_two_rank_gloo_aggregation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, fps, complete in res:
        assert complete
        assert fps == pytest.approx(1024 / 2.0)  # total frames / MAX time over ranks
