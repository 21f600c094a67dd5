"""Stream -> GPU sharding for the multi-GPU path.

Frames are independent (`run(&self)`, /root/reference/infer_server/src/nn.rs:179) and the reference keys every
stream by `hashed(name)` (infer_server/src/lib.rs:39-46, router.rs:58), so the path shards by stream with no
data-path collective: rank r of n owns the streams with `stream_id % n == r`. The only cross-rank traffic in this
repo is the benchmark's timing barrier / max-reduce (torch.distributed: NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence


def stream_id(name: str) -> int:
    """The reference's stream key `hashed(&id)` (lib.rs:39-46: DefaultHasher = SipHash-1-3 with zero keys), computed by
    the library (uf_stream_hash) so that the Rust server, the Python mirror and the batcher's routing agree."""
    from .batcher import stream_hash
    return stream_hash(name)


def owner(stream: int, world_size: int) -> int:
    return stream % world_size


def shard_streams(n_streams: int, rank: int, world_size: int) -> List[int]:
    """Stream ids owned by `rank` (disjoint over ranks, union = range(n_streams), sizes differ by at most 1)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    return list(range(rank, n_streams, world_size))


def batches(streams: Sequence[int], batch: int) -> List[List[int]]:
    """One step takes one frame from each stream of a batch; a rank with S streams needs ceil(S/batch) steps."""
    return [list(streams[i:i + batch]) for i in range(0, len(streams), batch)]


def aggregate_throughput(dist, frames_this_rank: int, seconds_this_rank: float) -> float:
    """Whole-job frames/s = total frames of all ranks / max time over ranks (bench.py contract)."""
    import torch
    t = torch.tensor([float(frames_this_rank), 0.0], dtype=torch.float64)
    m = torch.tensor([float(seconds_this_rank)], dtype=torch.float64)
    if dist is not None and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return float(t[0].item() / m[0].item())


def bind_host_thread_near_gpu(device_index: int) -> bool:
    """Pin the calling process to the CPU cores NVML reports as local to the GPU (same NUMA node / PCIe root), so
    that the pinned staging memory allocated afterwards and the H2D copies of this rank do not cross sockets. With 8
    ranks feeding 8 GPUs at ~55 GB/s each, host placement is what bounds end-to-end scaling. Best effort: returns
    False (and changes nothing) when NVML or sched_setaffinity is unavailable."""
    try:
        import os

        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            n_words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        finally:
            pynvml.nvmlShutdown()
        cpus = {w * 64 + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False
