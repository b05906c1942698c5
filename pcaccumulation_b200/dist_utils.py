"""Multi-GPU plumbing: scenes are independent, so ranks shard them round-robin and the
only collective is the reduction of a few metric counters (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_scenes(scenes, rank, world):
    """scene i -> rank i mod world"""
    return list(scenes[rank::world])


def reduce_metrics(t, op="sum"):
    """All-reduce a small tensor of counters (sum) or timings (max); no-op without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX)
    return t
