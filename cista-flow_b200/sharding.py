"""Multi-GPU plumbing: independent event streams sharded across ranks.

The hot path has no cross-stream coupling (each stream is a serial recurrence
over its own frames, test_wo_flow.py:109-149 in the reference) and the
reference has no distributed backend at all.  So: one process
per GPU, stream ``s`` lives on rank ``s % world_size``, NO collective on the data
path, and a single all_gather of per-stream metric rows at the end (NCCL on
GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_rank_world() -> tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1-process default)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_streams(n_streams: int, world_size: int, rank: int) -> list[int]:
    """Static round-robin partition: stream s -> rank s % world_size."""
    assert 0 <= rank < world_size
    return list(range(rank, n_streams, world_size))


def streams_per_rank(n_streams: int, world_size: int) -> list[int]:
    return [len(range(r, n_streams, world_size)) for r in range(world_size)]


def gather_stream_metrics(local_ids: list[int], local_rows: torch.Tensor, n_streams: int) -> torch.Tensor:
    """All-gather per-stream metric rows into a [n_streams, K] table on every rank.

    ``local_rows`` is [len(local_ids), K] (float64) on this rank's device (CUDA for
    NCCL, CPU for gloo).  Ranks may own different numbers of streams: rows are
    padded to the per-rank maximum so that one fixed-size all_gather suffices."""
    K = local_rows.shape[1]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        table = torch.zeros(n_streams, K, dtype=torch.float64)
        if local_ids:
            table[torch.tensor(local_ids)] = local_rows.double().cpu()
        return table
    world = dist.get_world_size()
    cap = max(streams_per_rank(n_streams, world))
    dev = local_rows.device
    buf = torch.full((cap, K + 1), -1.0, dtype=torch.float64, device=dev)
    if local_ids:
        buf[: len(local_ids), 0] = torch.tensor(local_ids, dtype=torch.float64, device=dev)
        buf[: len(local_ids), 1:] = local_rows.double()
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    table = torch.zeros(n_streams, K, dtype=torch.float64)
    for g in gathered:
        g = g.cpu()
        valid = g[:, 0] >= 0
        table[g[valid, 0].long()] = g[valid, 1:]
    return table


def max_over_ranks(value: float, device: torch.device) -> float:
    """Device-side timings are reduced as the MAX over ranks (never wall clock)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
