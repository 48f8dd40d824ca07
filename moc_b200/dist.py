"""Slide sharding across the GPUs of one box (one process per GPU, torch.distributed).

Slides are independent units for scoring, selection, gate and pooling (main_moc.py:472-498), so the eval
splits are partitioned statically over the ranks with longest-processing-time bin packing on patch count
(bag sizes vary 100x) and no feature byte ever crosses NVLink.  The only exchange is one all-gather of
``[n_local, C+1]`` floats (bag logits + label) per evaluation pass, so that every rank computes identical
loss / acc / AUC, and - in the optional data-parallel training mode - one all-reduce of the 33 092 gate
gradients per micro-batch.  The few-shot training bags are replicated: the reference's training is one
sequential Adam step per slide (main_moc.py:380-410) and stays bit-faithful that way.
"""
from __future__ import annotations

import heapq
import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def lpt_shards(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Deterministic LPT partition: slide ids per rank, each list ascending.  Ties by slide id."""
    loads = [(0, r) for r in range(world)]
    heapq.heapify(loads)
    out: List[List[int]] = [[] for _ in range(world)]
    for i in sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i)):
        load, r = heapq.heappop(loads)
        out[r].append(i)
        heapq.heappush(loads, (load + int(sizes[i]), r))
    return [sorted(ids) for ids in out]


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world) from the torchrun environment; initialises the process group if world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local, world


class Shard:
    """This rank's part of a split and how to put the pieces back together."""

    def __init__(self, sizes: Sequence[int], rank: int, world: int, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.n_global = len(sizes)
        self.all_ids = lpt_shards(sizes, world)
        self.ids = self.all_ids[rank]
        self.max_local = max((len(v) for v in self.all_ids), default=0)

    def global_len(self, ds) -> int:
        return self.n_global

    def gather(self, logits: torch.Tensor, labels: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """[n_local,C] logits + [n_local] labels of every rank -> ([n_global,C], [n_global]) in split order."""
        n_local, c = logits.shape
        assert n_local == len(self.ids)
        if self.world == 1:
            return logits, labels
        buf = torch.zeros(self.max_local, c + 1, dtype=torch.float32, device=logits.device)
        buf[:n_local, :c] = logits
        buf[:n_local, c] = labels.to(torch.float32)
        parts = [torch.empty_like(buf) for _ in range(self.world)]
        dist.all_gather(parts, buf, group=self.group)
        out = torch.empty(self.n_global, c, dtype=torch.float32, device=logits.device)
        lab = torch.empty(self.n_global, dtype=torch.int64, device=logits.device)
        for r, ids in enumerate(self.all_ids):
            if ids:
                idx = torch.tensor(ids, dtype=torch.int64, device=logits.device)
                out[idx] = parts[r][:len(ids), :c]
                lab[idx] = parts[r][:len(ids), c].to(torch.int64)
        return out, lab


def allreduce_sum(flat: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of the flat gate gradient over ranks (data-parallel training mode; 132 KB, latency-bound)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def barrier_max_ms(ms: float, device) -> float:
    """Max over ranks of a locally measured duration."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
