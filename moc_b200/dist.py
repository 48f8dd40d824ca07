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
        self._plans = {}

    def global_len(self, ds) -> int:
        return self.n_global

    def _plan(self, device):
        """Device-resident index of every global slide inside the padded all-gather buffer (built once per device)."""
        key = str(device)
        hit = self._plans.get(key)
        if hit is None:
            pos = [0] * self.n_global
            for r, ids in enumerate(self.all_ids):
                for k, i in enumerate(ids):
                    pos[i] = r * self.max_local + k
            hit = self._plans[key] = torch.tensor(pos, dtype=torch.int64, device=device)
        return hit

    def gather(self, logits: torch.Tensor, labels: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """[n_local,C] logits + [n_local] labels of every rank -> ([n_global,C], [n_global]) in split order: one padded
        all-gather (the only exchange of an evaluation pass, <= 276 KB even for EBRAINS-30) and one index-select."""
        n_local, c = logits.shape
        assert n_local == len(self.ids)
        if self.world == 1:
            return logits, labels
        buf = torch.zeros(self.max_local, c + 1, dtype=torch.float32, device=logits.device)
        buf[:n_local, :c] = logits
        buf[:n_local, c] = labels.to(torch.float32)
        flat = torch.empty(self.world * self.max_local, c + 1, dtype=torch.float32, device=logits.device)
        dist.all_gather_into_tensor(flat, buf, group=self.group)
        rows = flat.index_select(0, self._plan(logits.device))
        return rows[:, :c].contiguous(), rows[:, c].to(torch.int64)


def bind_to_gpu_numa_node(local_rank: int) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (read from sysfs), so that the pinned host
    buffers it allocates afterwards are first-touched on that node and the H2D copies do not cross the socket link.
    Returns the node, or None when the topology cannot be read (nothing is changed then)."""
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev_id = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0" % (dom, bus, dev_id)
        node = int(open(os.path.join(path, "numa_node")).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def sync_seed(seed: Optional[int] = None, group=None) -> Optional[int]:
    """Every rank seeds torch's CPU generator with the same value: rank 0's ``seed`` argument, or - when the run is
    unseeded like the reference's - a fresh random one that rank 0 draws.  The gate's initial weights and the per-step
    half masks (``torch.rand(N) > 0.5`` on the CPU default generator, main_moc.py:330) then come out identical on all
    ranks, which is what makes replicated training bit-identical.  Single process: seeds only if ``seed`` is given."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        if seed is not None:
            torch.manual_seed(int(seed))
        return seed
    val = int(seed) if seed is not None else int.from_bytes(os.urandom(7), "little")
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([val], dtype=torch.int64, device=dev)
    dist.broadcast(t, src=0, group=group)
    val = int(t.item())
    torch.manual_seed(val)
    return val


def broadcast_parameters(module: torch.nn.Module, group=None) -> None:
    """Rank 0's parameters and buffers overwrite every other rank's (in place)."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=0, group=group)


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
