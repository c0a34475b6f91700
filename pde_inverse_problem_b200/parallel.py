"""Data-parallel plumbing: one process per GPU, particles / batches sharded across ranks, one packed
all-reduce per iteration.

Replaces the reference's single DP site, jax.pmap(value_and_grad_fn, in_axes=(None, 0)) followed by a
host-side tree mean (core/trainer.py:44-53): parameters are replicated, every rank evaluates its own
shard, and the whole output (loss scalars + flat gradient [+ ensemble moments]) travels as ONE flat
float32 buffer through torch.distributed.all_reduce(SUM) (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  The payload is <= ~30 KB, i.e. latency-bound; nothing else crosses ranks on the KFP / FP paths.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


@dataclass
class Shard:
    rank: int = 0
    world: int = 1

    @staticmethod
    def current() -> "Shard":
        if dist.is_available() and dist.is_initialized():
            return Shard(dist.get_rank(), dist.get_world_size())
        return Shard(0, 1)

    def bounds(self, n_total: int) -> Tuple[int, int]:
        """Contiguous particle range [lo, hi) of this rank; global ids key the Philox streams, so the
        ensemble is invariant to the number of ranks."""
        base, rem = divmod(n_total, self.world)
        lo = self.rank * base + min(self.rank, rem)
        return lo, lo + base + (1 if self.rank < rem else 0)


def init_from_env(backend: str = "nccl") -> Shard:
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun) if world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=int(os.environ["RANK"]), world_size=world)
    return Shard.current()


def pack(tensors: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, List[Tuple[int, torch.Size]]]:
    flat = torch.cat([t.reshape(-1).float() for t in tensors])
    meta = [(t.numel(), t.shape) for t in tensors]
    return flat, meta


def unpack(flat: torch.Tensor, meta) -> List[torch.Tensor]:
    out, off = [], 0
    for n, shape in meta:
        out.append(flat[off:off + n].view(shape))
        off += n
    return out


def allreduce_sum_packed(tensors: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """One all-reduce(SUM) of all tensors packed into a single flat buffer (no-op for world == 1)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(tensors)
    flat, meta = pack(tensors)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return unpack(flat, meta)


def allreduce_mean_dict(d: Dict[str, torch.Tensor], flat_keys: Sequence[str]) -> Dict[str, torch.Tensor]:
    """The pmap + tree-mean of core/trainer.py:45-52 for a dict of tensors: every entry is averaged over
    ranks, all entries in one collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return d
    keys = list(flat_keys)
    red = allreduce_sum_packed([d[k] for k in keys])
    w = float(dist.get_world_size())
    out = dict(d)
    for k, t in zip(keys, red):
        out[k] = t / w
    return out
