"""Multi-GPU plumbing: one process per GPU (torchrun), trial years sharded contiguously, the small
integer accumulators combined with one all-reduce (NCCL over NVLink on the B200 box, gloo in the CPU
tests).  SURVEY.md section 8e: the path has no data-path collective -- years are independent because
every chain owns its Philox streams -- so the only exchange is this final sum of a few integers.

torch is used for the process group only; the compute path is libpsra_b200.so.
"""
from __future__ import annotations

from typing import Dict, Tuple

_KEYS = ("years", "sum_lol_hours", "sum_ens_fp", "sum_entries", "years_with_loss", "sum_lol_sq", "events")
_LIMB = 32


def shard_range(total: int, rank: int, world: int, multiple: int = 1) -> Tuple[int, int]:
    """Contiguous [start, stop) of `total` units for `rank`, boundaries aligned to `multiple`
    (years_per_chain); the union over ranks is exactly [0, total)."""
    if total % multiple:
        raise ValueError("total must be a multiple of the chain length")
    chunks = total // multiple
    base, rem = divmod(chunks, world)
    start = rank * base + min(rank, rem)
    stop = start + base + (1 if rank < rem else 0)
    return start * multiple, stop * multiple


def pack_raw(raw: Dict[str, int]):
    """Exact integer accumulators -> list of int64-safe limbs (the 128-bit sum of ENS^2 is split in
    32-bit limbs so that the element-wise sum over <= 2^31 ranks cannot overflow int64)."""
    vals = [int(raw.get(k, 0)) for k in _KEYS]
    e2 = int(raw.get("sum_ens_sq", 0))
    limbs = [(e2 >> (_LIMB * i)) & ((1 << _LIMB) - 1) for i in range(4)]
    return vals + limbs


def unpack_raw(vals) -> Dict[str, int]:
    vals = [int(v) for v in vals]
    raw = {k: vals[i] for i, k in enumerate(_KEYS)}
    n = len(_KEYS)
    raw["sum_ens_sq"] = sum(vals[n + i] << (_LIMB * i) for i in range(4))
    return raw


def allreduce_raw(raw: Dict[str, int], device=None, group=None) -> Dict[str, int]:
    """Sum the accumulators of all ranks (torch.distributed, int64)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(pack_raw(raw), dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return unpack_raw(t.cpu().tolist())


def allreduce_counts(counts, device=None, group=None):
    """Element-wise sum over ranks of an integer count vector (per-hour failure counts of
    tail_risk.jl:81,88 / psra_seq_outputs.fail_count, histogram bins of psra_tail): SURVEY.md section 8e."""
    import numpy as np
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.ascontiguousarray(counts).astype(np.int64), device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def gather_years(values, device=None, group=None):
    """Per-year (or per-group) vectors of the ranks' contiguous shards -> the vector of the whole experiment, in rank
    = year order, on every rank.  The shards may differ in length by one chain (shard_range)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    v = np.ascontiguousarray(values)
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return v.copy()
    world = dist.get_world_size(group)
    n = torch.tensor([v.size], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(x.item()) for x in sizes]
    width = max(sizes)
    pad = torch.zeros(width, dtype=torch.int64, device=device)
    pad[:v.size] = torch.as_tensor(v.astype(np.int64), device=device)
    parts = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return np.concatenate([p[:k].cpu().numpy() for p, k in zip(parts, sizes)]).astype(v.dtype)


def merged_history(group_lol, group: int, device=None, group_pg=None):
    """PowerSystemAdequacy.jl:263-265 across ranks: the ranks' per-group LOL-hour sums (psra_seq_outputs.group_lol,
    `group` years each, shards aligned to the group) -> running mean of the LOL hours every `group` years of the
    whole experiment."""
    import numpy as np
    g = gather_years(group_lol, device=device, group=group_pg).astype(np.float64)
    return np.cumsum(g) / (group * np.arange(1, g.size + 1, dtype=np.float64))


def tail_all_ranks(engine, ens_fp_local, alphas=(0.95, 0.99), device=None, group=None):
    """VaR / CVaR over the years of all ranks: the per-year ENS shards are gathered (8 B per year)
    and every rank runs the exact device reduction (psra_tail) on the whole vector."""
    x = gather_years(ens_fp_local, device=device, group=group)
    return engine.tail(x, alphas=alphas)
