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
