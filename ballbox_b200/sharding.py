"""Host-side logic of the multi-GPU path: which worlds a rank owns and the one
collective of the design (a sum of statistics after the last step).

Independent worlds never interact (the reference has no global state, SURVEY.md
section 8e), so the step path has no collective; `reduce_stats` is the only place
`torch.distributed` is used, with NCCL on GPUs and gloo in the CPU tests."""
from __future__ import annotations

BASE_SEED = 7000


def rank_world_ids(rank: int, world_size: int, worlds_per_rank: int):
    """Global ids of the worlds stepped by `rank` (contiguous block, weak scaling)."""
    if not (0 <= rank < world_size) or worlds_per_rank < 0:
        raise ValueError("bad rank / world_size / worlds_per_rank")
    lo = rank * worlds_per_rank
    return list(range(lo, lo + worlds_per_rank))


def world_seed(global_world_id: int) -> int:
    return BASE_SEED + global_world_id


STAT_KEYS = ("body_steps", "contacts", "worlds", "hash_lo", "hash_hi")


def pack_stats(body_steps: int, contacts: int, worlds: int, hash_sum: int):
    """Statistics vector; the 64-bit checksum of checksums is split so that it
    survives a float64 all-reduce exactly (each half < 2^53 after summing < 2^21 ranks)."""
    h = hash_sum & ((1 << 64) - 1)
    return [float(body_steps), float(contacts), float(worlds), float(h & 0xffffffff), float(h >> 32)]


def reduce_stats(vec, dist=None, device="cpu"):
    """Sum the statistics vector over ranks (identity when not distributed)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(zip(STAT_KEYS, vec))
    import torch
    t = torch.tensor(vec, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return dict(zip(STAT_KEYS, t.tolist()))
