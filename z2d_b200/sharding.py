"""Multi-GPU partitioning of the path: a batch of independent scenes shards across ranks with no
data-path collective -- scene s belongs to rank `s mod world`; every rank owns its scenes' surfaces and command
batches.  torch.distributed is used only for (a) the timing reduction (max over ranks, sum of work units) and
(b) gathering 32-byte per-scene checksums to every rank so a sharded run can be compared with an unsharded one.
Backend-agnostic (NCCL on the GPUs, gloo in the CPU tests)."""
import hashlib

import numpy as np

BASE_SEED_C2 = 0x7A326402
BASE_SEED_C5 = 0x7A326405  # seed = base + scene


def scenes_of_rank(n_scenes, world, rank):
    """Scene indices owned by `rank` (scene s -> rank s mod world)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_scenes, world))


def scene_seed(base, scene):
    return (base + scene) & 0xFFFFFFFFFFFFFFFF


def surface_checksum(raw):
    """32-byte digest of a surface's raw bytes (what a rank contributes instead of the pixels)."""
    return hashlib.sha256(np.ascontiguousarray(raw).tobytes()).digest()


def reduce_timing(ms, units, device=None, dist=None):
    """(max over ranks of each time in `ms`, sum over ranks of each count in `units`)."""
    import torch
    t = torch.tensor([float(v) for v in ms], dtype=torch.float64, device=device)
    u = torch.tensor([float(v) for v in units], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return [float(v) for v in t], [float(v) for v in u]


def gather_checksums(local, n_scenes, dist=None):
    """local: {scene index: 32-byte digest} for this rank's scenes -> list of all n_scenes digests (every rank)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        merged = dict(local)
    else:
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, local)
        merged = {}
        for p in parts:
            overlap = merged.keys() & p.keys()
            if overlap:
                raise RuntimeError(f"scenes rendered by more than one rank: {sorted(overlap)}")
            merged.update(p)
    missing = [s for s in range(n_scenes) if s not in merged]
    if missing:
        raise RuntimeError(f"scenes not rendered by any rank: {missing}")
    return [merged[s] for s in range(n_scenes)]
