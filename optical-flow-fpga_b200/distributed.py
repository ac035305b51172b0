"""Multi-GPU host logic: one process per GPU (torch.distributed), two ways to split the work.

* batch mode  -- frame pairs are independent (BASELINE configs 3/4): contiguous shards of the
  batch per rank, no data-path collective; an optional all_gather returns the full result.
* row-band mode (single-scale) -- one large frame pair split into row bands.  A band is
  computed on rows [a - halo, b + halo) and cropped to [a, b): flow at a pixel depends on a
  (2 + window)-row neighbourhood only, so with halo = window // 2 + 1 the cropped rows equal
  the full-frame result, and at the true image edges (no extra rows there) the kernel's own
  border rules apply.  No halo exchange is needed because every rank reads its halo rows
  straight from the (replicated or host-resident) input frame.

The collectives are plumbing around the C-ABI calls; `compute` defaults to the CUDA backend
and is injectable so the split / gather logic can be tested on CPU with gloo.
"""

from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) of `n_items` for `rank` (first ranks get the extra)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def band_plan(height: int, world: int, halo: int) -> List[Tuple[int, int, int, int]]:
    """Per rank (out_start, out_stop, read_start, read_stop): output rows and the rows that
    must be read (clipped to the frame) to compute them."""
    plan = []
    for r in range(world):
        a, b = shard_range(height, r, world)
        plan.append((a, b, max(0, a - halo), min(height, b + halo)))
    return plan


def _default_compute(prev, curr, window_size, mode):
    import of_b200

    return of_b200.lk_single_scale_batch(prev, curr, window_size, mode)


def _dist():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist
    return None


def _all_gather_rows(local: np.ndarray, counts: List[int]) -> np.ndarray:
    """Concatenate per-rank arrays along axis 0 (ragged counts allowed)."""
    dist = _dist()
    if dist is None:
        return local
    import torch

    world = dist.get_world_size()
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    mx = max(counts)
    pad = np.zeros((mx,) + local.shape[1:], dtype=local.dtype)
    pad[: local.shape[0]] = local
    t = torch.from_numpy(pad).to(dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return np.concatenate([o.cpu().numpy()[:n] for o, n in zip(out, counts)], axis=0)


def lk_single_scale_sharded(
    prev: np.ndarray,
    curr: np.ndarray,
    window_size: int = 5,
    mode: Optional[int] = None,
    gather: bool = True,
    compute: Callable = _default_compute,
):
    """Batch mode: every rank holds (or can read) the whole [B, H, W] stacks, computes its
    shard and, if `gather`, gets the full (u, v) back.  Without an initialised process group
    this is a plain single-process call."""
    dist = _dist()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
    b = prev.shape[0]
    s, e = shard_range(b, rank, world)
    if e > s:
        u, v = compute(prev[s:e], curr[s:e], window_size, mode)
    else:
        u = np.zeros((0,) + prev.shape[1:], np.float32)
        v = u.copy()
    if not gather or world == 1:
        return u, v
    counts = [shard_range(b, r, world)[1] - shard_range(b, r, world)[0] for r in range(world)]
    return _all_gather_rows(u, counts), _all_gather_rows(v, counts)


def lk_single_scale_rowbands(
    frame_prev: np.ndarray,
    frame_curr: np.ndarray,
    window_size: int = 5,
    mode: Optional[int] = None,
    gather: bool = True,
    compute: Callable = _default_compute,
):
    """Row-band mode for one [H, W] frame pair: rank r computes its band (with halo rows read
    from the input) and, if `gather`, every rank gets the full flow field."""
    dist = _dist()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
    h = frame_prev.shape[0]
    halo = window_size // 2 + 1
    plan = band_plan(h, world, halo)
    a, b, ra, rb = plan[rank]
    if b > a:
        u, v = compute(frame_prev[None, ra:rb], frame_curr[None, ra:rb], window_size, mode)
        u, v = u[0][a - ra : b - ra], v[0][a - ra : b - ra]
    else:
        u = np.zeros((0, frame_prev.shape[1]), np.float32)
        v = u.copy()
    if not gather or world == 1:
        return u, v
    counts = [p[1] - p[0] for p in plan]
    return _all_gather_rows(np.ascontiguousarray(u), counts), _all_gather_rows(np.ascontiguousarray(v), counts)
