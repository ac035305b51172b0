"""Multi-GPU host logic: one process per GPU (torch.distributed), two ways to split the work.

* batch mode  -- frame pairs are independent (BASELINE configs 3/4): contiguous shards of the
  batch per rank, no data-path collective; an optional all_gather returns the full result.
* row-band mode (single-scale) -- one large frame pair split into row bands.  A band is
  computed on rows [a - halo, b + halo) and cropped to [a, b): flow at a pixel depends on a
  (2 + window)-row neighbourhood only, so with halo = window // 2 + 1 the cropped rows equal
  the full-frame result, and at the true image edges (no extra rows there) the kernel's own
  border rules apply.  No halo exchange is needed because every rank reads its halo rows
  straight from the (replicated or host-resident) input frame.

* row-band mode (pyramidal) -- two drivers with the same row bookkeeping and the same bits:
  `lk_pyramidal_rowbands` (Python loop over the `_dev` building blocks, NCCL / gloo collectives,
  injectable backend: the one the CPU tests drive) and `PeerRowbands` (the native driver:
  `of_rowband_run`, collectives by peer stores inside the kernels, no NCCL in the data path).

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


# ======================================================================================
# Row-band mode for the PYRAMIDAL path (BASELINE config 5: one very large frame pair)
# ======================================================================================
#
# Every rank keeps full-size level images and flow planes but computes only its band of rows.
# Inside a level there is NO halo exchange: iteration i of I is computed on the band extended
# by band_grow(window) * (I - 1 - i) rows on both sides (a refinement iteration reads flow_in
# window // 2 + 1 rows beyond the rows it writes; one more row is left for the even alignment of the
# band start, rounded up to even: 4 rows for windows up to 5, 6 for window 7), so
# after the last iteration exactly the owned rows are still valid.  Communication happens
#   * once per iteration: all-reduce of (sum|du|, sum|dv|) over the owned rows -- 16 bytes,
#     needed for the reference's global early exit (lucas_kanade_pyramidal.py:213-223);
#   * once per level: all-gather of the owned flow rows, which feeds the next level's
#     upsampling (its taps and the next level's extended band reach into neighbour bands)
#     and, at the finest level, is the final gather.
# The warp gather has unbounded reach, so the current frame's pyramid is replicated (every
# rank builds both pyramids itself; they cost far less than the iterations).
#
# Backends: `CudaBackend` (torch CUDA tensors + the `_dev` C-ABI calls) is the product path;
# the split / collective logic is backend-agnostic so that tests can drive it on CPU tensors.

def band_grow(window_size: int) -> int:
    """Rows by which a band's computed range shrinks per iteration (csrc/of_rowband.inl: rb_grow)."""
    return max(4, (window_size // 2 + 3) & ~1)


class SingleProcessComm:
    rank, world = 0, 1

    def all_reduce_sum(self, t):
        return t

    def all_gather_rows(self, local_rows, counts):
        return local_rows


class TorchDistComm:
    """torch.distributed default group: NCCL for CUDA tensors, gloo for CPU tensors."""

    def __init__(self):
        import torch.distributed as dist

        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def all_reduce_sum(self, t):
        self.dist.all_reduce(t)
        return t

    def all_gather_rows(self, local_rows, counts):
        import torch

        mx = max(counts)
        pad = torch.zeros((mx,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
        pad[: local_rows.shape[0]] = local_rows
        out = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(out, pad)
        return torch.cat([o[:n] for o, n in zip(out, counts)], dim=0)


class ThreadComm:
    """R ranks as threads of one process (tests: emulate a multi-rank job on one device)."""

    def __init__(self, world: int):
        import threading

        self.world = world
        self._barrier = threading.Barrier(world)
        self._slots = [None] * world

    def view(self, rank: int):
        parent = self

        class _View:
            world = parent.world

            def __init__(self, r):
                self.rank = r

            def _exchange(self, item):
                parent._slots[self.rank] = item
                parent._barrier.wait()
                items = list(parent._slots)
                parent._barrier.wait()
                return items

            def all_reduce_sum(self, t):
                items = self._exchange(t.clone())
                total = items[0].clone()
                for it in items[1:]:
                    total += it.to(total.device)
                return total

            def all_gather_rows(self, local_rows, counts):
                import torch

                items = self._exchange(local_rows.clone())
                return torch.cat([it.to(local_rows.device) for it in items], dim=0)

        return _View(rank)


class CudaBackend:
    """Device arrays are torch CUDA tensors; arithmetic is the C-ABI `_dev` entry points."""

    def __init__(self, device=None):
        import torch

        import of_b200

        self.torch, self.ofb = torch, of_b200
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self._ws = None

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def from_host(self, a):
        if isinstance(a, self.torch.Tensor):  # already device-resident
            return a.to(device=self.device, dtype=self.torch.float32).contiguous()
        return self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device)

    def to_host(self, t):
        return t.cpu().numpy()

    def empty(self, h, w):
        return self.torch.empty((h, w), dtype=self.torch.float32, device=self.device)

    def zeros(self, h, w):
        return self.torch.zeros((h, w), dtype=self.torch.float32, device=self.device)

    def pyramid_down(self, img, lo=0, hi=None):
        """Rows [lo, hi) of the next coarser level (the other rows of the result are undefined)."""
        h, w = img.shape
        out = self.empty(int(h * 0.5), int(w * 0.5))
        self.ofb.pyramid_down_dev(img.data_ptr(), out.data_ptr(), 1, h, w, out.shape[0], out.shape[1], self._stream(),
                                  2.0, lo, out.shape[0] if hi is None else hi, getattr(self, "mode", 0))
        return out

    # ---- device-side iteration control (no host round trip inside a level) ----------------
    def int_zeros(self, n):
        return self.torch.zeros(n, dtype=self.torch.int32, device=self.device)

    def f32_zeros(self, n):
        return self.torch.zeros(n, dtype=self.torch.float32, device=self.device)

    def refine_pingpong(self, prev, curr, f0u, f0v, f1u, f1v, sel, done, window, mode, lo, hi, own_lo, own_hi):
        h, w = prev.shape
        need = self.ofb.lk_refine_workspace_bytes(1, h, w)
        if self._ws is None or self._ws.numel() < need:
            self._ws = self.torch.empty(need, dtype=self.torch.uint8, device=self.device)
        sums = self.torch.zeros(2, dtype=self.torch.float64, device=self.device)
        self.ofb.lk_refine_pingpong_dev(prev.data_ptr(), curr.data_ptr(), f0u.data_ptr(), f0v.data_ptr(), f1u.data_ptr(),
                                        f1v.data_ptr(), sel.data_ptr(), done.data_ptr(), 1, h, w, window, mode, lo, hi,
                                        own_lo, own_hi, sums.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                                        self._stream())
        return sums

    def select_rows(self, sel, buf0, buf1, r0, r1):
        """Plane whose rows [r0, r1) are those of buf1 if sel != 0 else buf0 (no host read)."""
        out = buf0 if (r0, r1) == (0, buf0.shape[0]) else self.torch.empty_like(buf0)
        out[r0:r1] = self.torch.where(sel[0] > 0, buf1[r0:r1], buf0[r0:r1])
        return out

    def convergence_update(self, sums, n_pixels, sel, done, iters_t, resid_t, max_iters, it):
        self.ofb.lk_convergence_update_dev(sums.data_ptr(), 1, n_pixels, sel.data_ptr(), done.data_ptr(),
                                           iters_t.data_ptr(), resid_t.data_ptr(), max_iters, it, self._stream())

    def upsample(self, cu, cv, th, tw, lo, hi):
        u, v = self.empty(th, tw), self.empty(th, tw)
        self.ofb.upsample_flow_dev(cu.data_ptr(), cv.data_ptr(), u.data_ptr(), v.data_ptr(), 1, cu.shape[0], cu.shape[1],
                                   th, tw, lo, hi, self._stream())
        return u, v

    def refine(self, prev, curr, fin_u, fin_v, fout_u, fout_v, window, mode, lo, hi, own_lo, own_hi):
        h, w = prev.shape
        need = self.ofb.lk_refine_workspace_bytes(1, h, w)
        if self._ws is None or self._ws.numel() < need:
            self._ws = self.torch.empty(need, dtype=self.torch.uint8, device=self.device)
        sums = self.torch.zeros(2, dtype=self.torch.float64, device=self.device)
        self.ofb.lk_refine_dev(prev.data_ptr(), curr.data_ptr(), fin_u.data_ptr(), fin_v.data_ptr(), fout_u.data_ptr(),
                               fout_v.data_ptr(), 1, h, w, window, mode, lo, hi, own_lo, own_hi, sums.data_ptr(),
                               self._ws.data_ptr(), self._ws.numel(), self._stream())
        return sums

    def zero_sums(self):
        return self.torch.zeros(2, dtype=self.torch.float64, device=self.device)


def lk_pyramidal_rowbands(
    frame_prev: np.ndarray,
    frame_curr: np.ndarray,
    num_levels: int = 3,
    window_size: int = 5,
    num_iterations: int = 3,
    mode: Optional[int] = None,
    comm=None,
    backend=None,
    trace: Optional[list] = None,
    to_host: bool = True,
):
    """Pyramidal LK of ONE [H, W] frame pair with the rows of every level split over the ranks.
    Every rank returns the full (u, v) -- NumPy arrays, or the backend's device arrays with
    to_host=False (frames may be passed as device arrays too).  Results equal the single-GPU
    path bit for bit (same kernels, same row pairing); the early-exit decision uses the
    all-reduced residual sums."""
    if comm is None:
        comm = TorchDistComm() if _dist() is not None else SingleProcessComm()
    if backend is None:
        backend = CudaBackend()
    if mode is None:
        import of_b200

        mode = of_b200.default_mode()
    rank, world = comm.rank, comm.world
    backend.mode = mode  # the fast drivers' pyramid uses the fast filter (same bits as the single-GPU fast path)
    GROW = band_grow(int(window_size))
    iters = int(num_iterations)

    def gather(arr, h):
        """Full plane from the ranks' owned rows (every rank computed rows shard_range(h) only)."""
        if world == 1:
            return arr
        a_, b_ = shard_range(h, rank, world)
        counts = [shard_range(h, r, world)[1] - shard_range(h, r, world)[0] for r in range(world)]
        return comm.all_gather_rows(arr[a_:b_].contiguous(), counts)

    # Gaussian pyramids: every rank smooths / decimates its rows of each level, then the level
    # is all-gathered (the next level's filter and the warp gather reach into other bands)
    lv_prev = [backend.from_host(frame_prev)]
    lv_curr = [backend.from_host(frame_curr)]
    for _ in range(1, int(num_levels)):
        oh = int(lv_prev[0].shape[0] * 0.5)
        a, b = shard_range(oh, rank, world)
        if b > a:
            nxt_p, nxt_c = backend.pyramid_down(lv_prev[0], a, b), backend.pyramid_down(lv_curr[0], a, b)
        else:
            nxt_p = nxt_c = backend.empty(oh, int(lv_prev[0].shape[1] * 0.5))
        lv_prev.insert(0, gather(nxt_p, oh))
        lv_curr.insert(0, gather(nxt_c, oh))

    device_control = hasattr(backend, "refine_pingpong")
    want_trace_sync = trace is not None
    h, w = lv_prev[0].shape
    flow_u, flow_v = backend.zeros(h, w), backend.zeros(h, w)
    for level, (img_prev, img_curr) in enumerate(zip(lv_prev, lv_curr)):
        h, w = img_prev.shape
        a, b = shard_range(h, rank, world)
        reach = GROW * max(iters, 1) + GROW
        if level > 0:
            lo, hi = max(0, a - reach), min(h, b + reach)
            if b > a:
                flow_u, flow_v = backend.upsample(flow_u, flow_v, h, w, lo, hi)
            else:
                flow_u, flow_v = backend.empty(h, w), backend.empty(h, w)
        out_u, out_v = backend.empty(h, w), backend.empty(h, w)
        n = float(h) * float(w)

        def rows_of(it):
            ext = GROW * (iters - 1 - it)
            lo_ = max(0, a - ext)
            return lo_ - (lo_ & 1), min(h, b + ext)  # even start: rows pair up identically on every rank

        if device_control:
            # all iterations are enqueued back to back; the convergence test, the ping-pong flip
            # and the skipping of iterations after convergence happen on the device
            sel, done, iters_t = backend.int_zeros(1), backend.int_zeros(1), backend.int_zeros(1)
            resid_t = backend.f32_zeros(2 * max(iters, 1))
            for it in range(iters):
                lo, hi = rows_of(it)
                if b > a:
                    sums = backend.refine_pingpong(img_prev, img_curr, flow_u, flow_v, out_u, out_v, sel, done,
                                                   window_size, mode, lo, hi, a, b)
                else:
                    sums = backend.zero_sums()
                sums = comm.all_reduce_sum(sums)
                backend.convergence_update(sums, n, sel, done, iters_t, resid_t, iters, it)
            if iters > 0:
                if hasattr(backend, "select_rows") and not want_trace_sync:
                    # pick the current ping-pong buffer on the device (owned rows only when they
                    # are gathered afterwards): the whole call never waits for the GPU, so it
                    # can be captured in a CUDA graph
                    r0, r1 = (a, b) if world > 1 else (0, h)
                    flow_u = backend.select_rows(sel, flow_u, out_u, r0, r1)
                    flow_v = backend.select_rows(sel, flow_v, out_v, r0, r1)
                elif int(sel.cpu()[0]) == 1:  # one host read per level
                    flow_u, flow_v = out_u, out_v
            if trace is not None:
                res = resid_t.cpu().numpy()
                for it in range(int(iters_t.cpu()[0])):
                    trace.append((level, it, float(res[2 * it]), float(res[2 * it + 1])))
        else:
            for it in range(iters):
                lo, hi = rows_of(it)
                if b > a:
                    sums = backend.refine(img_prev, img_curr, flow_u, flow_v, out_u, out_v, window_size, mode, lo, hi, a, b)
                else:
                    sums = backend.zero_sums()
                sums = comm.all_reduce_sum(sums)
                s = sums.detach().cpu().numpy() if hasattr(sums, "detach") else np.asarray(sums)
                mean_du, mean_dv = np.float32(s[0] / n), np.float32(s[1] / n)
                if trace is not None:
                    trace.append((level, it, float(mean_du), float(mean_dv)))
                flow_u, out_u = out_u, flow_u
                flow_v, out_v = out_v, flow_v
                if mean_du < np.float32(0.01) and mean_dv < np.float32(0.01):
                    break
        # level done: rows [a, b) are final on this rank -> every rank gets the whole plane
        flow_u, flow_v = gather(flow_u, h), gather(flow_v, h)
    if not to_host:
        return flow_u, flow_v
    return backend.to_host(flow_u), backend.to_host(flow_v)



# ======================================================================================
# Row-band mode, native driver: peer-memory collectives instead of NCCL + Python loop
# ======================================================================================
class _CudaArrayView:
    """Zero-copy view of device memory for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr: int, shape, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerRowbands:
    """lucas_kanade_pyramidal for one frame geometry with the rows of every level split over the
    ranks of the default process group (one process per GPU of one NVLink domain).

    Setup (once): every rank creates its arena and the 64-byte CUDA IPC handles are exchanged with
    `all_gather_object`.  After that torch.distributed is not used any more: `run` is one C call
    (`of_rowband_run`) that enqueues the whole coarse-to-fine computation; pyramid rows, final flow
    rows and the residual sums of every iteration move between the GPUs by peer stores and flag
    words inside the kernels (csrc/peer.cu).  Without a process group it degenerates to one rank.
    """

    def __init__(self, height: int, width: int, num_levels: int = 3, window_size: int = 5, num_iterations: int = 3,
                 mode: Optional[int] = None, device=None):
        import torch

        import of_b200

        self.torch, self.ofb = torch, of_b200
        dist = _dist()
        self.rank, self.world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.shape = (int(height), int(width))
        if mode is None:
            mode = of_b200.default_mode()
        self.ctx = of_b200.RowbandContext(self.rank, self.world, height, width, num_levels, window_size, num_iterations, mode)
        import os

        if os.environ.get("OF_B200_ROWBAND_REPL_PX"):  # A/B runs: 0 = split every level into row bands
            self.ctx.set_replicate_pixels(int(os.environ["OF_B200_ROWBAND_REPL_PX"]))
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, self.ctx.ipc_handle())
            self.ctx.open_peers_ipc(handles)
            dist.barrier()  # nobody starts storing into a peer before every rank has mapped every arena
        pu, pv = self.ctx.result_ptrs()
        self.u = torch.as_tensor(_CudaArrayView(pu, self.shape), device=self.device)
        self.v = torch.as_tensor(_CudaArrayView(pv, self.shape), device=self.device)

    def run(self, prev, curr):
        """prev, curr: full [H, W] float32 CUDA tensors on this rank's device (the same frames on
        every rank).  Returns (u, v): views of the arena planes that hold the gathered result on
        every rank once the current stream has finished (overwritten by the next run)."""
        torch = self.torch
        for t in (prev, curr):
            if tuple(t.shape) != self.shape or t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
                raise ValueError("frames must be contiguous float32 CUDA tensors of the planned shape")
        self.ctx.run(prev.data_ptr(), curr.data_ptr(), None, None, torch.cuda.current_stream(self.device).cuda_stream)
        return self.u, self.v

    def trace(self):
        """(iters_executed[levels], residuals[levels, iterations, 2]); raises if a peer timed out."""
        iters, resid, err = self.ctx.trace(self.torch.cuda.current_stream(self.device).cuda_stream)
        if err:
            raise self.ofb.OFBackendError("row-band run: a peer rank did not answer within the time-out")
        return iters, resid

    def close(self):
        self.u = self.v = None
        dist = _dist()
        if self.world > 1 and dist is not None:
            # the arena is mapped by every peer: nobody frees it while another rank may still store into it
            self.torch.cuda.synchronize(self.device)
            dist.barrier()
        self.ctx.close()


class PeerRowbandLanes:
    """A stream of frame pairs through `lanes` PeerRowbands plans at once.

    Frame pairs are independent, and at 8K most kernels of the coarse pyramid levels are bound by
    launch latency and by the cross-GPU flag round trip, not by the GPU.  Running `lanes` pairs
    concurrently -- each on its own CUDA stream with its own arena and flag words -- lets one pair's
    latency-bound kernels overlap with another pair's (12 % on 2 GPUs, 14 % on 4 at two lanes).
    `run_batch` forks the lanes off the current stream and joins them back, so it can be captured in
    a CUDA graph and replayed (the collectives' sequence numbers come from device-side counters)."""

    def __init__(self, height: int, width: int, num_levels: int = 3, window_size: int = 5, num_iterations: int = 3,
                 mode: Optional[int] = None, lanes: int = 2, device=None):
        import torch

        self.torch = torch
        self.plans = [PeerRowbands(height, width, num_levels, window_size, num_iterations, mode, device)
                      for _ in range(max(1, int(lanes)))]
        self.device = self.plans[0].device
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.plans]

    def run_batch(self, prev, curr, u, v):
        """prev, curr, u, v: [B, H, W] float32 CUDA tensors; the gathered flow of pair b lands in u[b], v[b]
        on every rank.  Only enqueues work."""
        torch = self.torch
        main = torch.cuda.current_stream(self.device)
        for s_ in self.streams:  # fork
            s_.wait_stream(main)
        for b in range(prev.shape[0]):
            k = b % len(self.plans)
            self.plans[k].ctx.run(prev[b].data_ptr(), curr[b].data_ptr(), u[b].data_ptr(), v[b].data_ptr(),
                                  self.streams[k].cuda_stream)
        for s_ in self.streams:  # join
            main.wait_stream(s_)

    def trace(self):
        """Waits for the device; raises if any lane saw a peer time out."""
        self.torch.cuda.synchronize(self.device)
        return [p.trace() for p in self.plans]

    def close(self):
        for p in self.plans:
            p.close()


def lk_pyramidal_rowbands_peer(frame_prev, frame_curr, num_levels=3, window_size=5, num_iterations=3, mode=None):
    """Convenience form of PeerRowbands for one call with NumPy frames: (u, v) as NumPy arrays."""
    import torch

    plan = PeerRowbands(frame_prev.shape[0], frame_prev.shape[1], num_levels, window_size, num_iterations, mode)
    try:
        dev = plan.device
        p = torch.from_numpy(np.ascontiguousarray(frame_prev, dtype=np.float32)).to(dev)
        c = torch.from_numpy(np.ascontiguousarray(frame_curr, dtype=np.float32)).to(dev)
        u, v = plan.run(p, c)
        plan.trace()  # waits for the stream and checks the error word
        return u.cpu().numpy(), v.cpu().numpy()
    finally:
        plan.close()
