"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding, row bands with
halo rows, ragged gathers.  The per-shard compute is the CPU oracle here (injected); on the
GPU box the same functions call the CUDA backend."""

import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import BACKEND_DIR, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_compute(prev, curr, window_size, mode):
    from oracle import lk_float_oracle as orc

    us, vs = zip(*(orc.lucas_kanade_single_scale(p, c, window_size) for p, c in zip(prev, curr)))
    return np.stack(us), np.stack(vs)


def _worker(rank, world, port, out_dir):
    for p in (str(ROOT), str(BACKEND_DIR)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist

    import distributed as ofd
    import synthetic

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prev, curr, _ = synthetic.make_pairs_numpy(5, 40, 56, seed=4)  # 5 pairs over 2 ranks: ragged
        u, v = ofd.lk_single_scale_sharded(prev, curr, 5, None, gather=True, compute=_oracle_compute)
        np.save(os.path.join(out_dir, f"batch_u_{rank}.npy"), u)
        np.save(os.path.join(out_dir, f"batch_v_{rank}.npy"), v)
        fp, fc, _ = synthetic.make_pairs_numpy(1, 75, 64, seed=5)  # odd height: uneven bands
        u, v = ofd.lk_single_scale_rowbands(fp[0], fc[0], 5, None, gather=True, compute=_oracle_compute)
        np.save(os.path.join(out_dir, f"band_u_{rank}.npy"), u)
        np.save(os.path.join(out_dir, f"band_v_{rank}.npy"), v)
        ul, _ = ofd.lk_single_scale_sharded(prev, curr, 5, None, gather=False, compute=_oracle_compute)
        np.save(os.path.join(out_dir, f"local_u_{rank}.npy"), ul)
    finally:
        dist.destroy_process_group()


def test_shard_range_and_band_plan():
    sys.path.insert(0, str(BACKEND_DIR))
    import distributed as ofd

    for n in (0, 1, 5, 8, 513):
        for world in (1, 2, 3, 8):
            spans = [ofd.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    plan = ofd.band_plan(4320, 8, 3)
    assert plan[0] == (0, 540, 0, 543) and plan[-1] == (3780, 4320, 3777, 4320)
    assert plan[3] == (1620, 2160, 1617, 2163)
    with pytest.raises(ValueError):
        ofd.shard_range(4, 2, 2)


def test_world_size_2_gloo(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, str(BACKEND_DIR))
    import synthetic

    prev, curr, _ = synthetic.make_pairs_numpy(5, 40, 56, seed=4)
    uo, vo = _oracle_compute(prev, curr, 5, None)
    fp, fc, _ = synthetic.make_pairs_numpy(1, 75, 64, seed=5)
    ubo, vbo = _oracle_compute(fp, fc, 5, None)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"batch_u_{r}.npy"), uo)
        assert np.array_equal(np.load(tmp_path / f"batch_v_{r}.npy"), vo)
        # row bands with halo rows reproduce the full-frame result exactly, seams included
        assert np.array_equal(np.load(tmp_path / f"band_u_{r}.npy"), ubo[0])
        assert np.array_equal(np.load(tmp_path / f"band_v_{r}.npy"), vbo[0])
    assert np.load(tmp_path / "local_u_0.npy").shape[0] == 3
    assert np.load(tmp_path / "local_u_1.npy").shape[0] == 2


# ---------------------------------------------------------------------------------------
# row-band PYRAMIDAL driver: split / overlap / collective logic on CPU tensors
# ---------------------------------------------------------------------------------------
class OracleBackend:
    """Stand-in for CudaBackend on CPU tensors: every op is the oracle applied to the whole
    plane, then only the requested rows are kept.  Rows a rank must not rely on are NaN, so any
    dependence on a row outside the validity argument of the driver poisons the result."""

    def __init__(self):
        import torch

        self.torch = torch

    def from_host(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32).copy())

    def to_host(self, t):
        return t.numpy().copy()

    def empty(self, h, w):
        return self.torch.full((h, w), float("nan"), dtype=self.torch.float32)

    def zeros(self, h, w):
        return self.torch.zeros((h, w), dtype=self.torch.float32)

    def zero_sums(self):
        return self.torch.zeros(2, dtype=self.torch.float64)

    def pyramid_down(self, img, lo=0, hi=None):
        from oracle import lk_float_oracle as orc

        full = orc.build_gaussian_pyramid(img.numpy(), 2)[0]
        out = self.empty(*full.shape)
        hi = full.shape[0] if hi is None else hi
        out[lo:hi] = self.torch.from_numpy(full[lo:hi].copy())
        return out

    def upsample(self, cu, cv, th, tw, lo, hi):
        from oracle import lk_float_oracle as orc

        u, v = orc.upsample_flow(cu.numpy(), cv.numpy(), (th, tw))
        ou, ov = self.empty(th, tw), self.empty(th, tw)
        ou[lo:hi] = self.torch.from_numpy(u[lo:hi].copy())
        ov[lo:hi] = self.torch.from_numpy(v[lo:hi].copy())
        return ou, ov

    def refine(self, prev, curr, fin_u, fin_v, fout_u, fout_v, window, mode, lo, hi, own_lo, own_hi):
        from oracle import lk_float_oracle as orc

        assert lo % 2 == 0
        with np.errstate(all="ignore"):
            warped = orc.warp_image(curr.numpy(), fin_u.numpy(), fin_v.numpy())
            du, dv = orc.lucas_kanade_single_scale(prev.numpy(), warped, window)
        fout_u[lo:hi] = self.torch.from_numpy((fin_u.numpy()[lo:hi] + du[lo:hi]).astype(np.float32))
        fout_v[lo:hi] = self.torch.from_numpy((fin_v.numpy()[lo:hi] + dv[lo:hi]).astype(np.float32))
        s = np.array([np.abs(du[own_lo:own_hi]).astype(np.float64).sum(), np.abs(dv[own_lo:own_hi]).astype(np.float64).sum()])
        return self.torch.from_numpy(s)


def _pyr_case():
    from scipy.ndimage import gaussian_filter, shift

    rng = np.random.default_rng(21)
    p = gaussian_filter((rng.random((97, 72)) * 255).astype(np.float32), 1.0)
    c = shift(p, (0.8, -1.1), order=1, mode="nearest").astype(np.float32)
    return p, c


def _pyr_thread(rank, comm, out, args):
    sys.path.insert(0, str(BACKEND_DIR))
    import distributed as ofd

    p, c, levels, iters, window = args
    out[rank] = ofd.lk_pyramidal_rowbands(p, c, levels, window, iters, mode=0, comm=comm.view(rank), backend=OracleBackend())


@pytest.mark.parametrize("world,levels,iters,window", [(3, 3, 3, 5), (2, 2, 1, 5), (4, 3, 2, 5), (3, 3, 3, 7), (4, 2, 2, 3)])
def test_pyramidal_rowbands_thread_ranks_equal_oracle(world, levels, iters, window):
    import threading

    sys.path.insert(0, str(BACKEND_DIR))
    import distributed as ofd
    from oracle import lk_float_oracle as orc

    p, c = _pyr_case()
    uo, vo = orc.lucas_kanade_pyramidal(p, c, levels, window, iters)
    comm = ofd.ThreadComm(world)
    out = [None] * world
    threads = [threading.Thread(target=_pyr_thread, args=(r, comm, out, (p, c, levels, iters, window))) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for r in range(world):
        u, v = out[r]
        assert not np.isnan(u).any() and not np.isnan(v).any(), "a rank relied on rows it does not hold"
        assert np.array_equal(u.view(np.uint32), uo.view(np.uint32)), f"rank {r} u"
        assert np.array_equal(v.view(np.uint32), vo.view(np.uint32)), f"rank {r} v"


def _pyr_gloo_worker(rank, world, port, out_dir):
    for p in (str(ROOT), str(BACKEND_DIR), str(ROOT / "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist

    import distributed as ofd

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p, c = _pyr_case()
        trace = []
        u, v = ofd.lk_pyramidal_rowbands(p, c, 3, 5, 3, mode=0, backend=OracleBackend(), trace=trace)
        np.save(os.path.join(out_dir, f"pyr_u_{rank}.npy"), u)
        np.save(os.path.join(out_dir, f"pyr_v_{rank}.npy"), v)
        np.save(os.path.join(out_dir, f"pyr_trace_{rank}.npy"), np.array(trace))
    finally:
        dist.destroy_process_group()


def test_pyramidal_rowbands_world_size_2_gloo(tmp_path):
    from oracle import lk_float_oracle as orc

    world, port = 2, _free_port()
    mp.spawn(_pyr_gloo_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    p, c = _pyr_case()
    ref_trace = []
    uo, vo = orc.lucas_kanade_pyramidal(p, c, 3, 5, 3, trace=ref_trace)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"pyr_u_{r}.npy").view(np.uint32), uo.view(np.uint32))
        assert np.array_equal(np.load(tmp_path / f"pyr_v_{r}.npy").view(np.uint32), vo.view(np.uint32))
        tr = np.load(tmp_path / f"pyr_trace_{r}.npy")
        assert tr.shape[0] == len(ref_trace)  # same iterations executed (global early-exit decisions)
        assert np.allclose(tr[:, 2:], np.array(ref_trace)[:, 2:], rtol=1e-5)
