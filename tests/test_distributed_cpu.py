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
