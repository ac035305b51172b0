#!/usr/bin/env python3
"""Real multi-GPU check of the row-band pyramidal mode (NCCL), run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/run_rowband_nccl.py [--height 4320 --width 7680 --levels 5 --iters 10]

Every rank computes its row bands; rank 0 compares the gathered flow with the single-GPU
result of the same frames (bit for bit) and prints timings."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "optical-flow-fpga_b200")):
    sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist

    import distributed as ofd
    import of_b200
    import synthetic

    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--levels", type=int, default=3)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--mode", choices=["exact", "fast"], default="fast")
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--timeline", default="", help="write rank 0's kernel timeline of one more run (CUPTI through torch.profiler) "
                    "to this CSV: analysis only, never a bench number")
    ap.add_argument("--driver", choices=["nccl", "peer"], default="peer",
                    help="nccl: Python loop + NCCL collectives; peer: native driver, peer-memory collectives")
    args = ap.parse_args()
    import faulthandler

    faulthandler.dump_traceback_later(60, exit=True)  # a wedged job must not hold the GPUs
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    of_b200.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mode = of_b200.MODE_FAST if args.mode == "fast" else of_b200.MODE_EXACT
    prev, curr, _ = synthetic.make_pairs_numpy(1, args.height, args.width, seed=77)
    p, c = prev[0], np.roll(curr[0], 2, axis=0)
    dev = torch.device("cuda", local)
    pd, cd = torch.from_numpy(p).to(dev), torch.from_numpy(c).to(dev)
    backend = ofd.CudaBackend()
    times = []
    plan = None
    if args.driver == "peer":
        plan = ofd.PeerRowbands(args.height, args.width, args.levels, 5, args.iters, mode)
    for _ in range(args.repeat + 1):  # first pass is warm-up
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if plan is not None:
            ud, vd = plan.run(pd, cd)
        else:
            ud, vd = ofd.lk_pyramidal_rowbands(pd, cd, args.levels, 5, args.iters, mode=mode, backend=backend, to_host=False)
        torch.cuda.synchronize()
        dist.barrier()
        times.append(time.perf_counter() - t0)
    if plan is not None:
        plan.trace()  # raises if a peer wait timed out
    times = times[1:]
    if args.timeline and plan is not None:
        from torch.profiler import ProfilerActivity, profile

        dist.barrier()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            plan.run(pd, cd)
            torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            evs = sorted((e for e in prof.events() if e.device_type.name == "CUDA"), key=lambda e: e.time_range.start)
            t0 = evs[0].time_range.start if evs else 0
            with open(args.timeline, "w") as f:
                f.write("start_us,duration_us,name\n")
                for e in evs:
                    f.write(f"{e.time_range.start - t0:.1f},{e.time_range.end - e.time_range.start:.1f},{e.name[:60]}\n")
    u, v = ud.cpu().numpy(), vd.cpu().numpy()
    ok = None
    t_single = None
    if rank == 0:
        # single-GPU reference: device-resident call of the fused driver on the same frames
        ws_bytes = of_b200.lk_pyramidal_workspace_bytes(1, args.height, args.width, args.levels, args.iters)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        u1d, v1d = torch.empty_like(pd), torch.empty_like(pd)
        ts = []
        for _ in range(args.repeat + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            of_b200.lk_pyramidal_dev(pd.data_ptr(), cd.data_ptr(), u1d.data_ptr(), v1d.data_ptr(), 1, args.height,
                                     args.width, args.levels, 5, args.iters, mode, ws.data_ptr(), ws_bytes, None, None,
                                     torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        t_single = min(ts[1:])
        u1, v1 = u1d.cpu().numpy(), v1d.cpu().numpy()
        ok = bool(np.array_equal(u.view(np.uint32), u1.view(np.uint32)) and np.array_equal(v.view(np.uint32), v1.view(np.uint32)))
    faulthandler.cancel_dump_traceback_later()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print(json.dumps({"world": world, "shape": [args.height, args.width], "levels": args.levels, "iters": args.iters,
                          "mode": args.mode, "driver": args.driver, "bit_equal_to_single_gpu": ok, "rowband_device_resident_s": min(times),
                          "single_gpu_device_resident_s": t_single,
                          "mpixel_per_s_rowband": args.height * args.width / min(times) / 1e6}))
        assert ok


if __name__ == "__main__":
    main()
