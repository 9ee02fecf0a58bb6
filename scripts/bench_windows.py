#!/usr/bin/env python
"""Windowed Wang-Landau (BASELINE.json configs[4]): 3-D Ising L=256, energy windows dealt to the ranks,
`walkers` chains per window.  Prints one JSON object (rank 0).  Auxiliary timing, not the headline bench.

    python scripts/bench_windows.py [--L 256] [--windows 8] [--walkers 4] [--sweeps 1]
    python -m torch.distributed.run --nproc-per-node 8 ... scripts/bench_windows.py --windows 8

One process per GPU; no collective while sampling, one all-gather of the window pieces at the end."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=256)
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--windows", type=int, default=8)
    ap.add_argument("--walkers", type=int, default=4)
    ap.add_argument("--overlap", type=float, default=0.5)
    ap.add_argument("--sweeps", type=int, default=1)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    import mcx_b200 as m
    backend = m.GPUBackend()
    wl = m.WangLandauWindows([a.L] * a.dim, nwindows=a.windows, walkers=a.walkers, overlap=a.overlap, seed=42,
                             backend=backend)
    t0 = time.perf_counter()
    wl.prepare_()
    t_prepare = time.perf_counter() - t0
    backend.barrier()
    t0 = time.perf_counter()
    wl.sweep_(a.sweeps)                      # reads the tables back: includes the wait for every local window
    backend.barrier()
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    try:
        visited = int(np.isfinite(wl.logdos().values).sum())
    except ValueError:                       # a few sweeps of a big lattice: neighbouring windows have not met yet
        visited = int((wl.pieces() != 0).sum())
    t_join = time.perf_counter() - t0
    attempts = a.sweeps * wl.N * a.walkers * a.windows
    inside = all(((wl.window_energies(wl.first + j)[0] <= e) & (e <= wl.window_energies(wl.first + j)[1])).all()
                 for j, e in enumerate(wl.energies()))
    if backend.is_root:
        print(json.dumps({"config": "C5 %d-D Ising L=%d Wang-Landau, %d windows x %d walkers over %d GPU(s), overlap %.2f"
                          % (a.dim, a.L, a.windows, a.walkers, world, a.overlap),
                          "bins_per_window": wl.width, "attempts_per_ns": attempts / dt / 1e9, "sweep_seconds": dt,
                          "prepare_seconds": t_prepare, "join_seconds": t_join, "walkers_inside_windows": bool(inside),
                          "bins_visited": visited, "bins": len(wl.bins), "scaling": "weak"}))
    wl.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
