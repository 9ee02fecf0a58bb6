"""3-D Ising sweep rate: vectorised k_ising3d vs the rows-of-8 kernel (MCX_ISING3D=1 / 0)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mcx_b200 as m
from mcx_b200._lib import check, lib

stream = torch.cuda.Stream()
ctx = m.Context(0, stream=stream.cuda_stream)
for dims, n in (([512, 512, 512], 1), ([256, 256, 256], 1), ([64, 64, 64], 64)):
    row = {"dims": dims, "chains": n}
    for mode in ("1", "0"):
        os.environ["MCX_ISING3D"] = mode
        for track in (False, True):
            s = m.Ising(dims, nchains=n, ctx=ctx); s.set_tracking(track)
            rng = m.PhiloxRNG(3); alg = m.Metropolis(rng, beta=0.2216)
            m.init_(s, "random", rng=rng)
            ns = 10
            m.sweep_(s, alg, ns)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(stream)
            check(lib().mcx_sweep(s.h_lat, ns))
            e1.record(stream); torch.cuda.synchronize()
            row[("ising3d" if mode == "1" else "rows8") + ("_tracked" if track else "")] = round(ns * n * s.N / (e0.elapsed_time(e1) * 1e6), 1)
            del s
    print(json.dumps(row), flush=True)
