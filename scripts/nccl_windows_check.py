"""Windowed Wang-Landau with neighbour-window exchanges over NCCL (point-to-point between ranks) against the one-rank run:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/nccl_windows_check.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import datetime
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank), timeout=datetime.timedelta(seconds=120))
import mcx_b200 as m


def run(backend):
    wl = m.WangLandauWindows([16, 16], nwindows=8, walkers=2, overlap=0.5, seed=11, backend=backend, device=rank)
    wl.prepare_().run_(0.1, 120, flatness=0.3, max_checks=2, exchange_every=20)
    out = np.concatenate([np.nan_to_num(wl.logdos().values, nan=-1.0), wl.exchange_rates()])
    acc = int(wl.exchange_accepted.sum())
    wl.close()
    return out, acc


got, _ = run(m.GPUBackend())
dist.barrier()
ok = True
if rank == 0:
    class One(m.GPUBackend):
        rank = property(lambda self: 0); size = property(lambda self: 1)
    ref, acc = run(One())
    ok = bool(np.array_equal(got, ref))
    print(json.dumps({"nccl_windows_exchange_equal_to_one_rank": ok, "world": world, "accepted_exchanges": acc}))
dist.barrier()
dist.destroy_process_group()
