"""Parity on MORE THAN ONE physical GPU (skipped on a one-GPU box): the cross-GPU paths -- slabs of one lattice over CUDA
IPC / NVLink, parallel tempering with peer-store all-gather, Wang-Landau windows dealt to ranks, and the multicanonical
histogram all-reduce over NCCL -- each rank on its OWN device, results equal to the one-process run bit for bit."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_ngpu() < 2, reason="needs at least two GPUs")


def _run(script_text, tmp_path, world, port, name):
    script = tmp_path / name
    script.write_text(script_text % {"root": ROOT})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    return json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])


_DEVICES = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
t = torch.tensor([torch.cuda.current_device()], device="cuda")
out = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(out, t)
if rank == 0:
    print(json.dumps({"devices": [int(o.item()) for o in out], "uuid_count": len({str(torch.cuda.get_device_properties(i).uuid) for i in range(world)})}))
dist.destroy_process_group()
'''


@needs2
def test_ranks_sit_on_distinct_devices(tmp_path):
    world = min(_ngpu(), 4)
    res = _run(_DEVICES, tmp_path, world, 29701, "devices.py")
    assert res["devices"] == list(range(world)) and res["uuid_count"] == world


@needs2
def test_slabs_on_distinct_gpus(tmp_path):
    from test_gpu_slab import _WORKER
    world = min(_ngpu(), 4)
    res = _run(_WORKER, tmp_path, world, 29702, "slab.py")
    assert res["ok"] and all(t == 0 for t, _ in res["status"]), res


@needs2
def test_parallel_tempering_peers_on_distinct_gpus(tmp_path):
    from test_gpu_slab import _PT_WORKER
    res = _run(_PT_WORKER, tmp_path, min(_ngpu(), 4), 29703, "pt.py")
    assert res["ok"], res


@needs2
def test_windows_on_distinct_gpus(tmp_path):
    from test_gpu_windows import _WINDOWS_WORKER
    res = _run(_WINDOWS_WORKER, tmp_path, 2, 29704, "windows.py")
    assert res["ok"], res


_MUCA = r'''
import os, sys, json, ctypes as C
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import mcx_b200 as m
from mcx_b200._lib import check, lib
from mcx_b200.parallel import _as_torch
L, chains, iters = 32, 8 * world, 3
N = L * L

def run(first, count, nranks):
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = m.Context(rank, stream=stream.cuda_stream)
        s = m.BlumeCapel([L, L], nchains=count, ctx=ctx)
        check(lib().mcx_lattice_set_first_chain_id(s.h_lat, first))
        s.set_rng(7, 0)
        h = C.c_void_p()
        check(lib().mcx_flat_create(s.h_lat, m._lib.FLAT_MUCA, m._lib.OBS_SPIN2_WITH_PAIR_BOLTZMANN, 0, 1, N + 1, 1 / 0.9, 0, C.byref(h)))
        p, nb = C.c_void_p(), C.c_int64()
        check(lib().mcx_flat_device_histogram(h, C.byref(p), C.byref(nb)))
        hist = _as_torch(p.value, nb.value, torch.int64, rank)
        for _ in range(iters):
            check(lib().mcx_flat_sweep(h, 2))
            check(lib().mcx_flat_reset_histogram(h))
            check(lib().mcx_flat_sweep(h, 6))
            if nranks > 1:
                dist.all_reduce(hist)          # merge_histograms! (parallel_multicanonical.jl:38-52) over NCCL, on the library's stream
            check(lib().mcx_flat_update(h))
        lw, hv = np.empty(N + 1), np.empty(N + 1)
        check(lib().mcx_flat_get_logweight(h, lw.ctypes.data))
        check(lib().mcx_flat_get_histogram(h, hv.ctypes.data))
        spins = np.asarray(s.spins).reshape(count, N).copy()
        check(lib().mcx_flat_destroy(h))
    return lw, hv, spins

count = chains // world
lw, hv, spins = run(rank * count, count, world)
gathered = [None] * world
dist.all_gather_object(gathered, spins)
ok = True
if rank == 0:
    lw1, hv1, spins1 = run(0, chains, 1)
    ok = bool(np.array_equal(lw, lw1) and np.array_equal(hv, hv1) and np.array_equal(np.concatenate(gathered), spins1) and hv.sum() == 6 * chains * N)
dist.barrier()
if rank == 0:
    print(json.dumps({"ok": ok, "visits": float(hv.sum())}))
dist.destroy_process_group()
'''


@needs2
def test_multicanonical_histogram_allreduce_on_distinct_gpus(tmp_path):
    """config 4's exchange step on hardware: chains sharded over ranks, one NCCL all-reduce of the device histogram per
    iteration, every rank applying the same update! -- weights, histogram and every chain equal the one-rank run"""
    res = _run(_MUCA, tmp_path, min(_ngpu(), 4), 29705, "muca.py")
    assert res["ok"], res
