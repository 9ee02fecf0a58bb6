"""GPU parity of the slab decomposition (k_slab.cu): a lattice split by rows into slabs whose half-sweeps
read the neighbour rows from the neighbour handle's memory must reproduce the unsplit lattice bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BETA_C = 0.440686793509772
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def m():
    import mcx_b200
    mcx_b200.lib()
    return mcx_b200


def _alg(m, rule, seed):
    rng = m.PhiloxRNG(seed, 2)
    return (m.Metropolis, m.Glauber, m.HeatBath)[rule](rng, beta=BETA_C)


@pytest.mark.parametrize("dims,nslabs", [([256, 256], 2), ([256, 256], 4), ([1024, 96], 3), ([64, 512], 8), ([512, 64], 1)])
@pytest.mark.parametrize("rule", [0, 1, 2])
def test_local_slabs_equal_whole_lattice(m, oracle, dims, nslabs, rule):
    nsweeps = 5
    os.environ["MCX_RESIDENT"] = "0"
    try:
        whole = m.Ising(dims)
        a0 = _alg(m, rule, 77)
        whole.init_("random", rng=a0.rng)
        m.sweep_(whole, a0, nsweeps)
    finally:
        os.environ.pop("MCX_RESIDENT", None)
    # the unsplit lattice the slabs are compared with is itself the oracle's trajectory (not only the device's own)
    ref = oracle.System(oracle.ISING, dims)
    ref.init_random(77, 2)
    ralg = oracle.Alg(rule, BETA_C)
    ref.sweep_checkerboard(ralg, 77, 2, 0, nsweeps)
    assert np.array_equal(whole.spins, ref.spins) and whole.pair_sum() == ref.pair_count()
    for tracking in (True, False):
        slabs = m.SlabIsing(dims, nslabs=nslabs)
        slabs.set_tracking(tracking)
        a1 = _alg(m, rule, 77)
        slabs.init_("random", rng=a1.rng)
        slabs.sweep_(a1, nsweeps)
        assert np.array_equal(slabs.spins, whole.spins)
        assert slabs.pair_sum() == whole.pair_sum()
        assert slabs.magnetization() == whole.magnetization()
        assert slabs.energy(full=True) == whole.energy(full=True)
        assert a1.steps == a0.steps
        if rule != 2:
            assert a1.accepted == a0.accepted
        assert all(t == 0 and e == 2 * nsweeps for t, e in slabs.status())


def test_slab_argument_errors(m):
    s = m.Ising([64, 64])
    from mcx_b200._lib import check, lib
    with pytest.raises(ValueError):
        check(lib().mcx_slab_configure(s.h_lat, 128, 3))        # odd offset
    with pytest.raises(ValueError):
        check(lib().mcx_slab_configure(s.h_lat, 64, 32))        # sticks out of the global lattice
    with pytest.raises(Exception):
        check(lib().mcx_slab_half_sweep(s.h_lat))                # not a slab
    with pytest.raises(ValueError):
        m.SlabIsing([64, 60], nslabs=4)                          # 15 rows per slab
    b = m.BlumeCapel([64, 64])
    with pytest.raises(Exception):
        check(lib().mcx_slab_configure(b.h_lat, 128, 0))


_WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
ngpu = torch.cuda.device_count()
torch.cuda.set_device(rank %% ngpu)
dist.init_process_group("gloo", rank=rank, world_size=world)
import mcx_b200 as m
ctx = m.Context(rank %% ngpu)
dims, nsweeps = [256, 128 * world], 3
s = m.SlabIsing(dims, backend=m.GPUBackend(), ctx=ctx)
rng = m.PhiloxRNG(123, 1)
alg = m.Metropolis(rng, beta=0.44)
s.init_("random", rng=rng)
s.sweep_(alg, nsweeps)
s.sweep_(alg, 1)
spins, pair, mag = s.spins, s.pair_sum(), s.magnetization()
st = s.status()
if rank == 0:
    os.environ["MCX_RESIDENT"] = "0"
    w = m.Ising(dims, ctx=ctx)
    a = m.Metropolis(m.PhiloxRNG(123, 1), beta=0.44)
    w.init_("random", rng=a.rng)
    m.sweep_(w, a, nsweeps + 1)
    ok = bool(np.array_equal(spins, w.spins)) and pair == w.pair_sum() and mag == w.magnetization() and alg.accepted == a.accepted
    print(json.dumps({"ok": ok, "status": st, "pair": int(pair), "ref_pair": int(w.pair_sum()), "acc": int(alg.accepted), "ref_acc": int(a.accepted)}))
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world,bands", [(2, None), (3, None), (2, "2"), (2, "4")])
def test_ipc_slabs_across_processes(m, world, bands, tmp_path):
    """one process per slab (sharing this GPU when the box has fewer GPUs than ranks): CUDA IPC mapping of
    the neighbour's planes, device flags for the ordering -- the multi-GPU path end to end"""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    port = 29600 + world + (10 * int(bands) if bands else 0)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    env = dict(os.environ)
    env.pop("MCX_BANDS", None)
    if bands:                      # every slab additionally split into row bands on auxiliary streams
        env["MCX_BANDS"] = bands
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["ok"], res
    assert all(t == 0 for t, _ in res["status"]), res


_PT_WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
ngpu = torch.cuda.device_count()
torch.cuda.set_device(rank %% ngpu)
dist.init_process_group("gloo", rank=rank, world_size=world)
import mcx_b200 as m
ctx = m.Context(rank %% ngpu)
L, n, rounds, seed = 64, 12, 16, 2025
betas = m.set_betas(n, 0.3, 0.6, "uniform")

def run(backend, every):
    pt = m.ParallelTempering(betas, seed=seed, backend=backend)
    first, count = backend.slots(n)
    reps = m.Ising([L, L], nchains=count, ctx=ctx)
    pt.attach(reps)
    reps.init_("random", rng=m.PhiloxRNG(seed, first))
    for _ in range(rounds):
        m.sweep_(reps, pt, every)
        m.update_(pt)
    out = (list(map(int, pt.index())), list(map(int, pt.steps)), list(map(int, pt.accepted)), [float(e) for e in pt.energies()])
    return out, pt

res = {}
for every in (1, 3):
    (idx, steps, acc, en), pt = run(m.GPUBackend(), every)
    assert pt._peers, "peer-store all-gather not attached"
    assert pt.peer_status() == 0
    dist.barrier()
    if rank == 0:
        class One(m.GPUBackend):        # the same ladder on one rank, no collective at all
            rank = property(lambda self: 0); size = property(lambda self: 1)
            def slots(self, n_global): return 0, n_global
            def barrier(self): pass
        (idx1, steps1, acc1, en1), _ = run(One(), every)
        res[every] = bool(idx == idx1 and steps == steps1 and acc == acc1 and en == en1 and sum(acc) > 0)
    dist.barrier()
if rank == 0:
    print(json.dumps({"ok": all(res.values()), "res": {str(k): v for k, v in res.items()}}))
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 3])
def test_pt_peer_store_allgather_across_processes(m, world, tmp_path):
    """parallel tempering over several processes with the energies exchanged by peer stores (CUDA IPC) and
    device-side arrival counters instead of a collective: ladder state and energies equal the one-rank run"""
    script = tmp_path / "pt_worker.py"
    script.write_text(_PT_WORKER % {"root": ROOT})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29650 + world), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    import json
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert res["ok"], res
