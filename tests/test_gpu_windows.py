"""GPU parity of windowed Wang-Landau (BASELINE.json configs[4]; SURVEY.md 8e "C5 WL windows").

The same driver is run twice -- on the device (mcx_b200.windows.DeviceWindow: checkerboard Glauber kernels for
the drive, k_flat_warp with out_of_range_policy = 1 for the walkers) and on the oracle-backed stand-in
(tests/_window_engine.OracleWindow) -- and every configuration, table and the joined log g must be
identical bit for bit.  The statistical gate is the exact 8x8 density of states."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def m():
    import mcx_b200
    mcx_b200.lib()
    return mcx_b200


@pytest.mark.parametrize("dims,nwindows,walkers,overlap", [([8, 8], 4, 2, 0.5), ([4, 4, 8], 2, 3, 0.6),
                                                           ([32, 16], 8, 1, 0.75)])
def test_windows_device_equals_oracle(m, oracle, dims, nwindows, walkers, overlap):
    from _window_engine import OracleWindow
    dev = m.WangLandauWindows(dims, nwindows=nwindows, walkers=walkers, overlap=overlap, seed=2024)
    ref = m.WangLandauWindows(dims, nwindows=nwindows, walkers=walkers, overlap=overlap, seed=2024,
                              window_factory=OracleWindow)
    dev.prepare_()
    ref.prepare_()
    for a, b in zip(dev.spins(), ref.spins()):                      # the drive: canonical Glauber sweeps, beta of both signs
        assert np.array_equal(a, b)
    for stage in range(3):
        dev.sweep_(40)
        ref.sweep_(40)
        for j in range(dev.count):
            assert np.array_equal(dev._lw[j], ref._lw[j]), (stage, j)
        assert dev.flatness() == ref.flatness()
        dev.exchange_()                                              # neighbour-window exchange: same decisions, same swaps
        ref.exchange_()
        assert np.array_equal(dev.exchange_accepted, ref.exchange_accepted)
        assert np.array_equal(dev.exchange_steps, ref.exchange_steps)
        for a, b in zip(dev.spins(), ref.spins()):
            assert np.array_equal(a, b)
        dev.update_()
        ref.update_()
    for a, b in zip(dev.spins(), ref.spins()):
        assert np.array_equal(a, b)
    for w, (ea, eb) in enumerate(zip(dev.energies(), ref.energies())):
        lo, hi = dev.window_energies(w)
        assert np.array_equal(ea, eb) and ((lo <= ea) & (ea <= hi)).all()
    assert np.array_equal(dev.logdos().values, ref.logdos().values, equal_nan=True)
    assert dev.steps == ref.steps == 3 * 40 * int(np.prod(dims)) * walkers * nwindows
    dev.close()


def test_windowed_wang_landau_vs_exact_8x8(m):
    """Tolerance: RMSE of log g(E) - log g(0) against the exact 8x8 DOS < 0.3 (typical 0.05; the single-window
    run of test_gpu_statistics.py is held to 0.5)."""
    from test_windows_cpu import exact_logdos, rmse_vs_exact
    wl = m.WangLandauWindows([8, 8], nwindows=4, walkers=2, overlap=0.5, seed=7)
    wl.prepare_().run_(2e-5, 2000)
    g = wl.logdos(anchor=(-128, np.log(2.0)))
    assert np.isnan(g[-124]) and np.isnan(g[124])
    assert rmse_vs_exact(g, exact_logdos(8)) < 0.3
    wl.close()


def test_windows_run_concurrently_on_their_own_streams(m):
    """the windows of one rank live on separate contexts: sweeps are queued for all of them before the first
    table is read back"""
    wl = m.WangLandauWindows([16, 16], nwindows=4, walkers=2, seed=5)
    assert len({w.ctx.h.value for w in wl.local}) == 4
    wl.prepare_()
    before = [w.ctx.launch_count() for w in wl.local]
    wl.sweep_(10)
    assert all(w.ctx.launch_count() > b for w, b in zip(wl.local, before))
    assert all(v.sum() == 10 * 256 * 2 for v in wl.visits())
    wl.close()


def test_windows_checkpoint_restart_on_device(m):
    """state() / restore() on the device: the restarted run continues the uninterrupted run's trajectory"""
    import pickle
    kw = dict(nwindows=4, walkers=2, overlap=0.5, seed=3)
    full = m.WangLandauWindows([16, 16], **kw).prepare_()
    part = m.WangLandauWindows([16, 16], **kw).prepare_()
    for wl in (full, part):
        wl.run_(0.2, 60, exchange_every=20)
    blob = pickle.dumps(part.state())
    part.close()
    resumed = m.WangLandauWindows.restore(pickle.loads(blob))
    for wl in (full, resumed):
        wl.sweep_(25)
        wl.exchange_()
        wl.run_(0.04, 60, exchange_every=20)
    for a, b in zip(full._lw, resumed._lw):
        assert np.array_equal(a, b)
    for a, b in zip(full.spins(), resumed.spins()):
        assert np.array_equal(a, b)
    assert np.array_equal(full.exchange_accepted, resumed.exchange_accepted)
    assert np.array_equal(full.logdos().values, resumed.logdos().values, equal_nan=True)
    full.close()
    resumed.close()


_WINDOWS_WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
ngpu = torch.cuda.device_count()
torch.cuda.set_device(rank %% ngpu)
dist.init_process_group("gloo", rank=rank, world_size=world)
import mcx_b200 as m

def run(backend):
    wl = m.WangLandauWindows([16, 16], nwindows=4, walkers=2, overlap=0.5, seed=11, backend=backend, device=rank %% ngpu)
    wl.prepare_().run_(0.1, 120, flatness=0.3, max_checks=2, exchange_every=20)
    out = np.concatenate([wl.logdos().values, wl.exchange_rates()])
    accepted = int(wl.exchange_accepted.sum())
    wl.close()
    return out, accepted

got, _ = run(m.GPUBackend())                 # windows 0-1 on rank 0, 2-3 on rank 1; pair (1, 2) crosses the ranks
dist.barrier()
ok = True
if rank == 0:
    class One(m.GPUBackend):                 # all four windows on one rank, nothing exchanged between processes
        rank = property(lambda self: 0); size = property(lambda self: 1)
    ref, accepted = run(One())
    ok = bool(np.array_equal(got, ref, equal_nan=True) and accepted > 0)
dist.barrier()
if rank == 0:
    print(json.dumps({"ok": ok}))
dist.destroy_process_group()
'''


def test_windows_over_two_processes_equal_one_process(m, tmp_path):
    """windows dealt to two processes (gloo for the plumbing, the device for everything else): cross-rank neighbour
    exchange and the final all-gather give the one-process result bit for bit"""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "wl_worker.py"
    script.write_text(_WINDOWS_WORKER % {"root": root})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29671", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    res = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert res["ok"], res
