"""mcx_pt_run as ONE persistent launch (k_persist.cu: sweeps, energies, replica exchange and labels decided inside the
kernel by the warp that finishes a round's last work item) must equal the rounds queued from the host
(mcx_sweep + mcx_pt_publish + mcx_pt_exchange, MCX_PT_PERSIST=0) bit for bit -- and those are held to the oracle in
tests/test_gpu_parity.py::test_parallel_tempering_matches_oracle; the oracle comparison is repeated here directly."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def m():
    import mcx_b200
    mcx_b200.lib()
    return mcx_b200


def _run(m, dims, n, seed, calls, rule="metropolis"):
    """calls = [(nrounds, sweeps_per_round), ...] successive run_ calls; returns the whole observable state"""
    betas = m.set_betas(n, 0.3, 0.6, "uniform")
    pt = m.ParallelTempering(betas, seed=seed, backend=m.GPUBackend())
    reps = m.Ising(dims, nchains=n)
    pt.attach(reps)
    reps.init_("random", rng=m.PhiloxRNG(seed, 0))
    paths = []
    for nrounds, every in calls:
        pt.run_(reps, nrounds, every)
        p, r = C.c_int32(), C.c_int32()
        m._lib.check(m.lib().mcx_pt_run_info(pt._pt, C.byref(p), C.byref(r)))
        paths.append((p.value, r.value))
    state = dict(index=list(map(int, pt.index())), steps=list(map(int, pt.steps)), accepted=list(map(int, pt.accepted)),
                 stage=int(pt.stage), round=int(pt.round), energies=[float(e) for e in pt.energies()],
                 pair=[int(v) for v in np.atleast_1d(reps.pair_sum())], mag=[int(v) for v in np.atleast_1d(reps.magnetization())],
                 acc=[int(v) for v in np.atleast_1d(reps.accepted())], sweep=int(reps.sweep_index))
    assert np.array_equal(np.atleast_1d(reps.energy()), np.atleast_1d(reps.energy(full=True)))
    spins = reps.spins.copy()
    code = C.c_int32(-1)
    m._lib.check(m.lib().mcx_ctx_async_error(reps.ctx.h, C.byref(code)))
    assert code.value == 0
    return state, spins, paths


@pytest.mark.parametrize("dims,n", [([64, 64], 10), ([256, 256], 5), ([1024, 64], 6), ([32, 128], 7), ([1024, 1024], 4)])
@pytest.mark.parametrize("calls", [[(12, 1)], [(6, 2), (3, 1)], [(4, 3)], [(3, 7), (2, 1), (1, 12)]])
def test_persistent_rounds_equal_host_queued_rounds(m, dims, n, calls, monkeypatch):
    monkeypatch.setenv("MCX_PT_PERSIST", "0")
    ref, ref_spins, ref_paths = _run(m, dims, n, 77, calls)
    assert all(p == 0 for p, _ in ref_paths)
    monkeypatch.setenv("MCX_PT_PERSIST", "1")
    got, got_spins, paths = _run(m, dims, n, 77, calls)
    assert all(p == 1 for p, _ in paths), paths
    assert got == ref
    assert np.array_equal(got_spins, ref_spins)
    assert sum(ref["steps"]) == sum((n - 1 + (k % 2 == 0)) // 2 for k in range(ref["round"]))
    if dims == [64, 64] and calls == [(12, 1)]:    # close enough temperatures for this lattice to exchange: the compared runs
        assert sum(ref["accepted"]) > 0            # do contain accepted swaps (a ~1 % event per attempt, fixed by the seed)


@pytest.mark.parametrize("rows", ["2", "6", "10", "16"])
def test_persistent_rounds_strip_heights(m, rows, monkeypatch):
    """every strip height, including heights that do not divide Ly (ceil(Ly / rows) strips of even heights that differ
    by at most two rows)"""
    calls = [(5, 1), (2, 4)]
    monkeypatch.setenv("MCX_PT_PERSIST", "0")
    ref, ref_spins, _ = _run(m, [1024, 64], 5, 5, calls)
    monkeypatch.setenv("MCX_PT_PERSIST", "1")
    monkeypatch.setenv("MCX_QUEUE_ROWS", rows)
    got, got_spins, paths = _run(m, [1024, 64], 5, 5, calls)
    assert all(abs(r - int(rows)) <= 2 for _, r in paths), paths
    assert got == ref and np.array_equal(got_spins, ref_spins)


def test_persistent_rounds_match_oracle(m, oracle, monkeypatch):
    """the persistent launch against the CPU restatement directly (replica_exchange.jl:158-178 in oracle.rx_update)"""
    monkeypatch.setenv("MCX_PT_PERSIST", "1")
    L, n, seed = 32, 8, 2025
    for every, rounds in ((1, 24), (3, 9)):
        betas = m.set_betas(n, 0.3, 0.6, "uniform")
        pt = m.ParallelTempering(betas, seed=seed, backend=m.GPUBackend())
        sys_ = m.Ising([L, L], nchains=n)
        pt.attach(sys_)
        sys_.init_("random", rng=m.PhiloxRNG(seed, 0))
        pt.run_(sys_, rounds, every)
        o_sys = []
        for r in range(n):
            s = oracle.System(oracle.ISING, [L, L])
            s.init_random(seed, r)
            o_sys.append(s)
        idx = np.arange(1, n + 1, dtype=np.int64)
        steps = np.zeros(n - 1, dtype=np.int64)
        acc = np.zeros(n - 1, dtype=np.int64)
        beta_of_slot = np.array(betas, dtype=np.float64)
        stage = 0
        for rd in range(rounds):
            for r in range(n):
                a = oracle.Alg(oracle.METROPOLIS, float(beta_of_slot[r]))
                o_sys[r].sweep_checkerboard(a, seed, r, rd * every, every)
            xs = [o_sys[r].energy() for r in range(n)]
            us = [oracle.lib().mcxo_exchange_u(seed, r, rd) for r in range(n)]
            stage = oracle.rx_update(stage, idx, steps, acc, beta_of_slot, xs, us)
        assert list(pt.index()) == list(idx)
        assert list(pt.steps) == list(steps) and list(pt.accepted) == list(acc)
        assert pt.stage == stage and acc.sum() > 0
        got = sys_.spins
        for r in range(n):
            assert np.array_equal(got[r], o_sys[r].spins)
        assert np.allclose(pt.energies(), [s.energy() for s in o_sys])
        p = C.c_int32()
        m._lib.check(m.lib().mcx_pt_run_info(pt._pt, C.byref(p), None))
        assert p.value == 1


@pytest.mark.parametrize("calls", [[(40, 1)], [(20, 2), (17, 1)], [(33, 1), (5, 1), (16, 2)]])
def test_graph_replayed_rounds_equal_host_queued_rounds(m, calls, monkeypatch):
    """short-interval rounds replayed from a CUDA graph (kernels read the half-sweep index and the round from a device
    clock) against the same rounds queued launch by launch"""
    monkeypatch.setenv("MCX_PT_PERSIST", "0")
    monkeypatch.setenv("MCX_PT_GRAPH", "0")
    ref, ref_spins, ref_paths = _run(m, [1024, 1024], 5, 123, calls)
    assert all(p == 0 for p, _ in ref_paths)
    monkeypatch.delenv("MCX_PT_GRAPH")
    got, got_spins, paths = _run(m, [1024, 1024], 5, 123, calls)
    assert any(p == 2 for p, _ in paths), paths
    assert got == ref
    assert np.array_equal(got_spins, ref_spins)


def test_persistent_then_manual_rounds(m, monkeypatch):
    """a persistent run followed by explicit sweep_ / update_ rounds (and the other way round) continues the same ladder"""
    L, n, seed = 64, 6, 3
    betas = m.set_betas(n, 0.3, 0.6, "uniform")
    out = []
    for mode in ("manual", "mixed"):
        monkeypatch.setenv("MCX_PT_PERSIST", "1" if mode == "mixed" else "0")
        pt = m.ParallelTempering(betas, seed=seed, backend=m.GPUBackend())
        reps = m.Ising([L, L], nchains=n)
        pt.attach(reps)
        reps.init_("random", rng=m.PhiloxRNG(seed, 0))
        pt.run_(reps, 5, 2)
        for _ in range(3):
            m.sweep_(reps, pt, 2)
            m.update_(pt)
        pt.run_(reps, 4, 1)
        out.append((list(pt.index()), list(pt.steps), list(pt.accepted), reps.spins.copy()))
    assert out[0][:3] == out[1][:3] and np.array_equal(out[0][3], out[1][3])


_WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
ngpu = torch.cuda.device_count()
torch.cuda.set_device(rank %% ngpu)
dist.init_process_group("gloo", rank=rank, world_size=world)
import ctypes as C
import mcx_b200 as m
ctx = m.Context(rank %% ngpu)
L, n, seed = 64, 12, 2025
betas = m.set_betas(n, 0.3, 0.6, "uniform")

def run(backend, calls):
    pt = m.ParallelTempering(betas, seed=seed, backend=backend)
    first, count = backend.slots(n)
    reps = m.Ising([L, L], nchains=count, ctx=ctx)
    pt.attach(reps)
    reps.init_("random", rng=m.PhiloxRNG(seed, first))
    paths = []
    for nrounds, every in calls:
        pt.run_(reps, nrounds, every)
        p = C.c_int32()
        m._lib.check(m.lib().mcx_pt_run_info(pt._pt, C.byref(p), None))
        paths.append(p.value)
    out = (list(map(int, pt.index())), list(map(int, pt.steps)), list(map(int, pt.accepted)), [float(e) for e in pt.energies()])
    return out, pt, paths, reps

res = {}
for k, calls in enumerate([[(10, 1)], [(4, 3), (3, 1)]]):
    os.environ["MCX_PT_PERSIST"] = "1"
    (idx, steps, acc, en), pt, paths, reps = run(m.GPUBackend(), calls)
    assert pt._peers, "peer-store all-gather not attached"
    assert all(p == 1 for p in paths), paths
    assert pt.peer_status() == 0
    sha = reps.spins.tobytes()
    dist.barrier()
    os.environ["MCX_PT_PERSIST"] = "0"
    (idx0, steps0, acc0, en0), pt0, paths0, reps0 = run(m.GPUBackend(), calls)
    same_as_host_path = idx == idx0 and steps == steps0 and acc == acc0 and en == en0 and sha == reps0.spins.tobytes()
    dist.barrier()
    ok = True
    if rank == 0:
        class One(m.GPUBackend):        # the same ladder on one rank, no peers at all
            rank = property(lambda self: 0); size = property(lambda self: 1)
            def slots(self, n_global): return 0, n_global
            def barrier(self): pass
        (idx1, steps1, acc1, en1), _, _, _ = run(One(), calls)
        ok = bool(idx == idx1 and steps == steps1 and acc == acc1 and en == en1 and sum(acc) > 0)
    flag = torch.tensor([int(ok and same_as_host_path)])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res[k] = bool(flag.item())
    dist.barrier()
if rank == 0:
    print(json.dumps({"ok": all(res.values()), "res": {str(k): v for k, v in res.items()}}))
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 3])
def test_persistent_rounds_across_processes(m, world, tmp_path):
    """ranks in separate processes (sharing this GPU when the box has fewer GPUs than ranks): each rank's persistent
    kernel stores its energies into every rank's buffer (CUDA IPC), waits for the others' arrival counters and decides
    the same exchanges -- equal to the host-queued rounds and to the one-rank run"""
    script = tmp_path / "ptp_worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29670 + world), str(script)]
    env = dict(os.environ)
    env.pop("MCX_PT_PERSIST", None)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=420, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert res["ok"], res
