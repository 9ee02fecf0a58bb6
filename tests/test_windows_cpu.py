"""Windowed Wang-Landau (BASELINE.json configs[4]; SURVEY.md 8e "C5 WL windows"): host logic on CPU.

The driver (mcx_b200.windows.WangLandauWindows) is run here against tests/_window_engine.OracleWindow,
the oracle-backed stand-in for the device window, so the window geometry, the drive into the windows,
the stage schedule, the join at the overlaps and the two-rank sharding are all checked without a GPU.
tests/test_gpu_windows.py then requires the device to reproduce these runs bit for bit."""
import csv
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def exact_logdos(L=8):
    out = {}
    with open(os.path.join(ROOT, "tests", "golden", "ising2d_%dx%d_logdos.csv" % (L, L))) as fh:
        for row in csv.reader(fh):
            if row and not row[0].startswith("#") and row[0] != "energy":
                out[int(row[0])] = float(row[2])
    return out


def rmse_vs_exact(g, exact):
    es = sorted(exact)
    est = np.array([g[e] for e in es])
    ref = np.array([exact[e] for e in es])
    assert np.isfinite(est).all()
    mid = es.index(0)
    return float(np.sqrt(np.mean(((est - est[mid]) - (ref - ref[mid])) ** 2)))


def test_partition_windows():
    import mcx_b200 as m
    assert m.partition_windows(33, 1) == [(0, 33)]
    for nbins, nw, ov in ((65, 4, 0.5), (65, 8, 0.75), (25165825, 8, 0.5), (1000, 16, 0.25), (9, 4, 0.5)):
        ws = m.partition_windows(nbins, nw, ov)
        assert len(ws) == nw and ws[0][0] == 0 and ws[-1][0] + ws[-1][1] == nbins
        assert len({w for _, w in ws}) == 1                       # equal widths: the all-gather needs them
        for (a, w), (b, _) in zip(ws[:-1], ws[1:]):
            assert a < b < a + w                                  # ordered, overlapping
    with pytest.raises(ValueError):
        m.partition_windows(65, 4, 1.0)
    with pytest.raises(ValueError):
        m.partition_windows(4, 8, 0.5)
    with pytest.raises(ValueError):
        m.partition_windows(65, 0)


def test_join_recovers_a_known_table_exactly():
    """Cut the exact 8x8 log g (golden, = SpinSystems/data/exact_solutions/ising2D_8x8.csv) into windows with
    arbitrary per-window constants: the join returns it up to one global constant, NaN where unreachable."""
    import mcx_b200 as m
    exact = exact_logdos(8)
    bins = range(-128, 129, 4)
    full = np.array([exact.get(e, 0.0) for e in bins])            # E = +-124 do not exist: "never visited"
    for nw, ov in ((2, 0.3), (4, 0.5), (8, 0.75)):
        ws = m.partition_windows(len(bins), nw, ov)
        consts = np.linspace(-300.0, 500.0, nw)
        pieces = [np.where(full[f:f + w] != 0, full[f:f + w] + c, 0.0) for (f, w), c in zip(ws, consts)]
        g = m.join_logdos(pieces, [f for f, _ in ws], len(bins))
        vis = full != 0
        assert np.isnan(g[~vis]).all() and np.isfinite(g[vis]).all()
        d = g[vis] - full[vis]
        assert np.max(np.abs(d - d[0])) < 1e-9
    with pytest.raises(ValueError):                                # no common visited bin
        m.join_logdos([np.array([1.0, 2.0, 0.0]), np.array([0.0, 3.0, 4.0])], [0, 2], 5)


def test_window_policy_of_the_oracle(oracle):
    """policy 1 = a proposal leaving the window is a rejected attempt that visits the current bin; policy 0 =
    the reference's BoundsError (test/test_multicanonical.jl:39-43)."""
    L, N = 8, 64
    s = oracle.System(oracle.ISING, [L, L])
    s.spins = np.ones(N, dtype=np.int8)
    f = oracle.Flat(-128, 4, 8, logf=0.5)                         # E in [-128, -100]
    a = oracle.Alg(0, 0.0)
    assert s.flat_sweep(a, f, 1, 0, 0.0, 3, 0, 0, 50, policy=1) == 0
    assert -128 <= s.energy(full=True) <= -100
    assert a.steps == 50 * N and -f.logweight.sum() / 0.5 == 50 * N       # every attempt lowers one bin by logf
    assert f.logweight[1] == 0.0                                  # E = -124 does not exist
    s2 = oracle.System(oracle.ISING, [L, L])
    s2.spins = np.ones(N, dtype=np.int8)
    assert s2.flat_sweep(oracle.Alg(0, 0.0), oracle.Flat(-128, 4, 8), 1, 0, 0.0, 3, 0, 0, 50, policy=0) == -1
    # a walker that STARTS outside its window is an error under both policies
    s3 = oracle.System(oracle.ISING, [L, L])
    s3.spins = np.ones(N, dtype=np.int8)
    assert s3.flat_sweep(oracle.Alg(0, 0.0), oracle.Flat(-64, 4, 8), 1, 0, 0.0, 3, 0, 0, 1, policy=1) == -1


def test_windowed_wang_landau_matches_exact_dos(oracle):
    """4 windows x 2 walkers on 8x8, user-driven schedule like the reference's (update! halves logf,
    ensembles/wang_landau.jl:23).  Tolerance: RMSE of log g(E) - log g(0) against the exact DOS < 0.3
    (typical 0.05), every existing energy visited, E = +-124 NaN."""
    import mcx_b200 as m
    from _window_engine import OracleWindow
    wl = m.WangLandauWindows([8, 8], nwindows=4, walkers=2, overlap=0.5, seed=7, window_factory=OracleWindow)
    assert [wl.window_energies(w) for w in range(4)] == [(-128, -28), (-76, 24), (-24, 76), (28, 128)]
    with pytest.raises(AssertionError):
        wl.sweep_(1)
    wl.prepare_()
    for w, E in enumerate(wl.energies()):
        lo, hi = wl.window_energies(w)
        assert ((lo <= E) & (E <= hi)).all()
    wl.sweep_(200)
    assert min(wl.flatness()) > 0.2
    assert all(v.sum() == 200 * 64 * 2 for v in wl.visits())       # one visit per attempt
    wl.update_()
    assert wl.logf == 0.5 and max(abs(v).max() for v in wl.visits()) == 0
    wl.run_(2e-5, 2000)
    for w, E in enumerate(wl.energies()):
        lo, hi = wl.window_energies(w)
        assert ((lo <= E) & (E <= hi)).all()
    g = wl.logdos(anchor=(-128, np.log(2.0)))
    assert np.isnan(g[-124]) and np.isnan(g[124]) and g[-128] == pytest.approx(np.log(2.0))
    assert rmse_vs_exact(g, exact_logdos(8)) < 0.3
    # total number of states: sum g(E) = 2^64 once one constant is fixed
    vals = g.values[np.isfinite(g.values)]
    assert abs(np.log(np.exp(vals - vals.max()).sum()) + vals.max() - 64 * np.log(2.0)) < 0.5


def test_three_dimensional_windows_and_argument_errors(oracle):
    import mcx_b200 as m
    from _window_engine import OracleWindow
    wl = m.WangLandauWindows([4, 4, 4], nwindows=4, walkers=1, seed=11, window_factory=OracleWindow)
    assert wl.bins == range(-192, 193, 4)
    wl.prepare_().run_(0.01, 300)
    g = wl.logdos()
    assert g[-192] == 0.0 and np.isfinite(g[0]) and g[0] > 35.0      # log g(0) ~ 64 log 2 - few
    assert np.nanargmax(g.values) in (47, 48, 49)                    # symmetric DOS peaks at E = 0
    for bad in ([5, 4], [4], [4, 4, 4, 4], [2, 4]):
        with pytest.raises(ValueError):
            m.WangLandauWindows(bad, nwindows=2, window_factory=OracleWindow)
    with pytest.raises(ValueError):
        m.WangLandauWindows([8, 8], nwindows=3, walkers=0, window_factory=OracleWindow)


def test_neighbour_window_exchange(oracle):
    """exchange_: the reference's exchange rule (replica_exchange.jl:110-115) between walkers of neighbouring windows.
    A stage permutes configurations (nothing is created or lost), keeps every walker inside its window, only pairs
    whose energies both lie in the overlap are attempted, and the density of states still converges (RMSE < 0.3)."""
    import hashlib
    import mcx_b200 as m
    from _window_engine import OracleWindow
    wl = m.WangLandauWindows([8, 8], nwindows=4, walkers=2, overlap=0.5, seed=7, window_factory=OracleWindow)
    wl.prepare_()
    swaps = 0
    for _ in range(40):
        wl.sweep_(20)
        before = [s.copy() for s in wl.spins()]
        E_before = [e.copy() for e in wl.energies()]
        acc0 = wl.exchange_accepted.sum()
        wl.exchange_()
        after, E_after = wl.spins(), wl.energies()
        digest = lambda groups: sorted(hashlib.sha256(r.tobytes()).hexdigest() for g_ in groups for r in g_)
        assert digest(before) == digest(after)
        assert sorted(np.concatenate(E_before)) == sorted(np.concatenate(E_after))
        changed = sum(int((a != b).any(axis=1).sum()) for a, b in zip(before, after))
        assert changed <= 2 * (wl.exchange_accepted.sum() - acc0)       # identical configurations may swap unseen
        swaps += wl.exchange_accepted.sum() - acc0
        for w, E in enumerate(E_after):
            lo, hi = wl.window_energies(w)
            assert ((lo <= E) & (E <= hi)).all()
    assert swaps > 0 and wl.exchange_round == 40 and wl.exchange_stage == 0
    assert (wl.exchange_accepted <= wl.exchange_steps).all() and wl.exchange_steps.sum() <= 40 * 2 * 2
    rates = wl.exchange_rates()
    assert rates.shape == (3,) and (rates >= 0).all() and (rates <= 1).all()
    wl.run_(2e-5, 2000, exchange_every=50)
    assert rmse_vs_exact(wl.logdos(), exact_logdos(8)) < 0.3


def test_checkpoint_restart_continues_the_trajectory(oracle, tmp_path):
    """state() / restore(): like the reference's checkpoint example (checkpoint_Ising2D.jl:40-106), a restarted run
    continues the same trajectory -- tables, configurations, exchange counters and the joined log g are identical to
    the uninterrupted run."""
    import pickle
    import mcx_b200 as m
    from _window_engine import OracleWindow
    kw = dict(nwindows=4, walkers=2, overlap=0.5, seed=3, window_factory=OracleWindow)
    full = m.WangLandauWindows([8, 8], **kw)
    full.prepare_()
    part = m.WangLandauWindows([8, 8], **kw)
    with pytest.raises(AssertionError):
        part.state()
    part.prepare_()
    for wl in (full, part):
        wl.run_(0.2, 200, exchange_every=20)
    path = tmp_path / "wl.ckpt"
    with open(path, "wb") as fh:
        pickle.dump(part.state(), fh)
    part.close()
    with open(path, "rb") as fh:
        resumed = m.WangLandauWindows.restore(pickle.load(fh), window_factory=OracleWindow)
    assert resumed.logf == full.logf and resumed.steps == full.steps
    for wl in (full, resumed):
        wl.sweep_(30)
        wl.exchange_()
        wl.run_(0.02, 200, exchange_every=20)
    for a, b in zip(full._lw, resumed._lw):
        assert np.array_equal(a, b)
    for a, b in zip(full.spins(), resumed.spins()):
        assert np.array_equal(a, b)
    assert np.array_equal(full.exchange_steps, resumed.exchange_steps)
    assert np.array_equal(full.exchange_accepted, resumed.exchange_accepted)
    assert np.array_equal(full.logdos().values, resumed.logdos().values, equal_nan=True)
    assert full.flatness() == resumed.flatness()


# ------------------------------------------------------------------ two ranks (gloo)
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    import mcx_b200 as m
    from _window_engine import OracleWindow
    wl = m.WangLandauWindows([8, 8], nwindows=4, walkers=2, seed=7, backend=m.GPUBackend(), window_factory=OracleWindow)
    assert (wl.first, wl.count) == (2 * rank, 2)
    wl.prepare_().run_(0.05, 300, flatness=0.5, max_checks=3, exchange_every=25)
    np.save(os.path.join(out, "rank%d.npy" % rank), np.concatenate([wl.logdos().values, wl.exchange_rates()]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_join_to_the_single_rank_result(tmp_path, oracle):
    """Windows dealt to two ranks (no collective while sampling, point-to-point neighbour-window exchanges, one
    all-gather of the pieces): both ranks hold the same joined table and exchange statistics, bit-identical to the
    one-rank run (streams keyed by global walker, exchange decisions by the counter-based EXCHANGE stream)."""
    import torch.multiprocessing as mp
    import mcx_b200 as m
    from _window_engine import OracleWindow
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (np.load(os.path.join(str(tmp_path), "rank%d.npy" % r)) for r in range(world))
    assert np.array_equal(r0, r1, equal_nan=True)
    wl = m.WangLandauWindows([8, 8], nwindows=4, walkers=2, seed=7, window_factory=OracleWindow)
    wl.prepare_().run_(0.05, 300, flatness=0.5, max_checks=3, exchange_every=25)
    assert np.array_equal(np.concatenate([wl.logdos().values, wl.exchange_rates()]), r0, equal_nan=True)
    assert wl.exchange_accepted.sum() > 0                             # pairs (0,1), (2,3) inside a rank, (1,2) across ranks
    with pytest.raises(ValueError):                                # 4 windows do not divide over 3 ranks
        m.partition_slots(4, 3, 0)
