"""GPU parity for the flat-histogram path (multicanonical / Wang-Landau) against the oracle and
against the exact 8x8 density of states (tests/golden/ising2d_8x8_logdos.csv, which
gen_exact_dos.py checked against the reference's ising2D_8x8.csv)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def m():
    import mcx_b200
    mcx_b200.lib()
    return mcx_b200


def _exact_logdos(L=8):
    rows = [ln.split(",") for ln in open(os.path.join(GOLD, "ising2d_%dx%d_logdos.csv" % (L, L))) if ln[0] in "-0123456789"]
    return {int(r[0]): float(r[2]) for r in rows}


def test_muca_ising_trajectory_bit_exact(m, oracle):
    L, seed, nsweeps = 8, 1000, 50
    bins = range(-2 * L * L, 2 * L * L + 1, 4)
    sys_ = m.Ising([L, L])
    rng = m.PhiloxRNG(seed, 0)
    sys_.init_("random", rng=rng)
    alg = m.Multicanonical(rng, bins)
    lw0 = np.linspace(0.0, 1.5, len(bins)) ** 2
    alg.ensemble.logweight_table.values[:] = lw0
    m.sweep_(sys_, alg, nsweeps)

    s = oracle.System(oracle.ISING, [L, L])
    s.init_random(seed, 0)
    a = oracle.Alg(oracle.METROPOLIS, 0.0)
    f = oracle.Flat(bins[0], 4, len(bins))
    f.logweight[:] = lw0
    assert s.flat_sweep(a, f, 0, 0, 0.0, seed, 0, 0, nsweeps) == 0
    assert np.array_equal(sys_.spins, s.spins)
    assert np.array_equal(alg.ensemble.histogram.values, f.histogram)
    assert alg.ensemble.histogram.values.sum() == nsweeps * L * L == alg.steps
    assert alg.accepted == a.accepted
    assert sys_.energy() == s.energy() == s.energy(full=True)
    # update!(ens; mode=:simple): lw -= log(h) where h > 0
    alg.ensemble.update_()
    f.muca_update()
    assert np.allclose(alg.ensemble.logweight_table.values, f.logweight, rtol=0, atol=1e-12)
    # a second iteration continues the same streams (sweep counter carried over)
    alg.reset_()
    f.histogram[:] = 0
    m.sweep_(sys_, alg, 20)
    a2 = oracle.Alg(oracle.METROPOLIS, 0.0)
    assert s.flat_sweep(a2, f, 0, 0, 0.0, seed, 0, nsweeps, 20) == 0
    assert np.array_equal(sys_.spins, s.spins) and np.array_equal(alg.ensemble.histogram.values, f.histogram)


def test_wang_landau_trajectory_bit_exact(m, oracle):
    L, seed, nsweeps = 8, 77, 40
    bins = range(-2 * L * L, 2 * L * L + 1, 4)
    sys_ = m.Ising([L, L])
    rng = m.PhiloxRNG(seed, 5)
    sys_.init_("random", rng=rng)
    alg = m.WangLandau(rng, bins, logf=1.0)
    m.sweep_(sys_, alg, nsweeps)
    alg.ensemble.update_()             # logf <- logf / 2
    m.sweep_(sys_, alg, nsweeps)

    s = oracle.System(oracle.ISING, [L, L])
    s.init_random(seed, 5)
    a = oracle.Alg(oracle.METROPOLIS, 0.0)
    f = oracle.Flat(bins[0], 4, len(bins), logf=1.0)
    assert s.flat_sweep(a, f, 1, 0, 0.0, seed, 5, 0, nsweeps) == 0
    f.f.logf = 0.5
    assert s.flat_sweep(a, f, 1, 0, 0.0, seed, 5, nsweeps, nsweeps) == 0
    assert alg.ensemble.logf == 0.5
    assert np.array_equal(sys_.spins, s.spins)
    assert np.array_equal(alg.ensemble.logweight_table.values, f.logweight)
    assert alg.accepted == a.accepted and alg.steps == a.steps


def test_muca_blume_capel_pair_spin2(m, oracle):
    """muca_BlumeCapel.jl: Boltzmann(T=0.9) on the pair term, multicanonical in sum s^2, bins 0:1:N"""
    L, seed, nsweeps, T = 8, 42, 30, 0.9
    N = L * L
    sys_ = m.BlumeCapel([L, L])
    rng = m.PhiloxRNG(seed, 0)
    ens = m.PairBoltzmannSpin2Ensemble(m.BoltzmannEnsemble(T=T), m.MulticanonicalEnsemble(range(0, N + 1)))
    alg = m.Metropolis(rng, ens)
    m.sweep_(sys_, alg, nsweeps)

    s = oracle.System(oracle.BLUME_CAPEL, [L, L], J=1.0, D=0.0)
    a = oracle.Alg(oracle.METROPOLIS, 0.0)
    f = oracle.Flat(0, 1, N + 1)
    assert s.flat_sweep(a, f, 0, 1, 1.0 / T, seed, 0, 0, nsweeps) == 0
    assert np.array_equal(sys_.spins, s.spins)
    assert np.array_equal(ens.spin2.histogram.values, f.histogram)
    assert sys_.spin2_sum() == s.spin2_sum() and alg.accepted == a.accepted


def test_muca_many_chains_share_weights_and_histogram(m, oracle):
    """ParallelMulticanonical on one GPU: chains share the weights; the device histogram is the sum
    merge_histograms! would form (parallel_multicanonical.jl:38-48)"""
    L, seed, nsweeps, nch = 8, 9, 25, 40
    bins = range(-2 * L * L, 2 * L * L + 1, 4)
    sys_ = m.Ising([L, L], nchains=nch)
    rng = m.PhiloxRNG(seed, 0)
    sys_.init_("random", rng=rng)
    alg = m.Multicanonical(rng, bins)
    m.sweep_(sys_, alg, nsweeps)
    total = np.zeros(len(bins))
    for c in range(nch):
        s = oracle.System(oracle.ISING, [L, L])
        s.init_random(seed, c)
        f = oracle.Flat(bins[0], 4, len(bins))
        assert s.flat_sweep(oracle.Alg(0, 0.0), f, 0, 0, 0.0, seed, c, 0, nsweeps) == 0
        total += f.histogram
        assert np.array_equal(sys_.spins[c], s.spins)
    assert np.array_equal(alg.ensemble.histogram.values, total)


def test_bounds_error_and_window_policy(m):
    L = 8
    sys_ = m.Ising([L, L])            # all up: E = -128
    rng = m.PhiloxRNG(3, 0)
    alg = m.Multicanonical(rng, range(-128, -100, 4))   # too narrow: the chain walks out
    with pytest.raises(IndexError):
        m.sweep_(sys_, alg, 50)
    # policy 1 (energy window): proposals leaving the window are rejected instead
    sys2 = m.Ising([L, L])
    alg2 = m.Multicanonical(m.PhiloxRNG(3, 0), range(-128, -100, 4))
    m.flat_for(sys2, alg2, policy=1).sweep_(50)
    assert -128 <= sys2.energy() <= -104
    assert alg2.ensemble.histogram.values.sum() == 50 * L * L


def test_muca_iterations_converge_to_exact_dos(m):
    """statistical gate: log g(E) from multicanonical iterations vs the exact 8x8 DOS
    (RMSE as in muca_Ising2D.jl:36-39).  Tolerance: RMSE < 0.35 after 8 iterations of
    2000 sweeps x 64 chains (typical value ~0.1)."""
    L, nch = 8, 64
    exact = _exact_logdos(L)
    bins = range(-2 * L * L, 2 * L * L + 1, 4)
    sys_ = m.Ising([L, L], nchains=nch)
    rng = m.PhiloxRNG(1000, 0)
    sys_.init_("random", rng=rng)
    alg = m.Multicanonical(rng, bins)
    for _ in range(8):
        m.sweep_(sys_, alg, 100)       # thermalise
        alg.reset_()
        m.sweep_(sys_, alg, 2000)
        alg.ensemble.update_()
    lw = alg.ensemble.logweight_table
    est = {e: -(lw[e] - lw[0]) for e in exact}
    ref = {e: exact[e] - exact[0] for e in exact}
    rmse = np.sqrt(np.mean([(est[e] - ref[e]) ** 2 for e in exact]))
    assert rmse < 0.35, rmse


def _wl_device_vs_oracle(m, oracle, model, dims, nch, first_chain, seed, init, lo, step, nbins, policy, logfs, nsweeps,
                         observable=0, beta_pair=0.0):
    """Wang-Landau chains through the C ABI (one table per chain) against the oracle, several logf stages."""
    import ctypes as C
    lib, check = m.lib(), m._lib.check
    N = int(np.prod(dims))
    sys_ = (m.Ising if model == 0 else m.BlumeCapel)(dims, nchains=nch)
    rng = m.PhiloxRNG(seed, first_chain)
    if init == "random":
        sys_.init_("random", rng=rng)
    else:
        check(lib.mcx_lattice_set_first_chain_id(sys_.h_lat, first_chain))
    sys_.set_rng(seed, 0)
    h = C.c_void_p()
    check(lib.mcx_flat_create(sys_.h_lat, m._lib.FLAT_WANG_LANDAU, observable, lo, step, nbins, beta_pair, policy, C.byref(h)))
    refs = []
    for c in range(nch):
        s = oracle.System(model, dims)
        if init == "random":
            s.init_random(seed, first_chain + c)
        else:
            s.spins = np.ones(N, dtype=np.int8)
        refs.append((s, oracle.Flat(lo, step, nbins), oracle.Alg(0, 0.0)))
    sweep0 = 0
    for logf in logfs:
        check(lib.mcx_flat_set_logf(h, logf))
        check(lib.mcx_flat_sweep(h, nsweeps))
        lw = np.empty((nch, nbins), dtype=np.float64)
        check(lib.mcx_flat_get_logweight(h, lw.ctypes.data))
        spins = np.asarray(sys_.spins).reshape(nch, N)
        acc = np.atleast_1d(sys_.accepted())
        for c, (s, f, a) in enumerate(refs):
            f.f.logf = logf
            assert s.flat_sweep(a, f, 1, observable, beta_pair, seed, first_chain + c, sweep0, nsweeps, policy=policy) == 0
            assert np.array_equal(lw[c], f.logweight), (logf, c)
            assert np.array_equal(spins[c], s.spins), (logf, c)
            assert acc[c] == a.accepted
        sweep0 += nsweeps
    check(lib.mcx_flat_destroy(h))
    return [a.accepted / a.steps for _, _, a in refs]


@pytest.mark.parametrize("spec", [None, "0", "8", "32"])
def test_wang_landau_group_decisions_bit_exact(m, oracle, spec, monkeypatch):
    """k_flat_warp decides several consecutive Wang-Landau attempts at once (lane j replays the j rejections before
    it); MCX_WL_SPEC pins the width (0 = the serial loop on lane 0, unset = adaptive).  Every width must reproduce
    the oracle bit for bit: power-of-two and inexact logf, full range and energy window, 2-D and 3-D, low and high
    acceptance, and the (pair, spin^2) observable of muca_BlumeCapel.jl."""
    if spec is None:
        monkeypatch.delenv("MCX_WL_SPEC", raising=False)
    else:
        monkeypatch.setenv("MCX_WL_SPEC", spec)
    _wl_device_vs_oracle(m, oracle, 0, [8, 8], 3, 5, 77, "random", -128, 4, 65, 0, [1.0, 0.5, 0.25], 40)
    _wl_device_vs_oracle(m, oracle, 0, [8, 8], 2, 0, 78, "random", -64, 4, 33, 1, [float(np.log(2.0)), 0.1], 60)
    _wl_device_vs_oracle(m, oracle, 0, [4, 4, 6], 2, 9, 79, "random", -288, 4, 145, 0, [0.3], 30)
    r = _wl_device_vs_oracle(m, oracle, 0, [16, 16], 2, 1, 80, "up", -512, 4, 257, 0, [1.0, 1e-3], 30)
    assert min(r) > 0.02
    N = 64
    _wl_device_vs_oracle(m, oracle, 1, [8, 8], 2, 3, 81, "up", 0, 1, N + 1, 0, [0.7, 0.35], 30,
                         observable=m._lib.OBS_SPIN2_WITH_PAIR_BOLTZMANN, beta_pair=1 / 0.9)


@pytest.mark.parametrize("window", [None, "0"])
def test_logweight_window_recentres_bit_exact(m, oracle, window, monkeypatch):
    """The chains keep a window of the log-weight table in shared memory (k_flat_warp<WIN>) and recentre it -- dirty
    Wang-Landau entries written back first -- when a batch could leave it.  Tables much wider than the window, walkers
    that start at the edge of the spectrum and travel: every table entry, spin and counter must still equal the
    oracle's, with the window (default) and without it (MCX_FLAT_WINDOW=0)."""
    if window is None:
        monkeypatch.delenv("MCX_FLAT_WINDOW", raising=False)
    else:
        monkeypatch.setenv("MCX_FLAT_WINDOW", window)
    # Wang-Landau, 2-D 64 x 64 from the ground state: 4097 bins against a 1025-entry window
    r = _wl_device_vs_oracle(m, oracle, 0, [64, 64], 3, 2, 90, "up", -8192, 4, 4097, 0, [1.0, 0.25], 12)
    assert min(r) > 0.05
    # the same in an energy window of 1500 bins (policy 1), and 3-D (window 1537 entries against 3073 bins)
    _wl_device_vs_oracle(m, oracle, 0, [64, 64], 2, 0, 91, "up", -8192, 4, 1500, 1, [0.5], 10)
    _wl_device_vs_oracle(m, oracle, 0, [16, 16, 16], 2, 4, 92, "up", -12288, 4, 3073 * 2 - 1, 0, [1.0], 6)
    # multicanonical chains on the (pair, sum s^2) observable, 4097 bins against a 513-entry window, non-flat weights
    import ctypes as C
    lib, check = m.lib(), m._lib.check
    L, nch, seed, nsweeps = 64, 3, 93, 6
    N = L * L
    lw0 = 1e-3 * (np.arange(N + 1, dtype=np.float64) - 0.6 * N) ** 2 / N
    sys_ = m.BlumeCapel([L, L], nchains=nch)
    check(lib.mcx_lattice_set_first_chain_id(sys_.h_lat, 1))
    sys_.set_rng(seed, 0)
    h = C.c_void_p()
    check(lib.mcx_flat_create(sys_.h_lat, m._lib.FLAT_MUCA, m._lib.OBS_SPIN2_WITH_PAIR_BOLTZMANN, 0, 1, N + 1, 1 / 0.9, 0, C.byref(h)))
    check(lib.mcx_flat_set_logweight(h, lw0.ctypes.data))
    check(lib.mcx_flat_sweep(h, nsweeps))
    hist = np.empty(N + 1, dtype=np.float64)
    check(lib.mcx_flat_get_histogram(h, hist.ctypes.data))
    spins = np.asarray(sys_.spins).reshape(nch, N)
    ref_hist = np.zeros(N + 1)
    for c in range(nch):
        s = oracle.System(oracle.BLUME_CAPEL, [L, L])
        s.spins = np.ones(N, dtype=np.int8)
        f = oracle.Flat(0, 1, N + 1)
        f.logweight[:] = lw0
        assert s.flat_sweep(oracle.Alg(0, 0.0), f, 0, 1, 1 / 0.9, seed, 1 + c, 0, nsweeps) == 0
        assert np.array_equal(spins[c], s.spins), c
        ref_hist += f.histogram
    assert np.array_equal(hist, ref_hist)
    check(lib.mcx_flat_destroy(h))
