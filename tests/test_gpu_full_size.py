"""Parity at BASELINE.json's full sizes (configs[1..4]) through size-independent properties and,
where the oracle finishes in seconds, bit-exact comparison of selected chains."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
BETA_C = 0.440686793509772


@pytest.fixture(scope="module")
def m():
    import mcx_b200
    mcx_b200.lib()
    return mcx_b200


def test_c2_ising_16384_single_chain(m):
    L = 16384
    for track in (True, False):
        sys_ = m.Ising([L, L])
        sys_.set_tracking(track)
        rng = m.PhiloxRNG(42, 0)
        alg = m.Metropolis(rng, beta=BETA_C)
        sys_.init_("random", rng=rng)
        m.sweep_(sys_, alg, 3)
        e, mag, acc = sys_.energy(), sys_.magnetization(), alg.accepted
        assert e == sys_.energy(full=True) and mag == sys_.magnetization(full=True)
        assert alg.steps == 3 * L * L and 0 < acc < alg.steps
        if track:
            ref = (e, mag, acc)
            g = sys_.spins.reshape(L, L)
            assert int(g.sum(dtype=np.int64)) == mag
            pair = int((g * np.roll(g, 1, 0)).sum(dtype=np.int64) + (g * np.roll(g, 1, 1)).sum(dtype=np.int64))
            assert e == -pair
        else:
            assert (e, mag, acc) == ref          # tracked and untracked sweeps: same trajectory, same sums
        del sys_


def test_c3_parallel_tempering_256_replicas_1024(m, oracle):
    L, n = 1024, 256
    betas = m.set_betas(n, 1 / 3.0, 1 / 1.5, "uniform")
    pt = m.ParallelTempering(betas, seed=42, backend=m.GPUBackend())
    reps = m.Ising([L, L], nchains=n)
    pt.attach(reps)
    reps.init_("random", rng=m.PhiloxRNG(42, 0))
    m.sweep_(reps, pt, 1)
    # before any exchange slot r holds beta_r: two replicas against the oracle, bit for bit
    first = reps.spins
    for r in (0, 255):
        s = oracle.System(oracle.ISING, [L, L])
        s.init_random(42, r)
        s.sweep_checkerboard(oracle.Alg(oracle.METROPOLIS, betas[r]), 42, r, 0, 1)
        assert np.array_equal(first[r], s.spins)
    del first
    for _ in range(6):
        m.update_(pt)
        m.sweep_(reps, pt, 1)
    idx = pt.index()
    assert sorted(idx) == list(range(1, n + 1))
    assert list(pt.steps[0::2]) == [3] * 128 and list(pt.steps[1::2]) == [3] * 127     # stages alternate (replica_exchange.jl:160,176)
    assert 0 < pt.accepted.sum() <= pt.steps.sum()
    assert np.array_equal(np.asarray(reps.energy()), np.asarray(reps.energy(full=True)))
    assert np.array_equal(reps.get_labels(), np.asarray(idx) - 1)
    assert [a.ensemble.beta for a in pt.replica.algs] == [betas[i - 1] for i in idx]


def test_c4_blume_capel_512_muca_1024_chains(m, oracle):
    L, nch, T = 512, 1024, 0.9
    N = L * L
    sys_ = m.BlumeCapel([L, L], nchains=nch)
    rng = m.PhiloxRNG(42, 0)
    ens = m.PairBoltzmannSpin2Ensemble(m.BoltzmannEnsemble(T=T), m.MulticanonicalEnsemble(range(0, N + 1)))
    alg = m.Metropolis(rng, ens)
    m.sweep_(sys_, alg, 1)
    h = ens.spin2.histogram.values
    assert h.sum() == nch * N == alg.steps and 0 < alg.accepted < alg.steps
    assert np.array_equal(np.asarray(sys_.spin2_sum()), np.asarray(sys_.spin2_sum()))
    e_cached = np.asarray(sys_.energy())
    assert np.array_equal(e_cached, np.asarray(sys_.energy(full=True)))
    spins = sys_.spins
    for c in (0, nch - 1):
        s = oracle.System(oracle.BLUME_CAPEL, [L, L], J=1.0, D=0.0)
        f = oracle.Flat(0, 1, N + 1)
        assert s.flat_sweep(oracle.Alg(0, 0.0), f, 0, 1, 1.0 / T, 42, c, 0, 1) == 0
        assert np.array_equal(spins[c], s.spins)
    ens.spin2.update_()
    assert np.isfinite(ens.spin2.logweight_table.values).all()


def test_c5_ising3d_256_wang_landau_walkers(m, oracle):
    """3-D Ising L=256, two Wang-Landau walkers with private tables on an energy window of
    2^19+1 bins around E=0 (C ABI directly: the Python layer binds one chain per ensemble)."""
    L, nch, seed = 256, 2, 42
    N = L ** 3
    lo, step, nbins = -(1 << 20), 4, (1 << 19) + 1
    sys_ = m.Ising([L, L, L], nchains=nch)
    rng = m.PhiloxRNG(seed, 0)
    sys_.init_("random", rng=rng)
    sys_.set_rng(seed)
    lib, check = m.lib(), m._lib.check
    h = C.c_void_p()
    check(lib.mcx_flat_create(sys_.h_lat, m._lib.FLAT_WANG_LANDAU, m._lib.OBS_ENERGY, lo, step, nbins, 0.0, 0, C.byref(h)))
    check(lib.mcx_flat_set_logf(h, 1.0))
    check(lib.mcx_flat_sweep(h, 1))
    lw = np.empty((nch, nbins), dtype=np.float64)
    check(lib.mcx_flat_get_logweight(h, lw.ctypes.data))
    spins = sys_.spins
    e = np.asarray(sys_.energy())
    assert np.array_equal(e, np.asarray(sys_.energy(full=True)))
    assert (-lw.sum(axis=1) == N).all()                     # every attempt lowers one bin by logf = 1
    for c in range(nch):
        s = oracle.System(oracle.ISING, [L, L, L])
        s.init_random(seed, c)
        f = oracle.Flat(lo, step, nbins, logf=1.0)
        assert s.flat_sweep(oracle.Alg(0, 0.0), f, 1, 0, 0.0, seed, c, 0, 1) == 0
        assert np.array_equal(spins[c], s.spins) and np.array_equal(lw[c], f.logweight)
        assert e[c] == s.energy()
    check(lib.mcx_flat_destroy(h))
