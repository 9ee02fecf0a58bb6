"""Statistical parity gates (BASELINE.json north_star): energy, |m| and the Binder cumulant at
beta_c from the GPU checkerboard/Philox sweeps must agree with the reference's own random-site
Xoshiro loop (oracle mode 2) within error bars; canonical P(E) and the Wang-Landau log g(E) must
agree with the exact 8x8 density of states (golden file pinned to the reference's csv)."""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
BETA_C = 0.440686793509772
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def m():
    import mcx_b200
    mcx_b200.lib()
    return mcx_b200


def _exact_logdos(L=8):
    rows = [ln.split(",") for ln in open(os.path.join(GOLD, "ising2d_%dx%d_logdos.csv" % (L, L))) if ln[0] in "-0123456789"]
    return {int(r[0]): float(r[2]) for r in rows}


def _gpu_chain_averages(m, L, beta, rule, nchains, therm, sweeps, interval, seed):
    sys_ = m.Ising([L, L], nchains=nchains)
    rng = m.PhiloxRNG(seed, 0)
    alg = (m.Metropolis, m.Glauber, m.HeatBath)[rule](rng, beta=beta)
    alg._track_counters = False
    sys_.init_("random", rng=rng)
    m.sweep_(sys_, alg, therm)
    N = L * L
    acc = np.zeros((nchains, 4))
    n = 0
    for _ in range(0, sweeps, interval):
        m.sweep_(sys_, alg, interval)
        e = np.asarray(sys_.energy(), dtype=np.float64) / N
        mm = np.asarray(sys_.magnetization(), dtype=np.float64) / N
        acc += np.stack([e, np.abs(mm), mm ** 2, mm ** 4], axis=1)
        n += 1
    return acc / n


@pytest.mark.parametrize("rule", [0, 1, 2])
def test_energy_magnetisation_binder_at_beta_c(m, oracle, rule):
    """C1-like: 2-D Ising L=32 at beta_c.  Tolerance: |difference| < 4.5 combined standard errors
    (independent chains give the error bars directly), for e, |m| and U4 = 1 - <m^4>/(3<m^2>^2)."""
    L, therm, sweeps, interval = 32, 2000, 20000, 10
    ref = oracle.stats_random_site(L, BETA_C, 32, therm, sweeps, interval, nthreads=8, seed=11)
    gpu = _gpu_chain_averages(m, L, BETA_C, rule, 256, therm, sweeps, interval, seed=5)

    def binder(a):     # jackknife over chains
        n = a.shape[0]
        tot2, tot4 = a[:, 2].sum(), a[:, 3].sum()
        full = 1 - (tot4 / n) / (3 * (tot2 / n) ** 2)
        jk = 1 - ((tot4 - a[:, 3]) / (n - 1)) / (3 * ((tot2 - a[:, 2]) / (n - 1)) ** 2)
        return full, math.sqrt((n - 1) / n * ((jk - jk.mean()) ** 2).sum())

    for k, name in ((0, "e"), (1, "|m|")):
        d = gpu[:, k].mean() - ref[:, k].mean()
        err = math.hypot(gpu[:, k].std(ddof=1) / math.sqrt(gpu.shape[0]), ref[:, k].std(ddof=1) / math.sqrt(ref.shape[0]))
        assert abs(d) < 4.5 * err, (name, d, err)
    (ug, eg), (ur, er) = binder(gpu), binder(ref)
    assert abs(ug - ur) < 4.5 * math.hypot(eg, er), ("U4", ug, ur, eg, er)
    assert 0.55 < ug < 0.66          # near the universal critical value ~0.61 for periodic squares


def test_canonical_energy_distribution_vs_exact_8x8(m):
    """P(E) at beta = 0.4 on 8x8 against distribution_exact_ising2D (ising2d_exact.jl:41-42):
    total-variation distance below 0.01 with 2.56e6 samples (statistical floor ~0.003)."""
    L, beta, nch = 8, 0.4, 2048
    exact = _exact_logdos(L)
    es = np.array(sorted(exact))
    logw = np.array([exact[e] for e in es]) - beta * es
    p_exact = np.exp(logw - logw.max())
    p_exact /= p_exact.sum()
    sys_ = m.Ising([L, L], nchains=nch)
    rng = m.PhiloxRNG(2024, 0)
    alg = m.Metropolis(rng, beta=beta)
    alg._track_counters = False
    sys_.init_("random", rng=rng)
    m.sweep_(sys_, alg, 500)
    counts = dict.fromkeys(es.tolist(), 0)
    for _ in range(1250):
        m.sweep_(sys_, alg, 4)
        for e, c in zip(*np.unique(np.asarray(sys_.energy()), return_counts=True)):
            counts[int(e)] += int(c)
    p = np.array([counts[e] for e in es.tolist()], dtype=np.float64)
    p /= p.sum()
    assert 0.5 * np.abs(p - p_exact).sum() < 0.01


def test_wang_landau_logdos_vs_exact_8x8(m):
    """Wang-Landau on 8x8 (user-driven schedule as in the reference: update! halves logf,
    ensembles/wang_landau.jl:23): RMSE of log g(E) - log g(0) against the exact DOS < 0.5."""
    L = 8
    exact = _exact_logdos(L)
    bins = range(-2 * L * L, 2 * L * L + 1, 4)
    sys_ = m.Ising([L, L])
    rng = m.PhiloxRNG(7, 0)
    sys_.init_("random", rng=rng)
    alg = m.WangLandau(rng, bins, logf=1.0)
    while alg.ensemble.logf > 2e-5:
        m.sweep_(sys_, alg, 3000)
        alg.ensemble.update_()
    lw = alg.ensemble.logweight_table
    est = np.array([-(lw[e] - lw[0]) for e in sorted(exact)])
    ref = np.array([exact[e] - exact[0] for e in sorted(exact)])
    rmse = float(np.sqrt(np.mean((est - ref) ** 2)))
    assert rmse < 0.5, rmse
    assert alg.acceptance_rate() > 0.1
