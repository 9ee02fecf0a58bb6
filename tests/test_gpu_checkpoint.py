"""Checkpoint/restore of device state and the on-device measurement series:
a restored run must continue the trajectory bit for bit (cf. test/test_checkpointing.jl:100-147)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
BETA_C = 0.440686793509772


@pytest.fixture(scope="module")
def m():
    import mcx_b200
    mcx_b200.lib()
    return mcx_b200


def test_deterministic_restart(m, oracle, tmp_path):
    L = 128
    ref = m.Ising([L, L])
    alg_ref = m.Metropolis(m.PhiloxRNG(42, 2), beta=BETA_C)
    ref.init_("random", rng=alg_ref.rng)
    m.sweep_(ref, alg_ref, 40)

    sys_ = m.Ising([L, L])
    alg = m.Metropolis(m.PhiloxRNG(42, 2), beta=BETA_C)
    sys_.init_("random", rng=alg.rng)
    m.sweep_(sys_, alg, 25)
    ck = m.init_checkpoint(str(tmp_path / "ckpt.mcx"), {"sys": sys_, "alg": alg}, sweep=25)
    del sys_, alg
    r = m.restore_checkpoint(ck.file)
    sys2, alg2 = r.sys, r.alg
    assert r.sweep == 25 and sys2.sweep_index == 25 and alg2.steps == 25 * L * L
    m.sweep_(sys2, alg2, 15)
    assert np.array_equal(sys2.spins, ref.spins)
    assert sys2.energy() == ref.energy() and alg2.accepted == alg_ref.accepted and alg2.steps == alg_ref.steps
    # ... and both are the oracle's trajectory, not merely each other's
    o = oracle.System(oracle.ISING, [L, L])
    o.init_random(42, 2)
    oa = oracle.Alg(oracle.METROPOLIS, BETA_C)
    o.sweep_checkerboard(oa, 42, 2, 0, 40)
    assert np.array_equal(sys2.spins, o.spins) and alg2.accepted == oa.accepted and alg2.steps == oa.steps
    assert sys2.energy() == o.energy(full=True) and sys2.magnetization() == o.magnetization(full=True)
    # the lattice-side counters travel with the checkpoint (importance_sampling.jl:26-27)
    assert int(np.atleast_1d(sys2.accepted())[0]) == oa.accepted
    m.finalize_(ck)


def test_sweep_series_matches_per_sweep_reads(m, oracle):
    L, nch = 64, 3
    a = m.Ising([L, L], nchains=nch)
    b = m.Ising([L, L], nchains=nch)
    alg_a = m.Glauber(m.PhiloxRNG(9, 0), beta=0.4)
    alg_b = m.Glauber(m.PhiloxRNG(9, 0), beta=0.4)
    a.init_("random", rng=alg_a.rng)
    b.init_("random", rng=alg_b.rng)
    series = m.sweep_series_(a, alg_a, 12, interval=3)
    for k in range(12):
        m.sweep_(b, alg_b, 3)
        assert np.array_equal(series["energy"][k], np.asarray(b.energy()))
        assert np.array_equal(series["magnetization"][k], np.asarray(b.magnetization()))
    assert alg_a.accepted == alg_b.accepted and alg_a.steps == alg_b.steps
    assert m.tau_int(series["energy"][:, 0].astype(float)) >= 0.5
    # the series against the oracle: every snapshot of every chain
    acc_total = 0
    for c in range(nch):
        o = oracle.System(oracle.ISING, [L, L])
        o.init_random(9, c)
        oa = oracle.Alg(oracle.GLAUBER, 0.4)
        for k in range(12):
            o.sweep_checkerboard(oa, 9, c, 3 * k, 3)
            assert series["energy"][k, c] == o.energy() and series["magnetization"][k, c] == o.magnetization()
            assert series["accepted"][k, c] == oa.accepted
        acc_total += oa.accepted
    # one algorithm object drives the whole batch: its counters are sums over the chains and the rate stays a rate
    assert alg_a.accepted == acc_total and alg_a.steps == 36 * L * L * nch
    assert 0.0 < alg_a.acceptance_rate() < 1.0


def test_batched_sweep_counters_are_a_rate(m):
    """acceptance_rate(alg) on a batched lattice (importance_sampling.jl:95-101): steps and accepted both count all chains"""
    L, nch = 32, 5
    s = m.Ising([L, L], nchains=nch)
    alg = m.Metropolis(m.PhiloxRNG(3, 0), beta=0.3)
    s.init_("random", rng=alg.rng)
    m.sweep_(s, alg, 4)
    assert alg.steps == 4 * L * L * nch
    assert alg.accepted == int(np.sum(s.accepted()))
    assert 0.0 < alg.acceptance_rate() < 1.0


def test_parallel_tempering_restart(m, oracle, tmp_path):
    """a checkpointed ladder (indices, per-edge counters, stage, exchange round, lattices) continues exactly
    (checkpointing.jl:95-101 restoring replica_exchange.jl:13-19)"""
    L, n, seed = 32, 8, 11
    betas = m.set_betas(n, 0.3, 0.6, "uniform")

    def fresh():
        pt = m.ParallelTempering(betas, seed=seed, backend=m.GPUBackend())
        reps = m.Ising([L, L], nchains=n)
        pt.attach(reps)
        reps.init_("random", rng=m.PhiloxRNG(seed, 0))
        return pt, reps

    ref_pt, ref_reps = fresh()
    ref_pt.run_(ref_reps, 30, 1)
    pt, reps = fresh()
    pt.run_(reps, 13, 1)
    ck = m.init_checkpoint(str(tmp_path / "pt.mcx"), {"pt": pt}, round=13)
    del pt, reps
    r = m.restore_checkpoint(ck.file)
    pt2 = r.pt
    reps2 = pt2._sys
    assert pt2.round == 13 and reps2.sweep_index == 13
    pt2.run_(reps2, 17, 1)
    assert list(pt2.index()) == list(ref_pt.index())
    assert list(pt2.steps) == list(ref_pt.steps) and list(pt2.accepted) == list(ref_pt.accepted)
    assert pt2.stage == ref_pt.stage and pt2.round == ref_pt.round == 30
    assert np.array_equal(reps2.spins, ref_reps.spins)
    assert [a.ensemble.beta for a in pt2.replica.algs] == [a.ensemble.beta for a in ref_pt.replica.algs]
    # per-replica counters (ReplicaExchange credits alg.accepted from the per-chain device counters)
    ref_pt.sync_counters(); pt2.sync_counters()
    assert [a.accepted for a in pt2.replica.algs] == [a.accepted for a in ref_pt.replica.algs]
    assert all(0 < a.accepted < a.steps for a in ref_pt.replica.algs)
    assert sum(ref_pt.accepted) > 0
    m.finalize_(ck)


def test_device_tau_int_matches_host_estimator(m):
    """mcx_series_tau_int reduces the on-device series to one tau_int per chain (autocorrelations.jl:28-65); it must agree
    with the reference's estimator applied to the copied-back series (deterministic block sums: tolerance 1e-10 relative),
    give 0.5 for a constant signal (test/test_measurements.jl:154) and reject the same arguments."""
    L, nch, n = 32, 5, 600
    sys_ = m.Ising([L, L], nchains=nch)
    alg = m.Metropolis(m.PhiloxRNG(17, 0), beta=0.42)
    sys_.init_("random", rng=alg.rng)
    m.sweep_(sys_, alg, 50)
    series = m.sweep_series_(sys_, alg, n, interval=2)
    for obs, key, f in (("energy", "energy", lambda x: x), ("magnetization", "magnetization", lambda x: x),
                        ("abs_magnetization", "magnetization", np.abs)):
        for max_lag, c in ((None, 5.0), (50, 5.0), (None, 2.5)):
            dev = m.series_tau_int_(sys_, n, obs, max_lag=max_lag, c=c)
            host = [m.integrated_autocorrelation_time(f(np.asarray(series[key][:, ch], dtype=np.float64)), max_lag=max_lag, c=c)
                    for ch in range(nch)]
            assert np.allclose(dev, host, rtol=1e-10, atol=0), (obs, max_lag, c, dev, host)
            assert np.all(dev >= 0.5)
    # shorter prefix of the same series
    dev = m.series_tau_int_(sys_, 100, "energy")
    host = [m.integrated_autocorrelation_time(np.asarray(series["energy"][:100, ch], dtype=np.float64)) for ch in range(nch)]
    assert np.allclose(dev, host, rtol=1e-10, atol=0)
    # a frozen lattice (beta -> infinity from all-up): constant series, tau_int == 0.5
    cold = m.Ising([L, L])
    calg = m.Metropolis(m.PhiloxRNG(1, 0), beta=50.0)
    m.sweep_series_(cold, calg, 64, interval=1)
    assert m.series_tau_int_(cold, 64, "energy")[0] == 0.5
    with pytest.raises(ValueError):
        m.series_tau_int_(sys_, n, "energy", max_lag=n // 2 + 1)
    with pytest.raises(ValueError):
        m.series_tau_int_(sys_, n, "energy", c=0.0)
    with pytest.raises(AssertionError):
        m.series_tau_int_(sys_, n + 1, "energy")
