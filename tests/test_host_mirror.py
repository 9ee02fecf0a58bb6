"""CPU tests of the host-side mirror of the reference API (no GPU): PhiloxRNG, BinnedObject,
ensembles, scalar accept!, replica-exchange helpers -- the known answers of SURVEY.md 8c again, this
time through the product's Python layer, and agreement of the host rule tables / RNG with the oracle."""
import math

import numpy as np
import pytest

import mcx_b200 as m


def test_philox_matches_oracle_and_kat(oracle):
    assert m.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    r = m.PhiloxRNG(0x1234567890abcdef, chain=7)
    o = oracle.Rng(0x1234567890abcdef, 7)
    for tag, t, q in ((0, 5, 13), (3, 2 ** 40 + 9, 123456789), (0, 0, 0)):
        r.position(tag, t, q)
        o.position(tag, t, q)
        assert r.rand() == o.rand() and r.rand_bool() == o.rand_bool() and r.rand() == o.rand()
    assert m.exchange_u(99, 3, 17) == oracle.lib().mcxo_exchange_u(99, 3, 17)


def test_rule_tables_match_oracle(oracle):
    for model in (0, 1):
        for rule in (0, 1, 2):
            for nd in (1, 2, 3):
                for beta, J, h, D in ((0.44, 1, 0, 0), (0.7, 1.0, 0.1, 0.2), (1.3, 2, 0, 0.5), (30.0, 1, 0, 0)):
                    assert np.array_equal(m.build_table(model, rule, nd, beta, J, h, D),
                                          oracle.build_table(model, rule, nd, beta, J, h, D))


def test_binned_object():
    # test/test_binned_objects.jl
    bo = m.BinnedObject(range(0, 11, 2), 0.0)
    assert bo.size == (6,)
    bo[4] = 2.5
    assert bo.values[2] == 2.5 and bo(4) == 2.5
    with pytest.raises(IndexError):
        bo[12]
    c = m.BinnedObject([0.0, 1.0, 2.0, 3.0], 0.0)
    c[1.2] = 7.0
    assert c.values[1] == 7.0
    assert m.get_centers(bo) == [0, 2, 4, 6, 8, 10]
    assert m.BinnedObject(range(-128, 129, 4), 0.0)[-128] == 0.0
    with pytest.raises(ValueError):
        m.BinnedObject([0, 1, 3], 0.0)


def test_scalar_accept_and_counters():
    class Fixed:
        def __init__(self, u):
            self.u = u

        def rand(self):
            return self.u
    a = m.Metropolis(Fixed(0.5), beta=1.0)
    assert a.accept_(-1.0) is True and a.steps == 1 and a.accepted == 1      # log_ratio > 0: no draw needed
    assert a.accept_(1.0) is False and a.steps == 2                          # exp(-1) = 0.37 < 0.5
    assert a.accept_(0.5) is True                                            # exp(-0.5) = 0.61 > 0.5
    assert m.acceptance_rate(a) == pytest.approx(2 / 3)
    a.reset_()
    assert a.steps == 0 and a.accepted == 0
    g = m.Glauber(Fixed(0.49), beta=1.0)
    assert g.accept_(0.0) is True and g.steps == 1                           # logistic(0) = 0.5
    assert m.logistic(0.0) == 0.5 and abs(m.logistic(20.0) - 1) < 1e-8
    with pytest.raises(ValueError):
        m.BoltzmannEnsemble()
    assert m.BoltzmannEnsemble(T=2.0).beta == 0.5


def test_multicanonical_and_wang_landau_host():
    # test/test_multicanonical.jl:16-48, test/test_wang_landau.jl:44-63
    ens = m.MulticanonicalEnsemble(range(0, 4))
    ens.histogram.values[:] = [0.2, 0.8, 1.1, 2.5]
    lw0 = ens.logweight_table.values.copy()
    ens.update_()
    assert np.allclose(ens.logweight_table.values, lw0 - np.log([0.2, 0.8, 1.1, 2.5]))
    alg = m.Multicanonical(m.PhiloxRNG(1), range(0, 4))
    alg.rng.position(3, 0, 0)
    n0 = alg.steps
    with pytest.raises(IndexError):
        alg.accept_(7, 1)
    assert alg.steps == n0
    assert alg.accept_(2, 2) in (True, False) and alg.ensemble.histogram.values.sum() == 1
    wl = m.WangLandau(m.PhiloxRNG(2), range(0, 5), init=1.5, logf=0.25)
    wl.rng.position(3, 0, 0)
    assert wl.accept_(2, 2) is True and wl.ensemble.logweight_table[2] == 1.5 - 0.25
    wl.ensemble.update_()
    assert wl.ensemble.logf == 0.125


def test_exchange_helpers():
    # test/test_parallel_ensembles.jl:152-206
    assert m.set_betas(4, 0.4, 1.0, "uniform") == [1.0, 0.8, 0.6, 0.4]
    g = m.set_betas(5, 0.25, 2.0, "geometric")
    assert g[0] == pytest.approx(2.0) and g[-1] == pytest.approx(0.25)
    from mcx_b200.parallel import _resolve_pair
    assert _resolve_pair(1, 0, 4) == (True, 1, 2) and _resolve_pair(2, 0, 4) == (True, 1, 1)
    assert _resolve_pair(1, 1, 4) == (False, 0, 0) and _resolve_pair(3, 1, 4) == (True, 2, 2)
    ai, aj = m.Metropolis(m.PhiloxRNG(21), beta=1.0), m.Metropolis(m.PhiloxRNG(22), beta=0.5)
    assert abs(m.exchange_log_ratio(ai.ensemble, aj.ensemble, 0.0, -5.0) - 2.5) < 1e-12
    assert m.attempt_exchange_pair_(ai, aj, 0.0, -5.0, 0.0) and ai.ensemble.beta == 0.5 and aj.ensemble.beta == 1.0
    bi, bj = m.Metropolis(m.PhiloxRNG(23), beta=1.0), m.Metropolis(m.PhiloxRNG(24), beta=0.5)
    assert not m.attempt_exchange_pair_(bi, bj, -5.0, 0.0, 1.0) and bi.ensemble.beta == 1.0
    with pytest.raises(ValueError):
        m.attempt_exchange_pair_(bi, bj, 0.0, 0.0, float("nan"))
    with pytest.raises(ValueError):
        m.ParallelTempering([1.0])
    pt = m.ParallelTempering([1.0, 0.5], seed=77)
    assert pt.size == 2 and list(pt.index()) == [1, 2] and pt.acceptance_rate() == 0.0
    with pytest.raises(ValueError):
        pt.update_([1.0])
    # merge / distribute on two chains (test_parallel_ensembles.jl:96-113)
    algs = [m.Multicanonical(m.PhiloxRNG(1, c), range(0, 4)) for c in range(2)]
    algs[0].ensemble.histogram.values[:] = [1, 2, 3, 4]
    algs[1].ensemble.histogram.values[:] = [4, 3, 2, 1]
    pc = m.ParallelMulticanonical(m.ThreadsBackend(2), algs)
    m.merge_histograms_(pc)
    assert list(algs[0].ensemble.histogram.values) == [5, 5, 5, 5] and list(algs[1].ensemble.histogram.values) == [4, 3, 2, 1]
    algs[0].ensemble.logweight_table.values[:] = [0.1, 0.2, 0.3, 0.4]
    m.distribute_logweight_(pc)
    assert list(algs[1].ensemble.logweight_table.values) == [0.1, 0.2, 0.3, 0.4]


def test_exact_dos_api():
    # SpinSystems/test/test_ising.jl:174-191
    b = m.logdos_exact_ising2D(8)
    assert b[-128] == math.log(2) and abs(b[0] - 42.41274640460084) < 1e-12 and math.isnan(b[-124])
    v = m.logdos_exact_ising2D(8, format="vector")
    assert v[0] == (-128, math.log(2)) and v[-1] == (128, math.log(2))
    with pytest.raises(RuntimeError):
        m.logdos_exact_ising2D(10)
    d = m.distribution_exact_ising2D(8, 0.4)
    assert abs(np.nansum(d.values) - 1.0) < 1e-12


def test_tau_int_and_retuning():
    # src/measurements/autocorrelations.jl:28-65 on an AR(1) series: tau = 1/2 + phi/(1-phi)
    rng = np.random.default_rng(0)
    x = np.zeros(40000)
    for i in range(1, x.size):
        x[i] = 0.8 * x[i - 1] + rng.normal()
    assert abs(m.tau_int(x) - 4.5) < 0.6
    assert m.tau_int(np.ones(10)) == 0.5
    with pytest.raises(ValueError):
        m.tau_int([1.0])
    with pytest.raises(ValueError):
        m.tau_int(x, max_lag=x.size)
    taus = m.integrated_autocorrelation_times([x, x[:3], [1.0]], min_points=4)
    assert taus[0] > 3 and math.isnan(taus[1]) and math.isnan(taus[2])
    from mcx_b200.measurements import _retune_exchange_sweeps_
    assert _retune_exchange_sweeps_([0, 0, 0], [1.0, 2.0, float("nan")], 100, 10, 150) == [67, 133, 100]
    pt = m.ParallelTempering([1.0, 0.5], seed=3)
    sweeps = [10, 10]
    samples = [(1, float(v)) for v in x[:2000]] + [(2, float(v)) for v in rng.normal(size=2000)]
    m.optimize_exchange_interval_(pt, samples, sweeps, base_sweeps=100, min_points=400)
    assert sweeps[0] > sweeps[1]


def test_checkpoint_session_roundtrip(tmp_path):
    # src/infrastructure/checkpointing.jl:48-111
    f = str(tmp_path / "run" / "ckpt.mcx")
    ck = m.init_checkpoint(f, {"rng": m.PhiloxRNG(42, 3), "x": 1.5}, sweep=0)
    m.checkpoint_(ck, sweep=100)
    r = m.restore_checkpoint(f)
    assert r.sweep == 100 and r.x == 1.5 and r.rng.seed == 42 and r.rng.chain == 3
    m.finalize_(ck)
    import os
    assert not os.path.exists(f)


def test_checkpoint_kwargs_do_not_leak_into_the_session(tmp_path):
    """checkpoint!(ckpt; kwargs...) merges the keywords into the written snapshot only (checkpointing.jl:48-56)"""
    f = str(tmp_path / "c.mcx")
    ck = m.init_checkpoint(f, {"x": 1.5})
    m.checkpoint_(ck, sweep=7)
    assert m.restore_checkpoint(f).sweep == 7
    with pytest.raises(AttributeError):
        ck.sweep


def test_philox_rng_many_draws_do_not_alias_other_streams():
    """more than 128 draws at one position: the 8-bit plane field must not spill into the tag field"""
    rng = m.PhiloxRNG(5, 1).position(0, 3, 17)
    first = [rng.rand() for _ in range(300)]
    assert len(set(first)) == 300
    other = m.PhiloxRNG(5, 1).position(1, 3, 17)      # tag 1 = EXCHANGE stream
    assert rng.tag == 0 and all(0.0 <= u < 1.0 for u in first)
    again = m.PhiloxRNG(5, 1).position(0, 3, 17)
    assert [again.rand() for _ in range(300)] == first


def test_graph_host_helpers():
    """graph_systems.py host side: Graphs.SimpleGraphs.grid numbering and the symmetric CSR a SimpleGraph / a sparse J gives"""
    import numpy as np
    import mcx_b200 as m
    from mcx_b200.graph_systems import _csr_from_edges
    e, n = m.grid_graph([3, 2], periodic=False)          # sites i = x + 3 y
    assert n == 6 and sorted(map(tuple, e.tolist())) == [(0, 1), (0, 3), (1, 2), (1, 4), (2, 5), (3, 4), (4, 5)]
    e, n = m.grid_graph([4, 4], periodic=True)
    assert len(e) == 32 and n == 16                      # 2 N edges on a periodic square lattice
    e2, _ = m.grid_graph([2, 2], periodic=True)
    assert len(e2) == 4                                  # a SimpleGraph keeps one edge per pair (ne(grid([2, 2])) == 4)
    rowptr, col, val = _csr_from_edges([[0, 1], [1, 2], [2, 3], [3, 0]], 4, [1.0, 2.0, 3.0, 4.0])
    assert rowptr.tolist() == [0, 2, 4, 6, 8] and col.tolist() == [1, 3, 0, 2, 1, 3, 0, 2]
    assert val.tolist() == [1.0, 4.0, 1.0, 2.0, 2.0, 3.0, 4.0, 3.0]          # ascending neighbours, J_ij == J_ji
    import pytest
    with pytest.raises(AssertionError):
        _csr_from_edges([[0, 1], [1, 2]], 3, [1.0])                          # ne(graph) != length(J) (ising.jl:384)
    with pytest.raises(IndexError):
        _csr_from_edges([[0, 5]], 3)
