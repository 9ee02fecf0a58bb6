"""Pins the CPU oracle against every known-answer value the reference's own tests hold for
the hot path.  Each test cites the reference test it restates."""
import math

import numpy as np
import pytest


def test_philox_known_answers(oracle):
    # Random123 kat vectors for philox4x32-10 (constants curand_philox4x32_x.h:88-91)
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_ising_4x4_bookkeeping(oracle):
    # SpinSystems/test/test_ising.jl:8-35
    s = oracle.System(oracle.ISING, [4, 4])
    assert s.energy() == -32 and s.magnetization() == 16
    assert 2 * s.local_pair_interactions(0) == 8
    assert s.delta_energy(0) == 8
    # force one flip at site 0: Metropolis at beta=0 always accepts
    alg = oracle.Alg(oracle.METROPOLIS, 0.0)
    s.attempt_at(alg, 0, oracle.Rng(1).position(oracle.TAG_SWEEP, 0, 0))
    assert s.energy() == -24 and s.magnetization() == 14 and s.spins[0] == -1
    assert s.energy(full=True) == -24 and s.magnetization(full=True) == 14
    assert alg.steps == 1 and alg.accepted == 1


def test_ising_cached_equals_full_after_updates(oracle):
    # test_ising.jl:149-172 (4x4, beta=0.4, 1 and 51 random updates)
    s = oracle.System(oracle.ISING, [4, 4])
    s.init_random(2024)
    alg = oracle.Alg(oracle.METROPOLIS, 0.4)
    xo = oracle.xoshiro(2024)
    for n in (1, 50):
        s.sweep_random_site(alg, xo, n)
        assert s.energy() == s.energy(full=True)
        assert s.magnetization() == s.magnetization(full=True)
    assert alg.steps == 51


def test_blume_capel_4x4_bookkeeping(oracle):
    # SpinSystems/test/test_blume_capel.jl:8-31 (J=1, D=0.5, all up; site 1 -> 0)
    s = oracle.System(oracle.BLUME_CAPEL, [4, 4], J=1.0, D=0.5)
    assert s.energy() == -24.0 and s.magnetization() == 16
    assert s.delta_energy(0, 0) == 3.5
    sp = s.spins
    sp[0] = 0
    s.spins = sp
    assert s.magnetization() == 15 and s.spin2_sum() == 15
    assert s.energy() == -24.0 + 3.5


def test_blume_capel_cached_equals_full(oracle):
    # test_blume_capel.jl:87-107 (J=1, D=0.2, h=0.1; 100 Metropolis + 100 heat-bath updates)
    s = oracle.System(oracle.BLUME_CAPEL, [4, 4], J=1.0, h=0.1, D=0.2)
    am = oracle.Alg(oracle.METROPOLIS, 0.7)
    ah = oracle.Alg(oracle.HEATBATH, 0.7)
    xo = oracle.xoshiro(7)
    s.sweep_random_site(am, xo, 100)
    s.sweep_random_site(ah, xo, 100)
    assert ah.steps == 100 and am.steps == 100
    assert s.energy() == pytest.approx(s.energy(full=True), abs=1e-9)
    assert s.magnetization() == s.magnetization(full=True)


def test_propose_state(oracle):
    # blume_capel.jl:21-30
    L = oracle.lib()
    assert [L.mcxo_propose_state(1, -1), L.mcxo_propose_state(0, -1)] == [0, 1]
    assert [L.mcxo_propose_state(1, 0), L.mcxo_propose_state(0, 0)] == [-1, 1]
    assert [L.mcxo_propose_state(1, 1), L.mcxo_propose_state(0, 1)] == [-1, 0]


def test_logistic(oracle):
    # test/test_utils.jl:56-67
    f = oracle.lib().mcxo_logistic
    assert f(0.0) == 0.5
    assert abs(f(20.0) - 1.0) < 1e-8 and abs(f(-20.0)) < 1e-8
    for x in (0.3, 2.0, 7.5):
        assert abs(f(x) + f(-x) - 1.0) < 1e-12


def test_exchange_known_answers(oracle):
    # test/test_parallel_ensembles.jl:183-206
    assert oracle.resolve_pair(1, 0, 4) == (True, 1, 2)
    assert oracle.resolve_pair(2, 0, 4) == (True, 1, 1)
    assert oracle.resolve_pair(1, 1, 4) == (False, 0, 0)
    assert oracle.resolve_pair(3, 1, 4) == (True, 2, 2)
    L = oracle.lib()
    lr = L.mcxo_exchange_log_ratio(1.0, 0.5, 0.0, -5.0)
    assert abs(lr - 2.5) < 1e-12
    assert L.mcxo_accept_exchange(lr, 0.0) == 1
    lr2 = L.mcxo_exchange_log_ratio(1.0, 0.5, -5.0, 0.0)
    assert L.mcxo_accept_exchange(lr2, 1.0) == 0
    # update! on two replicas: accept swaps betas and indices, stage toggles
    idx = np.array([1, 2], dtype=np.int64)
    steps = np.zeros(1, dtype=np.int64)
    acc = np.zeros(1, dtype=np.int64)
    betas = np.array([1.0, 0.5])
    stage = oracle.rx_update(0, idx, steps, acc, betas, [0.0, -5.0], [0.0, 0.0])
    assert stage == 1 and list(idx) == [2, 1] and list(betas) == [0.5, 1.0] and steps[0] == 1 and acc[0] == 1
    # odd stage on 2 replicas: no pair
    stage = oracle.rx_update(stage, idx, steps, acc, betas, [0.0, -5.0], [0.0, 0.0])
    assert stage == 0 and steps[0] == 1


def test_set_betas(oracle):
    # test_parallel_ensembles.jl:152-158
    assert list(oracle.set_betas(4, 0.4, 1.0, "uniform")) == [1.0, 0.8, 0.6, 0.4]
    g = oracle.set_betas(5, 0.25, 2.0, "geometric")
    assert g[0] == pytest.approx(2.0, rel=1e-15) and g[-1] == pytest.approx(0.25, rel=1e-15)
    assert np.allclose(g[1:] / g[:-1], g[1] / g[0])


def test_binned_object_index(oracle):
    # test/test_binned_objects.jl:9-20,51: BinnedObject(0:2:10): size 6, bo[4] <-> values[3]
    L = oracle.lib()
    assert L.mcxo_binindex(0, 2, 4) == 3
    assert L.mcxo_binindex(0, 2, 10) == 6
    assert L.mcxo_binindex_f(0.0, 2.0, 4.0) == 3
    # Julia div truncates toward zero
    assert L.mcxo_binindex(-128, 4, -129) == 1


def test_muca_update_and_accept(oracle):
    # test/test_multicanonical.jl:16-48
    f = oracle.Flat(0, 1, 4)
    f.histogram[:] = [0.2, 0.8, 1.1, 2.5]
    f.logweight[:] = [0.5, -1.0, 2.0, 0.0]
    lw0 = f.logweight.copy()
    f.muca_update()
    assert np.allclose(f.logweight, lw0 - np.log([0.2, 0.8, 1.1, 2.5]))
    f.histogram[:] = 0
    alg = oracle.Alg(oracle.METROPOLIS, 0.0)
    n = 0
    rng = np.random.default_rng(0)
    x = 1
    for _ in range(200):
        xn = int(np.clip(x + rng.integers(-1, 2), 0, 3))
        r = f.accept(alg, 0, xn, x, float(rng.random()))
        assert r in (0, 1)
        x = xn if r else x
        n += 1
    assert f.histogram.sum() == alg.steps == n
    # out-of-range -> BoundsError, steps unchanged
    assert f.accept(alg, 0, 7, x, 0.5) == -1 and alg.steps == n
    # zero-count bins leave weights untouched
    f2 = oracle.Flat(0, 1, 3)
    f2.histogram[:] = [0, 3, 0]
    f2.muca_update()
    assert f2.logweight[0] == 0 and f2.logweight[2] == 0 and f2.logweight[1] == -math.log(3)


def test_wang_landau_accept(oracle):
    # test/test_wang_landau.jl:44-63: self-move accepted, lw[x] == w0 - logf
    f = oracle.Flat(0, 1, 5, logf=0.25)
    f.logweight[:] = 1.5
    alg = oracle.Alg(oracle.METROPOLIS, 0.0)
    assert f.accept(alg, 1, 2, 2, 0.3) == 1
    assert f.logweight[2] == 1.5 - 0.25 and alg.steps == 1 and alg.accepted == 1


def test_threshold_tables_match_float_compare(oracle):
    """u < p  <=>  m < ceil(p*2^32): the integer table must reproduce the float decision."""
    rng = np.random.default_rng(5)
    for rule in (oracle.METROPOLIS, oracle.GLAUBER, oracle.HEATBATH):
        for beta in (0.0, 0.2, 0.440686793509772, 1.7):
            T = oracle.build_table(oracle.ISING, rule, 2, beta)
            assert T.size == 10 and T.max() <= 2 ** 32
            for sb in (0, 1):
                for nup in range(5):
                    s = 1 if sb else -1
                    dE = 2.0 * s * (2 * nup - 4)
                    if rule == oracle.HEATBATH:
                        p = oracle.lib().mcxo_logistic(beta * s * dE)
                    elif rule == oracle.GLAUBER:
                        p = oracle.lib().mcxo_logistic(-beta * dE)
                    else:
                        p = 2.0 if -beta * dE > 0 else math.exp(-beta * dE)
                    t = int(T[sb * 5 + nup])
                    ms = np.concatenate([rng.integers(0, 2 ** 32, 64), [0, 2 ** 32 - 1, max(t - 1, 0), min(t, 2 ** 32 - 1)]])
                    for m in ms:
                        assert (float(m) / 2 ** 32 < p) == (int(m) < t)


def test_oracle_reproduces_committed_trajectories(oracle):
    """tests/golden/trajectories.json (written by tests/golden/gen_trajectories.py) freezes RNG layout v1:
    the oracle must still produce every committed trajectory.  The device is held to the same file in
    tests/test_gpu_parity.py::test_device_reproduces_committed_trajectories."""
    import json
    import os
    import _golden_cases as g
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trajectories.json")) as fh:
        gold = json.load(fh)
    assert gold["rng_layout"] == 1
    cs = g.cases()
    assert sorted(g.name_of(c) for c in cs) == sorted(gold["cases"])
    for c in cs:
        assert g.oracle_result(c) == gold["cases"][g.name_of(c)], g.name_of(c)
