"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, bit-exact.

Integer/byte work: spins, sum s_i s_j, sum s, accepted counts must be identical (tolerance 0)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BETA_C = 0.440686793509772


@pytest.fixture(scope="module")
def m():
    import mcx_b200
    mcx_b200.lib()
    return mcx_b200


def _oracle_run(oracle, model, dims, rule, beta, J, h, D, seed, chain, nsweeps, spins0=None):
    s = oracle.System(model, dims, J=float(J), h=float(h), D=float(D))
    if spins0 is None:
        s.init_random(seed, chain)
    else:
        s.spins = spins0
    a = oracle.Alg(rule, beta)
    s.sweep_checkerboard(a, seed, chain, 0, nsweeps)
    return s, a


def _make_alg(m, rule, beta, seed, chain):
    rng = m.PhiloxRNG(seed, chain)
    if rule == 0:
        return m.Metropolis(rng, beta=beta)
    if rule == 1:
        return m.Glauber(rng, beta=beta)
    return m.HeatBath(rng, beta=beta)


def _check(sys_gpu, s_or, a_or, alg, nsweeps):
    assert np.array_equal(sys_gpu.spins, s_or.spins)
    assert sys_gpu.pair_sum() == s_or.pair_count()
    assert sys_gpu.magnetization() == s_or.magnetization(full=True)
    assert sys_gpu.magnetization() == sys_gpu.magnetization(full=True)
    assert sys_gpu.energy() == sys_gpu.energy(full=True)
    assert alg.steps == a_or.steps == nsweeps * s_or.N
    if hasattr(alg, "accepted"):
        assert alg.accepted == a_or.accepted


@pytest.mark.parametrize("L", [8, 16, 64, 256])
@pytest.mark.parametrize("rule", [0, 1, 2])
def test_ising2d_bit_exact(m, oracle, L, rule):
    for beta, seed in ((BETA_C, 42), (0.2, 7), (1.0, 123456789012345)):
        nsweeps = 12 if L <= 64 else 6
        s_or, a_or = _oracle_run(oracle, oracle.ISING, [L, L], rule, beta, 1, 0, 0, seed, 3, nsweeps)
        sys_ = m.Ising([L, L])
        alg = _make_alg(m, rule, beta, seed, 3)
        sys_.init_("random", rng=alg.rng)
        m.sweep_(sys_, alg, nsweeps)
        _check(sys_, s_or, a_or, alg, nsweeps)


@pytest.mark.parametrize("L", [64, 128, 512])
def test_fast_kernel_equals_generic_kernel(m, L):
    """same trajectories from the vectorised and the generic kernels, any strip height"""
    outs = []
    keys = ("MCX_FORCE_GENERIC", "MCX_ROWS_PER_STRIP", "MCX_VARIANT", "MCX_RESIDENT", "MCX_FULL")
    envs = [{"MCX_FORCE_GENERIC": "1"}, {}, {"MCX_RESIDENT": "0"}, {"MCX_FULL": "0"}, {"MCX_FULL": "1"}, {"MCX_FORCE_GENERIC": "2"},
            {"MCX_ROWS_PER_STRIP": "2"}, {"MCX_ROWS_PER_STRIP": "64"}]
    envs += [{"MCX_VARIANT": str(v), "MCX_ROWS_PER_STRIP": r} for v in (0, 3) for r in ("4", "16")]
    for env in envs:
        for k in keys:
            os.environ.pop(k, None)
        os.environ.update(env)
        sys_ = m.Ising([L, L])
        alg = _make_alg(m, 0, BETA_C, 99, 0)
        sys_.init_("random", rng=alg.rng)
        m.sweep_(sys_, alg, 10)
        outs.append((sys_.spins.copy(), sys_.pair_sum(), sys_.magnetization(), alg.accepted))
    for k in keys:
        os.environ.pop(k, None)
    for o in outs[1:]:
        assert np.array_equal(o[0], outs[0][0]) and o[1:] == outs[0][1:]


@pytest.mark.parametrize("rule,track", [(0, False), (0, True), (2, False), (1, True)])
def test_graph_replayed_sweep_series_equals_launches(m, rule, track):
    """a long series of row-band sweeps replayed from a CUDA graph (kernels add a device clock to their half-sweep index)
    gives the trajectory of the same launches queued one by one, over several calls"""
    keys = ("MCX_BANDS", "MCX_SWEEP_GRAPH", "MCX_RESIDENT")

    def run(env):
        for k in keys:
            os.environ.pop(k, None)
        os.environ.update(env)
        s = m.Ising([1024, 1024])
        s.set_tracking(track)
        alg = _make_alg(m, rule, BETA_C, 5, 0)
        s.init_("random", rng=alg.rng)
        l0 = s.ctx.launch_count()
        for n in (30, 3, 17):
            m.sweep_(s, alg, n)
        out = (s.spins.copy(), np.array(s.pair_sum()), np.array(s.magnetization()), np.array(s.accepted()), s.energy(full=True))
        return out, s.ctx.launch_count() - l0

    try:
        ref, l_ref = run({"MCX_BANDS": "4", "MCX_SWEEP_GRAPH": "0", "MCX_RESIDENT": "0"})
        got, l_got = run({"MCX_BANDS": "4", "MCX_SWEEP_GRAPH": "4", "MCX_RESIDENT": "0"})      # 4 sweeps per replay (default 32)
        assert l_got != l_ref, "the graph replay was not taken"
        assert all(np.array_equal(a, b) for a, b in zip(ref, got))
    finally:
        for k in keys:
            os.environ.pop(k, None)


@pytest.mark.parametrize("dims,nchains,envs", [
    ([1024, 1024], 1, [{"MCX_BANDS": "2"}, {"MCX_BANDS": "4"}, {"MCX_BANDS": "8"}]),
    ([2048, 512], 1, [{"MCX_BANDS": "4"}]),
    ([512, 512], 6, [{"MCX_GROUPS": "2"}, {"MCX_GROUPS": "4"}, {"MCX_GROUPS": "6"}]),
    ([64, 32, 32], 1, [{"MCX_BANDS": "4"}, {"MCX_BANDS": "8"}, {"MCX_BANDS": "16"}]),        # 3-D: bands of z-planes
])
def test_bands_and_groups_equal_plain_launches(m, dims, nchains, envs):
    """a half-sweep issued as several overlapping launches (row bands of one lattice, chain groups of a
    batch) gives the trajectory of the single launch, sums and counters included"""
    keys = ("MCX_BANDS", "MCX_GROUPS", "MCX_RESIDENT")

    def run(rule, track, env):
        for k in keys:
            os.environ.pop(k, None)
        os.environ.update(env)
        s = m.Ising(dims, nchains=nchains)
        s.set_tracking(track)
        alg = _make_alg(m, rule, BETA_C, 11, 0)
        s.init_("random", rng=alg.rng)
        l0 = s.ctx.launch_count()
        m.sweep_(s, alg, 4)
        m.sweep_(s, alg, 1)
        out = (s.spins.copy(), np.array(s.pair_sum()), np.array(s.magnetization()), np.array(s.accepted()))
        return out, s.ctx.launch_count() - l0

    try:
        for rule in (0, 2):
            for track in (True, False):
                ref, l_ref = run(rule, track, {"MCX_BANDS": "0", "MCX_GROUPS": "0", "MCX_RESIDENT": "0"})
                for env in envs:
                    got, l_got = run(rule, track, dict(env, MCX_RESIDENT="0"))
                    assert l_got > l_ref, "the banded / grouped path was not taken"
                    assert all(np.array_equal(a, b) for a, b in zip(ref, got)), (rule, track, env)
    finally:
        for k in keys:
            os.environ.pop(k, None)


def test_ising2d_untracked_sums_match(m, oracle):
    L, nsweeps = 128, 5
    s_or, a_or = _oracle_run(oracle, oracle.ISING, [L, L], 0, BETA_C, 1, 0, 0, 5, 0, nsweeps)
    sys_ = m.Ising([L, L])
    sys_.set_tracking(False)
    alg = _make_alg(m, 0, BETA_C, 5, 0)
    sys_.init_("random", rng=alg.rng)
    m.sweep_(sys_, alg, nsweeps)
    _check(sys_, s_or, a_or, alg, nsweeps)


def test_ising2d_rectangular_and_field(m, oracle):
    for dims, J, h in (([64, 8], 1, 0), ([8, 64], 1, 0), ([32, 12], 2, 0.3), ([12, 10], 1.5, -0.2), ([48, 10], 1, 0.1)):
        s_or, a_or = _oracle_run(oracle, oracle.ISING, dims, 0, 0.5, J, h, 0, 11, 1, 8)
        sys_ = m.Ising(dims, J=J, h=h)
        alg = _make_alg(m, 0, 0.5, 11, 1)
        sys_.init_("random", rng=alg.rng)
        m.sweep_(sys_, alg, 8)
        assert np.array_equal(sys_.spins, s_or.spins)
        assert sys_.pair_sum() == s_or.pair_count()
        assert alg.accepted == a_or.accepted
        assert sys_.energy(full=True) == pytest.approx(s_or.energy(full=True), abs=1e-9)


@pytest.mark.parametrize("dims", [[8, 6, 4], [16, 8, 6], [32, 4, 4]])
@pytest.mark.parametrize("rule", [0, 1, 2])
def test_ising3d_bit_exact(m, oracle, rule, dims):
    s_or, a_or = _oracle_run(oracle, oracle.ISING, dims, rule, 0.2216, 1, 0, 0, 2024, 0, 10)
    sys_ = m.Ising(dims)
    alg = _make_alg(m, rule, 0.2216, 2024, 0)
    sys_.init_("random", rng=alg.rng)
    m.sweep_(sys_, alg, 10)
    _check(sys_, s_or, a_or, alg, 10)


@pytest.mark.parametrize("dims", [[32, 8, 4], [64, 16, 6], [32, 32, 8], [128, 4, 4], [64, 12, 10]])
@pytest.mark.parametrize("rule", [0, 1, 2])
def test_ising3d_vectorised_kernel(m, oracle, dims, rule):
    """k_ising3d (z-neighbour rows, run-time row parity) against the oracle and the rows-of-8 kernel"""
    for beta, seed, tracking in ((0.2216, 2024, True), (0.35, 5, False), (0.1, 99, True)):
        nsweeps = 6
        outs = []
        for env in ({}, {"MCX_ISING3D": "0"}):
            os.environ.pop("MCX_ISING3D", None)
            os.environ.update(env)
            sys_ = m.Ising(dims, nchains=2)
            sys_.set_tracking(tracking)
            alg = _make_alg(m, rule, beta, seed, 1)
            sys_.init_("random", rng=alg.rng)
            l0 = sys_.ctx.launch_count()
            m.sweep_(sys_, alg, nsweeps)
            outs.append((sys_.spins.copy(), list(sys_.pair_sum()), list(sys_.magnetization()), list(sys_.accepted())))
        os.environ.pop("MCX_ISING3D", None)
        assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1:] == outs[1][1:]
        for c in range(2):
            s_or, a_or = _oracle_run(oracle, oracle.ISING, dims, rule, beta, 1, 0, 0, seed, 1 + c, nsweeps)
            assert np.array_equal(outs[0][0][c], s_or.spins)
            assert outs[0][1][c] == s_or.pair_count() and outs[0][2][c] == s_or.magnetization(full=True)
            if rule != 2:
                assert outs[0][3][c] == a_or.accepted


@pytest.mark.parametrize("dims", [[64, 64], [256, 64], [32, 128], [512, 256]])
@pytest.mark.parametrize("rule", [0, 1, 2])
def test_blume_capel_vectorised_kernel(m, oracle, dims, rule):
    """k_bc2d (packed 15-bit decisions; four Philox blocks per thread-row for Metropolis / Glauber, two and a pair of
    thresholds per draw for the heat bath) against the oracle and against the rows-of-8 kernel, tracked and untracked
    sums, several couplings"""
    for beta, J, D, h, tracking in ((0.8, 1, 0, 0, True), (1.1, 1.0, 0.5, 0.0, False), (0.6, 1.0, 0.2, 0.1, True), (2.5, 1, 1.9, 0, True)):
        nsweeps = 6
        outs = []
        for env in ({}, {"MCX_BC2D": "0"}):
            os.environ.pop("MCX_BC2D", None)
            os.environ.update(env)
            sys_ = m.BlumeCapel(dims, J=J, D=D, h=h)
            sys_.set_tracking(tracking)
            alg = _make_alg(m, rule, beta, 77, 2)
            sys_.init_("random", rng=alg.rng)
            m.sweep_(sys_, alg, nsweeps)
            outs.append((sys_.spins.copy(), sys_.pair_sum(), sys_.magnetization(), sys_.spin2_sum(), getattr(alg, "accepted", None),
                         sys_.accepted()))
        os.environ.pop("MCX_BC2D", None)
        assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1:] == outs[1][1:]
        s_or, a_or = _oracle_run(oracle, oracle.BLUME_CAPEL, dims, rule, beta, J, h, D, 77, 2, nsweeps)
        assert np.array_equal(outs[0][0], s_or.spins)
        assert outs[0][1] == s_or.pair_count() and outs[0][2] == s_or.magnetization(full=True) and outs[0][3] == s_or.spin2_sum()
        if rule != 2:
            assert outs[0][4] == a_or.accepted


@pytest.mark.parametrize("dims", [[64, 16, 12], [32, 8, 24], [128, 32, 16]])
@pytest.mark.parametrize("rule", [0, 1, 2])
def test_blume_capel_3d_vectorised_kernel(m, oracle, dims, rule):
    """k_bc3d (78-entry tables, z-neighbour rows as extra loads) against the oracle and against the rows-of-8 kernel"""
    for beta, J, D, h, tracking in ((0.5, 1, 0, 0, True), (0.7, 1.0, 0.5, 0.1, False), (1.5, 1, 2.8, 0, True)):
        nsweeps = 5
        outs = []
        for env in ({}, {"MCX_BC2D": "0"}):
            os.environ.pop("MCX_BC2D", None)
            os.environ.update(env)
            sys_ = m.BlumeCapel(dims, J=J, D=D, h=h, nchains=2)
            sys_.set_tracking(tracking)
            alg = _make_alg(m, rule, beta, 78, 4)
            sys_.init_("random", rng=alg.rng)
            l0 = sys_.ctx.launch_count()
            m.sweep_(sys_, alg, nsweeps)
            outs.append((sys_.spins.copy(), np.array(sys_.pair_sum()), np.array(sys_.magnetization()), np.array(sys_.spin2_sum()),
                         np.array(sys_.accepted())))
        os.environ.pop("MCX_BC2D", None)
        assert all(np.array_equal(a, b) for a, b in zip(*outs))
        for c in range(2):
            s_or, a_or = _oracle_run(oracle, oracle.BLUME_CAPEL, dims, rule, beta, J, h, D, 78, 4 + c, nsweeps)
            assert np.array_equal(outs[0][0][c], s_or.spins)
            assert outs[0][1][c] == s_or.pair_count() and outs[0][2][c] == s_or.magnetization(full=True) and outs[0][3][c] == s_or.spin2_sum()
            if rule != 2:
                assert outs[0][4][c] == a_or.accepted


def test_blume_capel_vectorised_batched_labels(m, oracle):
    """chains with different tables (labels) and chain ids in one k_bc2d launch"""
    dims, n, nsweeps = [64, 64], 5, 5
    betas = [0.5, 0.8, 1.0, 1.3, 2.0]
    sys_ = m.BlumeCapel(dims, J=1, D=0.3, nchains=n)
    tables = np.stack([m.build_table(1, 0, 2, b, 1.0, 0.0, 0.3) for b in betas])
    sys_.set_rule(0, tables)
    labels = [4, 0, 3, 1, 2]
    sys_.set_labels(labels)
    sys_.set_rng(4711, 0)
    sys_.init_("random", rng=m.PhiloxRNG(4711, 0))
    m.lib().mcx_sweep(sys_.h_lat, nsweeps)
    got = sys_.spins
    for c in range(n):
        s_or, a_or = _oracle_run(oracle, oracle.BLUME_CAPEL, dims, 0, betas[labels[c]], 1.0, 0.0, 0.3, 4711, c, nsweeps)
        assert np.array_equal(got[c], s_or.spins)
        assert sys_.pair_sum()[c] == s_or.pair_count() and sys_.spin2_sum()[c] == s_or.spin2_sum()
        assert sys_.accepted()[c] == a_or.accepted


@pytest.mark.parametrize("dims", [[8, 8], [32, 16], [6, 4, 8], [64, 12], [16, 6, 4]])
@pytest.mark.parametrize("rule", [0, 1, 2])
def test_blume_capel_bit_exact(m, oracle, dims, rule):
    for beta, J, D, h in ((0.8, 1, 0, 0), (1.1, 1.0, 0.5, 0.0), (0.6, 1.0, 0.2, 0.1)):
        s_or, a_or = _oracle_run(oracle, oracle.BLUME_CAPEL, dims, rule, beta, J, h, D, 77, 2, 10)
        sys_ = m.BlumeCapel(dims, J=J, D=D, h=h)
        alg = _make_alg(m, rule, beta, 77, 2)
        sys_.init_("random", rng=alg.rng)
        m.sweep_(sys_, alg, 10)
        assert np.array_equal(sys_.spins, s_or.spins)
        assert sys_.pair_sum() == s_or.pair_count()
        assert sys_.magnetization() == s_or.magnetization(full=True)
        assert sys_.spin2_sum() == s_or.spin2_sum()
        assert sys_.energy() == pytest.approx(s_or.energy(full=True), abs=1e-9)
        if hasattr(alg, "accepted"):
            assert alg.accepted == a_or.accepted


def test_constructor_defaults_and_known_answers(m):
    # SpinSystems/test/test_ising.jl:8-35 and test_blume_capel.jl:8-31 through the device path
    s = m.Ising([4, 4])
    assert s.energy() == -32 and s.magnetization() == 16
    sp = s.spins.copy()
    assert sp.dtype == np.int8 and (sp == 1).all()
    sp[0] = -1
    s.spins = sp
    assert s.energy() == -24 and s.magnetization() == 14 and s.spins[0] == -1
    b = m.BlumeCapel([4, 4], J=1, D=0.5)
    assert b.energy() == -24.0 and b.magnetization() == 16
    sp = b.spins.copy()
    sp[0] = 0
    b.spins = sp
    assert b.magnetization() == 15 and b.energy() == -24.0 + 3.5
    b.init_("zero")
    assert b.energy() == 0 and b.spin2_sum() == 0
    with pytest.raises(ValueError):
        m.Ising([5, 4])
    with pytest.raises(ValueError):
        s.init_("zero")


def test_batched_chains_with_labels(m, oracle):
    """many independent chains in one launch, each with its own table (label) and chain id"""
    L, n, nsweeps = 32, 6, 8
    betas = [0.2, 0.3, 0.4, 0.44, 0.5, 0.7]
    sys_ = m.Ising([L, L], nchains=n)
    tables = np.stack([m.build_table(0, 0, 2, b) for b in betas])
    sys_.set_rule(0, tables)
    labels = [5, 0, 3, 1, 4, 2]
    sys_.set_labels(labels)
    sys_.set_rng(31337, 0)
    sys_.init_("random", rng=m.PhiloxRNG(31337, 0))
    m.lib().mcx_sweep(sys_.h_lat, nsweeps)
    got = sys_.spins
    for c in range(n):
        s_or, a_or = _oracle_run(oracle, oracle.ISING, [L, L], 0, betas[labels[c]], 1, 0, 0, 31337, c, nsweeps)
        assert np.array_equal(got[c], s_or.spins)
        assert sys_.pair_sum()[c] == s_or.pair_count()
        assert sys_.accepted()[c] == a_or.accepted


def test_parallel_tempering_matches_oracle(m, oracle):
    """sweep + exchange rounds: label permutation, per-edge counters and lattices vs the oracle
    (replica_exchange.jl:158-178 restated in oracle.rx_update)"""
    L, n, rounds, seed = 32, 8, 24, 2025
    betas = m.set_betas(n, 0.3, 0.6, "uniform")
    pt = m.ParallelTempering(betas, seed=seed, backend=m.GPUBackend())
    sys_ = m.Ising([L, L], nchains=n)
    pt.attach(sys_)
    sys_.init_("random", rng=m.PhiloxRNG(seed, 0))
    # oracle side
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
        m.sweep_(sys_, pt, 1)
        m.update_(pt)
        for r in range(n):
            a = oracle.Alg(oracle.METROPOLIS, float(beta_of_slot[r]))
            o_sys[r].sweep_checkerboard(a, seed, r, rd, 1)
        xs = [o_sys[r].energy() for r in range(n)]
        us = [oracle.lib().mcxo_exchange_u(seed, r, rd) for r in range(n)]
        stage = oracle.rx_update(stage, idx, steps, acc, beta_of_slot, xs, us)
    assert list(pt.index()) == list(idx)
    assert list(pt.steps) == list(steps) and list(pt.accepted) == list(acc)
    assert pt.stage == stage
    assert acc.sum() > 0
    got = sys_.spins
    for r in range(n):
        assert np.array_equal(got[r], o_sys[r].spins)
        assert pt.algorithm(r).ensemble.beta == beta_of_slot[r]
    assert np.allclose(pt.energies(), [s.energy() for s in o_sys])


def test_pt_run_equals_user_loop(m):
    """ParallelTempering.run_ (one library call for rounds x (sweeps, update!)) == the explicit loop"""
    L, n, seed = 64, 10, 77
    betas = m.set_betas(n, 0.3, 0.6, "uniform")
    outs = []
    for mode in ("loop", "run"):
        for every in (1, 4):
            pt = m.ParallelTempering(betas, seed=seed, backend=m.GPUBackend())
            reps = m.Ising([L, L], nchains=n)
            pt.attach(reps)
            reps.init_("random", rng=m.PhiloxRNG(seed, 0))
            if mode == "loop":
                for _ in range(12):
                    m.sweep_(reps, pt, every)
                    m.update_(pt)
            else:
                pt.run_(reps, 12, every)
            outs.append((mode, every, list(pt.index()), list(pt.steps), list(pt.accepted), reps.spins.copy(),
                         [a.steps for a in pt.replica.algs]))
    for every in (1, 4):
        a = [o for o in outs if o[0] == "loop" and o[1] == every][0]
        b = [o for o in outs if o[0] == "run" and o[1] == every][0]
        assert a[2:5] == b[2:5] and np.array_equal(a[5], b[5]) and a[6] == b[6]
        assert sum(a[4]) > 0


def test_large_lattice_properties(m):
    """L = 4096 (beyond what the oracle sweeps in seconds): size-independent properties"""
    L = 4096
    sys_ = m.Ising([L, L])
    alg = _make_alg(m, 0, BETA_C, 42, 0)
    sys_.init_("random", rng=alg.rng)
    m.sweep_(sys_, alg, 5)
    e_cached, m_cached, acc = sys_.energy(), sys_.magnetization(), alg.accepted
    assert e_cached == sys_.energy(full=True) and m_cached == sys_.magnetization(full=True)
    assert 0 < acc < alg.steps
    sp = sys_.spins
    assert int(sp.astype(np.int64).sum()) == m_cached and set(np.unique(sp)) == {-1, 1}
    # host recomputation of the energy from the downloaded spins
    g = sp.reshape(L, L).astype(np.int32)
    pair = int((g * np.roll(g, 1, 0)).sum() + (g * np.roll(g, 1, 1)).sum())
    assert e_cached == -pair
    # determinism and restart: 5 sweeps == 2 + 3 sweeps with the counter carried over
    sys2 = m.Ising([L, L])
    alg2 = _make_alg(m, 0, BETA_C, 42, 0)
    sys2.init_("random", rng=alg2.rng)
    m.sweep_(sys2, alg2, 2)
    assert sys2.sweep_index == 2
    m.sweep_(sys2, alg2, 3)
    assert np.array_equal(sys2.spins, sp) and alg2.accepted == acc
    # upload/download round trip is the identity
    sys2.spins = sp
    assert np.array_equal(sys2.spins, sp) and sys2.energy() == e_cached


def test_device_reproduces_committed_trajectories(m):
    """The committed golden trajectories (tests/golden/trajectories.json, RNG layout v1) without the oracle
    in the loop: spins (sha256), integer sums, accepted counts and flat-histogram tables, bit for bit;
    energies that involve a Float64 field / crystal-field term to 1e-9."""
    import json
    import os
    import _golden_cases as g
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trajectories.json")) as fh:
        gold = json.load(fh)["cases"]
    for c in g.cases():
        want, got = gold[g.name_of(c)], g.device_result(c, m)
        assert got.pop("energy") == pytest.approx(want.pop("energy"), abs=1e-9), g.name_of(c)
        assert got == want, g.name_of(c)


def test_split_upload_equals_upload(m):
    """mcx_lattice_upload_begin / _commit: the copy runs beside sweeps that are already queued and does not touch the
    lattice until the commit; afterwards the handle behaves as after mcx_lattice_upload."""
    import torch
    L, seed = 256, 9
    rs = np.random.default_rng(1)
    spins1 = (2 * rs.integers(0, 2, L * L) - 1).astype(np.int8)

    def fresh():
        s = m.Ising([L, L])
        a = m.Metropolis(m.PhiloxRNG(seed, 0), beta=0.44)
        s.init_("random", rng=a.rng)
        return s, a

    ref, aref = fresh()
    m.sweep_(ref, aref, 2)
    before = ref.spins.copy()
    ref.spins = spins1                       # plain upload
    ref.set_rng(seed, 10)
    m.sweep_(ref, aref, 3)

    for pinned in (False, True):
        s, a = fresh()
        if pinned:
            buf = torch.empty(L * L, dtype=torch.int8).pin_memory()
            buf.copy_(torch.from_numpy(spins1))
            ptr = buf.data_ptr()
        else:
            buf = spins1.copy()
            ptr = buf.ctypes.data
        s._bind_alg(a)
        m._lib.check(m.lib().mcx_sweep(s.h_lat, 2))      # queued; the copy below must not disturb it
        s.upload_begin(ptr)
        with pytest.raises(AssertionError):
            s.upload_begin(ptr)                           # one pending upload per handle
        with pytest.raises(AssertionError):
            s.spins                                       # download needs the staging buffer
        with pytest.raises(AssertionError):
            s.spins = spins1
        assert s.energy() == int(-_pair_sum_2d(before, L))   # still the swept old lattice
        s.upload_commit()
        with pytest.raises(AssertionError):
            s.upload_commit()
        assert np.array_equal(s.spins, spins1)
        assert s.energy() == s.energy(full=True) == int(-_pair_sum_2d(spins1, L))
        s.set_rng(seed, 10)
        m.sweep_(s, a, 3)
        assert np.array_equal(s.spins, ref.spins) and s.energy() == ref.energy()
        # a second round trip on the same handle (events are reused)
        s.upload_begin(ptr)
        s.upload_commit()
        assert np.array_equal(s.spins, spins1)


def _pair_sum_2d(spins, L):
    a = np.asarray(spins, dtype=np.int64).reshape(L, L)
    return int((a * np.roll(a, 1, 0)).sum() + (a * np.roll(a, 1, 1)).sum())
