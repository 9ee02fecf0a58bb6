"""GPU parity for general topologies (IsingGraph / IsingMatrix with site fields): the coloured sweeps of
mcx_graph_* against the oracle's restatement of ising.jl:86-360, bit-exact spins and counters; energies exact for
exactly representable couplings."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m():
    import mcx_b200
    mcx_b200.lib()
    return mcx_b200


def _alg(m, rule, beta, seed, chain):
    rng = m.PhiloxRNG(seed, chain)
    return (m.Metropolis, m.Glauber, m.HeatBath)[rule](rng, beta=beta)


def _random_graph(rng, n, m_edges):
    e = set()
    while len(e) < m_edges:
        i, j = (int(v) for v in rng.integers(0, n, size=2))
        if i != j:
            e.add((min(i, j), max(i, j)))
    return np.array(sorted(e), dtype=np.int64)


def _oracle_of(oracle, sys_):
    return oracle.Graph(sys_.rowptr, sys_.col, sys_.val, J=float(sys_.J) if sys_.val is None else 0.0, h=sys_.h)


def _compare(m, oracle, sys_, rule, beta, seed, nsweeps):
    ncol, colour = sys_.colours()
    ref = _oracle_of(oracle, sys_)
    rc, rn = ref.colour()
    assert rn == ncol and np.array_equal(rc, colour)
    k = sys_.nchains
    alg = _alg(m, rule, beta, seed, 3)
    sys_.init_("random", rng=alg.rng)
    m.sweep_(sys_, alg, nsweeps - 2)
    m.sweep_(sys_, alg, 2)
    spins = np.asarray(sys_.spins).reshape(k, sys_.N)
    E, M = np.atleast_1d(sys_.energy()), np.atleast_1d(sys_.magnetization())
    acc, steps = 0, 0
    for c in range(k):
        g = _oracle_of(oracle, sys_)
        g.init_random(seed, 3 + c)
        a = oracle.Alg(rule, beta)
        g.sweep_coloured(a, seed, 3 + c, 0, nsweeps, rc, rn)
        assert np.array_equal(spins[c], g.spins), (rule, c)
        assert E[c] == g.energy(full=True) and M[c] == g.magnetization()
        acc += a.accepted
        steps += a.steps
    assert alg.steps == steps
    if hasattr(alg, "accepted"):
        assert alg.accepted == acc


def test_sparse_J_known_answers_on_device(m):
    J = np.zeros((4, 4))
    J[0, 1] = J[1, 0] = 1.0
    J[1, 2] = J[2, 1] = 2.0
    J[2, 3] = J[3, 2] = 3.0
    J[3, 0] = J[0, 3] = 4.0
    sys_ = m.IsingMatrix(J)                              # SpinSystems/test/test_ising.jl:37-50
    assert sys_.energy() == -10.0 and sys_.energy(full=True) == -10.0 and sys_.magnetization() == 4
    sp = sys_.spins.copy()
    sp[0] = -1
    sys_.spins = sp                                      # the state after modify!(sys, 1, ...) (:55-60)
    assert sys_.energy() == 0.0 and sys_.magnetization() == 2
    bad = np.zeros((3, 3))
    bad[0, 1], bad[1, 0] = 1.0, 0.5
    with pytest.raises(AssertionError):                  # @test_throws AssertionError Ising(J_bad) (:62-65)
        m.IsingMatrix(bad)


def test_field_known_answers_on_device(m):
    sys_ = m.Ising([2, 2], J=1.0, h=[1.0, -1.0, 0.5, 0.0])      # test_ising.jl:75-79 (4 edges: a SimpleGraph merges duplicates)
    assert sys_.energy() == -4.0 - 0.5
    assert m.Ising([2, 2], J=1.0, h=0.5, periodic=False).energy() == -4.0 - 2.0


@pytest.mark.parametrize("rule", [0, 1, 2])
def test_periodic_grid_graph_equals_lattice_kernels(m, oracle, rule):
    """the graph path on the periodic grid graph reproduces the lattice path (and the oracle) bit for bit"""
    dims = [32, 16]
    e, n = m.grid_graph(dims)
    g = m.IsingGraph(e, n, J=1)
    lat = m.Ising(dims)
    ag, al = _alg(m, rule, 0.44, 5, 0), _alg(m, rule, 0.44, 5, 0)
    g.init_("random", rng=ag.rng)
    lat.init_("random", rng=al.rng)
    m.sweep_(g, ag, 6)
    m.sweep_(lat, al, 6)
    assert np.array_equal(g.spins, lat.spins) and g.energy() == lat.energy() and g.magnetization() == lat.magnetization()
    if rule != 2:
        assert ag.accepted == al.accepted


@pytest.mark.parametrize("rule", [0, 1, 2])
def test_random_graphs_bit_exact(m, oracle, rule):
    rng = np.random.default_rng(3)
    n = 600
    e = _random_graph(rng, n, 1800)
    # one global coupling, no field / uniform field / field vector
    _compare(m, oracle, m.IsingGraph(e, n, J=1), rule, 0.3, 21, 8)
    _compare(m, oracle, m.IsingGraph(e, n, J=-1.5, h=0.25, nchains=3), rule, 0.2, 22, 6)
    _compare(m, oracle, m.IsingGraph(e, n, J=2, h=rng.integers(-4, 5, size=n) * 0.25), rule, 0.15, 23, 6)
    # one coupling per edge (the reference's IsingMatrix), dyadic values of both signs: a spin glass
    Jv = rng.integers(-8, 9, size=len(e)) * 0.125
    _compare(m, oracle, m.IsingGraph(e, n, J=Jv), rule, 0.5, 24, 8)
    _compare(m, oracle, m.IsingGraph(e, n, J=Jv, h=rng.integers(-4, 5, size=n) * 0.5, nchains=2), rule, 0.4, 25, 6)


def test_open_boundaries_and_arbitrary_couplings(m, oracle):
    # Ising(dims; periodic=false) in 2-D and 3-D, odd sizes included (no checkerboard restriction on this path)
    _compare(m, oracle, m.Ising([9, 7], periodic=False), 0, 0.44, 31, 8)
    _compare(m, oracle, m.Ising([5, 4, 3], J=1, h=0.1, periodic=False), 1, 0.3, 32, 6)
    # couplings that are not exactly representable: trajectories still bit-exact (same Float64 expression per attempt)
    rng = np.random.default_rng(4)
    e = _random_graph(rng, 300, 700)
    sys_ = m.IsingGraph(e, 300, J=rng.normal(size=len(e)), h=rng.normal(size=300) * 0.3)
    ref = _oracle_of(oracle, sys_)
    alg = _alg(m, 0, 0.7, 33, 0)
    sys_.init_("random", rng=alg.rng)
    m.sweep_(sys_, alg, 10)
    ref.init_random(33, 0)
    a = oracle.Alg(0, 0.7)
    ref.sweep_coloured(a, 33, 0, 0, 10)
    assert np.array_equal(sys_.spins, ref.spins) and alg.accepted == a.accepted
    assert abs(sys_.energy() - ref.energy(full=True)) <= 1e-9 * max(1.0, abs(ref.energy(full=True)))   # summation order differs
