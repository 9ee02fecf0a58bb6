"""The oracle's general-topology extension (IsingGraph / IsingMatrix with fields, ising.jl:86-360) against every value
the reference's own tests hold for it (SpinSystems/test/test_ising.jl:37-147), plus the colouring contract."""
import numpy as np

import mcx_b200 as m


def _ring4(oracle, vals=(1.0, 2.0, 3.0, 4.0), h=0):
    # J[1,2] = 1, J[2,3] = 2, J[3,4] = 3, J[4,1] = 4 (test_ising.jl:38-42), 0-based CSR, ascending neighbours
    a, b, c, d = vals
    return oracle.Graph([0, 2, 4, 6, 8], [1, 3, 0, 2, 1, 3, 0, 2], [a, d, a, b, b, c, d, c], h=h)


def test_sparse_J_known_answers(oracle):
    g = _ring4(oracle)                                    # test_ising.jl:37-60
    assert g.energy() == -10.0 and g.energy(full=True) == -10.0 and g.magnetization() == 4
    dE = g.delta_energy(0)
    assert dE == 10.0 and dE == 2 * g.local_pair_interactions(0)
    g.flip(0)                                             # modify!(sys, 1, dpair, dspin)
    assert g.energy() == 0.0 and g.energy(full=True) == 0.0 and g.magnetization() == 2


def test_fields_known_answers(oracle):
    rp, col = oracle.grid_csr([2, 2])                     # Ising([2, 2], J=1.0, h=0.5), test_ising.jl:68-79
    g = oracle.Graph(rp, col, None, J=1.0, h=0.5)
    pair = sum(g.local_pair_interactions(i) for i in range(4)) / 2
    assert g.energy() == -pair - 0.5 * 4 and g.energy(full=True) == g.energy()
    h_i = [1.0, -1.0, 0.5, 0.0]
    g2 = oracle.Graph(rp, col, None, J=1.0, h=h_i)
    assert g2.energy() == -pair - sum(h_i) and g2.energy(full=True) == g2.energy()


def test_delta_energy_api_consistency(oracle):
    rp, col = oracle.grid_csr([4, 4])                     # test_ising.jl:123-147
    g = oracle.Graph(rp, col, None, J=1, h=0.1)
    i = 2
    dpair, dspin = -2 * 1 * g.local_pair_interactions(i), -2 * int(g.spins[i])
    assert g.delta_energy(i) == -dpair - 0.1 * dspin
    gm = _ring4(oracle, h=0.2)
    i = 1
    dpair, dspin = -2 * gm.local_pair_interactions(i), -2 * int(gm.spins[i])
    assert gm.delta_energy(i) == -dpair - 0.2 * dspin


def test_cached_equals_full_after_random_flips(oracle):
    rng = np.random.default_rng(1)
    rp, col = oracle.grid_csr([6, 5], periodic=False)
    val = rng.integers(-3, 4, size=len(col)).astype(np.float64)
    # symmetrise: value of (i, j) = value of (j, i)
    pos = {}
    for i in range(30):
        for p in range(rp[i], rp[i + 1]):
            pos[(i, int(col[p]))] = p
    for (i, j), p in pos.items():
        if i < j:
            val[pos[(j, i)]] = val[p]
    for g in (oracle.Graph(rp, col, None, J=2.0, h=rng.integers(-2, 3, size=30).astype(float)), oracle.Graph(rp, col, val, h=0.5)):
        for i in rng.integers(0, 30, size=200):
            g.flip(int(i))
        assert g.energy() == g.energy(full=True)          # integer-valued couplings: the running sums are exact


def test_grid_graph_matches_lattice_and_colouring(oracle):
    # the periodic grid graph is the lattice: same energies as the lattice oracle; the greedy colouring of an even
    # periodic grid is the checkerboard
    dims = [6, 4]
    rp, col = oracle.grid_csr(dims)
    g = oracle.Graph(rp, col, None, J=1.0)
    s = oracle.System(oracle.ISING, dims)
    g.init_random(7, 2)
    s.init_random(7, 2)
    assert np.array_equal(g.spins, s.spins) and g.energy() == s.energy()
    colour, ncol = g.colour()
    x, y = np.arange(24) % 6, np.arange(24) // 6
    assert ncol == 2 and np.array_equal(colour, (x + y) & 1)
    # and the host mirror builds the same graph
    e, n = m.grid_graph(dims)
    from mcx_b200.graph_systems import _csr_from_edges
    rp2, col2, _ = _csr_from_edges(e, n)
    assert np.array_equal(rp, rp2) and np.array_equal(col, col2)
    e2, _ = m.grid_graph([2, 2])
    assert len(e2) == 4                                    # a SimpleGraph keeps one edge per pair (ne(grid([2, 2])) == 4)


def test_coloured_sweep_equals_checkerboard_on_the_grid(oracle):
    """on an even periodic grid the coloured sweep with the graph's integer J is the checkerboard sweep of the lattice:
    same classes, same slots (rank inside the class = row * Lx/2 + x >> 1), same draws"""
    dims = [8, 6]
    rp, col = oracle.grid_csr(dims)
    for rule in (oracle.METROPOLIS, oracle.GLAUBER, oracle.HEATBATH):
        g = oracle.Graph(rp, col, None, J=1.0)
        s = oracle.System(oracle.ISING, dims)
        g.init_random(11, 1)
        s.init_random(11, 1)
        ag, as_ = oracle.Alg(rule, 0.4), oracle.Alg(rule, 0.4)
        g.sweep_coloured(ag, 11, 1, 0, 5)
        s.sweep_checkerboard(as_, 11, 1, 0, 5)
        assert np.array_equal(g.spins, s.spins) and ag.accepted == as_.accepted and ag.steps == as_.steps
