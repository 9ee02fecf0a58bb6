"""GPU parity of the shared-memory-resident sweep series (k_resident.cu): whole nsweeps-series in one
cluster launch must give the trajectories of the streaming kernel and of the CPU oracle, bit for bit,
for every cluster size, rule, tracking mode and label assignment."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BETA_C = 0.440686793509772
KEYS = ("MCX_RESIDENT", "MCX_RESIDENT_CLUSTER", "MCX_FORCE_GENERIC")


@pytest.fixture(scope="module")
def m():
    import mcx_b200
    mcx_b200.lib()
    return mcx_b200


@pytest.fixture(autouse=True)
def _clean_env():
    for k in KEYS:
        os.environ.pop(k, None)
    yield
    for k in KEYS:
        os.environ.pop(k, None)


def _alg(m, rule, beta, seed, chain=0):
    rng = m.PhiloxRNG(seed, chain)
    return (m.Metropolis, m.Glauber, m.HeatBath)[rule](rng, beta=beta)


def _run(m, dims, rule, beta, seed, nsweeps, nchains=1, tracking=True, env=None, calls=1):
    os.environ.update(env or {})
    sys_ = m.Ising(dims, nchains=nchains)
    sys_.set_tracking(tracking)
    alg = _alg(m, rule, beta, seed)
    sys_.init_("random", rng=alg.rng)
    before = sys_.ctx.launch_count()
    for _ in range(calls):
        m.sweep_(sys_, alg, nsweeps)
    launches = sys_.ctx.launch_count() - before
    out = (sys_.spins.copy(), np.array(sys_.pair_sum()), np.array(sys_.magnetization()), np.array(sys_.accepted()))
    for k in KEYS:
        os.environ.pop(k, None)
    return out, launches


def _same(a, b):
    return all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("dims,cluster", [([256, 256], 1), ([256, 256], 2), ([256, 256], 4), ([256, 256], 8),
                                          ([512, 512], 2), ([512, 512], 8), ([1024, 1024], 8), ([512, 128], 1),
                                          ([64, 1024], 4)])
@pytest.mark.parametrize("rule", [0, 1, 2])
def test_resident_equals_streaming(m, dims, cluster, rule):
    nsweeps = 7
    ref, l_ref = _run(m, dims, rule, BETA_C, 17, nsweeps, nchains=3, env={"MCX_RESIDENT": "0"})
    got, l_got = _run(m, dims, rule, BETA_C, 17, nsweeps, nchains=3,
                      env={"MCX_RESIDENT": "1", "MCX_RESIDENT_CLUSTER": str(cluster)})
    assert l_ref >= 2 * nsweeps
    assert l_got == 1, "the resident kernel was not used"
    assert _same(ref, got)


def test_resident_matches_oracle(m, oracle):
    L, nsweeps, seed = 256, 6, 4242
    for rule in (0, 1, 2):
        for tracking in (True, False):
            got, launches = _run(m, [L, L], rule, BETA_C, seed, nsweeps, tracking=tracking, env={"MCX_RESIDENT": "1"})
            s = oracle.System(oracle.ISING, [L, L])
            s.init_random(seed, 0)
            a = oracle.Alg(rule, BETA_C)
            s.sweep_checkerboard(a, seed, 0, 0, nsweeps)
            assert np.array_equal(got[0], s.spins)
            assert got[1] == s.pair_count() and got[2] == s.magnetization(full=True)
            if rule != 2:
                assert got[3] == a.accepted


def test_resident_series_is_resumable(m):
    """3 calls of 4 sweeps == 1 call of 12 sweeps == streaming: the sweep counter carries across launches"""
    a, _ = _run(m, [256, 256], 0, 0.5, 5, 4, nchains=2, env={"MCX_RESIDENT": "1"}, calls=3)
    b, _ = _run(m, [256, 256], 0, 0.5, 5, 12, nchains=2, env={"MCX_RESIDENT": "1"})
    c, _ = _run(m, [256, 256], 0, 0.5, 5, 12, nchains=2, env={"MCX_RESIDENT": "0"})
    assert _same(a, b) and _same(b, c)


def test_resident_more_chains_than_clusters_and_labels(m):
    """a batch larger than the number of co-resident clusters (clusters loop over chains), per-chain tables"""
    L, n, nsweeps = 256, 200, 3
    betas = np.linspace(0.2, 0.7, 8)
    outs = []
    for env in ({"MCX_RESIDENT": "0"}, {"MCX_RESIDENT": "1"}, {"MCX_RESIDENT": "1", "MCX_RESIDENT_CLUSTER": "8"}):
        os.environ.update(env)
        sys_ = m.Ising([L, L], nchains=n)
        sys_.set_rule(0, np.stack([m.build_table(0, 0, 2, b) for b in betas]))
        sys_.set_labels(np.arange(n) % 8)
        sys_.set_rng(99, 0)
        sys_.init_("random", rng=m.PhiloxRNG(99, 0))
        m.lib().mcx_sweep(sys_.h_lat, nsweeps)
        outs.append((sys_.spins.copy(), np.array(sys_.pair_sum()), np.array(sys_.accepted())))
        for k in KEYS:
            os.environ.pop(k, None)
    assert _same(outs[0], outs[1]) and _same(outs[0], outs[2])
    assert len(set(outs[0][2].tolist())) > 8          # chains differ


def test_default_policy_uses_resident_for_small_batches(m):
    """default: series of >= 2 sweeps over small batches take the resident path (one launch); single
    sweeps, big batches and lattices that need an 8-CTA cluster per chain in several waves do not"""
    _, l_series = _run(m, [512, 512], 0, BETA_C, 1, 10, nchains=4)
    assert l_series == 1
    _, l_tiny = _run(m, [64, 64], 0, BETA_C, 1, 10, nchains=4)
    assert l_tiny == 1
    _, l_single = _run(m, [512, 512], 0, BETA_C, 1, 1, nchains=4)
    assert l_single == 2
    _, l_big = _run(m, [1024, 1024], 0, BETA_C, 1, 3, nchains=40)
    assert l_big >= 6                                             # streaming launches (chain groups), not the resident kernel
    _, l_odd = _run(m, [48, 64], 0, BETA_C, 1, 3, nchains=2)       # Lx % 32 != 0: not a row-aligned shape
    assert l_odd == 6
