"""k_queue.cu: a whole series of sweeps in one launch, the work items of all half-sweeps taken from one ticket
counter and ordered by per-item progress words (MCX_QUEUE=1).  Must reproduce the launch-per-half-sweep path
(itself held to the oracle in tests/test_gpu_parity.py) bit for bit: spins, sums, accepted counts."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m():
    import mcx_b200
    mcx_b200.lib()
    return mcx_b200


def _run(m, dims, nch, rule, track, nsweeps, seed, parts):
    """`parts` successive mcx_sweep calls of nsweeps each on a batch with one table (label) per chain"""
    betas = np.linspace(0.25, 0.6, nch)
    sys_ = m.Ising(dims, nchains=nch)
    sys_.set_rule(rule, np.stack([m.build_table(0, rule, 2, float(b)) for b in betas]))
    sys_.set_labels(np.arange(nch, dtype=np.int32)[::-1].copy())
    sys_.set_tracking(track)
    sys_.set_rng(seed, 0)
    sys_.init_("random", rng=m.PhiloxRNG(seed, 3))
    before = sys_.ctx.launch_count()
    for _ in range(parts):
        m._lib.check(m.lib().mcx_sweep(sys_.h_lat, nsweeps))
    launches = sys_.ctx.launch_count() - before
    out = (sys_.spins.copy(), np.array(sys_.pair_sum()), np.array(sys_.magnetization()), np.array(sys_.accepted()))
    assert np.array_equal(np.atleast_1d(sys_.energy()), np.atleast_1d(sys_.energy(full=True)))
    return out, launches


@pytest.mark.parametrize("dims,nch", [([64, 64], 5), ([256, 256], 3), ([1024, 64], 2), ([32, 1024], 2), ([512, 512], 1),
                                       ([4096, 48], 1), ([1024, 1024], 4)])
@pytest.mark.parametrize("rule,track", [(0, True), (0, False), (1, True), (2, False)])
def test_queue_series_equals_half_sweep_launches(m, dims, nch, rule, track, monkeypatch):
    nsweeps, parts = (7, 2) if dims != [1024, 1024] else (5, 1)
    monkeypatch.setenv("MCX_QUEUE", "0")
    monkeypatch.setenv("MCX_RESIDENT", "0")
    ref, _ = _run(m, dims, nch, rule, track, nsweeps, 99, parts)
    monkeypatch.setenv("MCX_QUEUE", "1")
    got, launches = _run(m, dims, nch, rule, track, nsweeps, 99, parts)
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)
    assert launches == parts                      # one launch per series (tracking off: + nothing; sums recomputed lazily)


@pytest.mark.parametrize("dims,rows", [([2048, 2048], 2), ([4096, 4096], 4), ([1024, 1024], 2), ([4096, 1024], 2)])
def test_single_mid_size_lattice_takes_the_queue_by_default(m, dims, rows, monkeypatch):
    """one lattice of 1024 ... 8192 rows: the default policy runs a series of >= 4 sweeps as ONE launch with short strips
    (2 / 4 rows); same trajectory as one launch per half-sweep"""
    monkeypatch.setenv("MCX_RESIDENT", "0")
    for rule, track in ((0, False), (2, True)):
        monkeypatch.setenv("MCX_QUEUE", "0")
        monkeypatch.setenv("MCX_BANDS", "0")                  # one launch per half-sweep
        ref, l_ref = _run(m, dims, 1, rule, track, 6, 5, 1)
        monkeypatch.delenv("MCX_QUEUE", raising=False)
        monkeypatch.delenv("MCX_BANDS", raising=False)
        got, launches = _run(m, dims, 1, rule, track, 6, 5, 1)
        assert launches == 1 and l_ref >= 12, (launches, l_ref)
        for a, b in zip(ref, got):
            assert np.array_equal(a, b)


def test_queue_series_matches_oracle(m, oracle, monkeypatch):
    monkeypatch.setenv("MCX_QUEUE", "1")
    L, nsweeps, seed = 64, 9, 2024
    sys_ = m.Ising([L, L])
    alg = m.Glauber(m.PhiloxRNG(seed, 1), beta=0.44)
    sys_.init_("random", rng=alg.rng)
    m.sweep_(sys_, alg, nsweeps)
    s = oracle.System(oracle.ISING, [L, L])
    s.init_random(seed, 1)
    a = oracle.Alg(oracle.GLAUBER, 0.44)
    s.sweep_checkerboard(a, seed, 1, 0, nsweeps)
    assert np.array_equal(sys_.spins, s.spins) and alg.accepted == a.accepted
    assert sys_.energy() == s.energy(full=True)


def test_default_policy_takes_the_queue_for_a_pt_rank_share(m, monkeypatch):
    """unset MCX_QUEUE: a batch whose half-sweep is about one work item per resident CTA (32 replicas of 1024 x 1024,
    what a parallel-tempering rank holds at 8 GPUs) runs as one launch per series, bit-identical to MCX_QUEUE=0"""
    monkeypatch.setenv("MCX_QUEUE", "0")
    ref, ref_launches = _run(m, [1024, 1024], 32, 0, False, 6, 7, 1)
    monkeypatch.delenv("MCX_QUEUE", raising=False)
    got, launches = _run(m, [1024, 1024], 32, 0, False, 6, 7, 1)
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)
    assert launches == 1 and ref_launches >= 12
    # big batches keep the chain-group launches
    _, launches = _run(m, [1024, 1024], 96, 0, False, 4, 7, 1)
    assert launches >= 8


@pytest.mark.parametrize("rows", ["16", "8", "2"])
def test_queue_strip_heights(m, rows, monkeypatch):
    """every strip height the launcher can pick (MCX_QUEUE_ROWS pins it)"""
    monkeypatch.setenv("MCX_QUEUE", "0")
    ref, _ = _run(m, [1024, 256], 3, 1, True, 5, 11, 2)
    monkeypatch.setenv("MCX_QUEUE", "1")
    monkeypatch.setenv("MCX_QUEUE_ROWS", rows)
    got, launches = _run(m, [1024, 256], 3, 1, True, 5, 11, 2)
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)
    assert launches == 2
