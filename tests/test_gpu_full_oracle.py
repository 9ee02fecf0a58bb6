"""One oracle sweep at the benchmarked size and path: the default banded L = 16384 launch sequence (and the bit planes)
against the CPU oracle directly, not through smaller lattices (parity item 1)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
BETA_C = 0.440686793509772


@pytest.fixture(scope="module")
def m():
    import mcx_b200
    mcx_b200.lib()
    return mcx_b200


def test_headline_path_against_the_oracle_at_16384(m, oracle):
    L, seed, nsweeps = 16384, 42, 1
    ref = oracle.System(oracle.ISING, [L, L])
    ref.init_random(seed, 0)
    ralg = oracle.Alg(oracle.METROPOLIS, BETA_C)
    ref.sweep_checkerboard(ralg, seed, 0, 0, nsweeps)
    want = ref.spins.copy()
    pair, mag = ref.pair_count(), ref.magnetization(full=True)
    del ref
    for storage, track in (("int8", False), ("int8", True), ("bit", False)):
        sys_ = m.Ising([L, L], storage=storage)
        sys_.set_tracking(track)
        alg = m.Metropolis(m.PhiloxRNG(seed, 0), beta=BETA_C)
        sys_.init_("random", rng=alg.rng)
        l0 = sys_.ctx.launch_count()
        m.sweep_(sys_, alg, nsweeps)
        assert sys_.ctx.launch_count() - l0 >= 16 * nsweeps, "the banded path (8 launches per half-sweep) was not taken"
        assert np.array_equal(sys_.spins, want), (storage, track)
        assert sys_.pair_sum() == pair and sys_.magnetization() == mag and alg.accepted == ralg.accepted
        del sys_
