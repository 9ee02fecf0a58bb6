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


def test_deterministic_restart(m, tmp_path):
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
    m.finalize_(ck)


def test_sweep_series_matches_per_sweep_reads(m):
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
