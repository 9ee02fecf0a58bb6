"""GPU parity tests of MCX_STORAGE_BIT (one bit per spin, north_star "int8/bit-packed spins"): the bit planes must give
the oracle's trajectories -- and therefore the int8 planes' -- bit for bit (tolerance 0), for every rule, in 2-D and 3-D,
whole launches, chain groups and row bands; plus the host bit buffers (an eighth of the PCIe bytes) on both storages."""
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


def _alg(m, rule, beta, seed, chain):
    rng = m.PhiloxRNG(seed, chain)
    return (m.Metropolis, m.Glauber, m.HeatBath)[rule](rng, beta=beta)


def _oracle_run(oracle, dims, rule, beta, J, h, seed, chain, nsweeps):
    s = oracle.System(oracle.ISING, dims, J=float(J), h=float(h), D=0.0)
    s.init_random(seed, chain)
    a = oracle.Alg(rule, beta)
    s.sweep_checkerboard(a, seed, chain, 0, nsweeps)
    return s, a


def _same_as_oracle(sys_, alg, s_or, a_or):
    assert np.array_equal(sys_.spins, s_or.spins)
    assert sys_.pair_sum() == s_or.pair_count()
    assert sys_.magnetization() == s_or.magnetization(full=True)
    assert sys_.energy() == sys_.energy(full=True) and sys_.magnetization() == sys_.magnetization(full=True)
    if hasattr(alg, "accepted"):
        assert alg.accepted == a_or.accepted
    assert alg.steps == a_or.steps


@pytest.mark.parametrize("dims", [[32, 4], [32, 32], [64, 6], [128, 64], [256, 256], [96, 20]])
@pytest.mark.parametrize("rule", [0, 1, 2])
def test_bits_2d_bit_exact_with_oracle(m, oracle, dims, rule):
    for beta, seed, J, h in ((BETA_C, 42, 1, 0), (0.25, 7, 1, 0.2), (0.9, 123456789012345, 2, 0)):
        nsweeps = 8
        s_or, a_or = _oracle_run(oracle, dims, rule, beta, J, h, seed, 3, nsweeps)
        for track in (True, False):
            sys_ = m.Ising(dims, J=J, h=h, storage="bit")
            sys_.set_tracking(track)
            alg = _alg(m, rule, beta, seed, 3)
            sys_.init_("random", rng=alg.rng)
            m.sweep_(sys_, alg, nsweeps - 3)
            m.sweep_(sys_, alg, 3)
            _same_as_oracle(sys_, alg, s_or, a_or)


@pytest.mark.parametrize("dims", [[32, 4, 4], [32, 8, 6], [64, 16, 8], [96, 10, 4]])
@pytest.mark.parametrize("rule", [0, 1, 2])
def test_bits_3d_bit_exact_with_oracle(m, oracle, dims, rule):
    for beta, seed in ((0.2216544, 42), (0.4, 9)):
        nsweeps = 6
        s_or, a_or = _oracle_run(oracle, dims, rule, beta, 1, 0, seed, 1, nsweeps)
        for track in (True, False):
            sys_ = m.Ising(dims, storage="bit")
            sys_.set_tracking(track)
            alg = _alg(m, rule, beta, seed, 1)
            sys_.init_("random", rng=alg.rng)
            m.sweep_(sys_, alg, nsweeps)
            _same_as_oracle(sys_, alg, s_or, a_or)


@pytest.mark.parametrize("dims,nchains,env", [
    ([1024, 1024], 1, {}),
    ([1024, 1024], 1, {"MCX_BANDS": "8"}),
    ([2048, 512], 1, {"MCX_BANDS": "4"}),
    ([512, 512], 6, {"MCX_GROUPS": "3"}),
    ([512, 512], 6, {"MCX_GROUPS": "0"}),
    ([128, 64, 32], 3, {}),
    ([128, 32, 32], 1, {"MCX_BANDS": "4"}),
])
def test_bits_equal_int8(m, dims, nchains, env):
    """the two storages side by side at sizes the oracle would take long for: spins, sums and counters identical,
    through whole launches, row bands and chain groups"""
    keys = ("MCX_BANDS", "MCX_GROUPS")
    try:
        for rule in (0, 2):
            out = []
            for storage in ("int8", "bit"):
                for k in keys:
                    os.environ.pop(k, None)
                os.environ.update(env)
                s = m.Ising(dims, nchains=nchains, storage=storage)
                alg = _alg(m, rule, BETA_C if len(dims) == 2 else 0.22, 11, 0)
                s.init_("random", rng=alg.rng)
                m.sweep_(s, alg, 5)
                out.append((s.spins.copy(), np.array(s.pair_sum()), np.array(s.magnetization()), np.array(s.accepted())))
            assert all(np.array_equal(a, b) for a, b in zip(*out)), (rule, env)
    finally:
        for k in keys:
            os.environ.pop(k, None)


def test_bits_init_modes_and_upload_download(m):
    rng = np.random.default_rng(5)
    for dims in ([64, 8], [32, 6, 4]):
        s = m.Ising(dims, nchains=2, storage="bit")
        N = int(np.prod(dims))
        assert np.all(s.spins == 1) and np.all(s.magnetization() == N)
        s.init_("down")
        assert np.all(s.spins == -1) and np.all(s.magnetization() == -N)
        host = (2 * rng.integers(0, 2, size=(2, N)) - 1).astype(np.int8)
        s.spins = host
        assert np.array_equal(s.spins, host)
        ref = m.Ising(dims, nchains=2)
        ref.spins = host
        assert np.array_equal(s.pair_sum(), ref.pair_sum()) and np.array_equal(s.magnetization(), ref.magnetization())


@pytest.mark.parametrize("storage", ["int8", "bit"])
def test_host_bit_buffers(m, storage):
    """mcx_lattice_upload_bits / _download_bits / _upload_bits_begin: numpy's little-endian packbits of (spins > 0)"""
    rng = np.random.default_rng(6)
    dims, nch = [64, 16], 3
    N = dims[0] * dims[1]
    host = (2 * rng.integers(0, 2, size=(nch, N)) - 1).astype(np.int8)
    bits = np.packbits(host > 0, axis=1, bitorder="little")
    s = m.Ising(dims, nchains=nch, storage=storage)
    s.spin_bits = bits
    assert np.array_equal(s.spins, host)
    assert np.array_equal(s.spin_bits, bits)
    ref = m.Ising(dims, nchains=nch)
    ref.spins = host
    assert np.array_equal(s.pair_sum(), ref.pair_sum())
    # split upload of a bit buffer
    host2 = -host
    bits2 = np.ascontiguousarray(np.packbits(host2 > 0, axis=1, bitorder="little"))
    s.upload_bits_begin(bits2.ctypes.data)
    s.upload_commit()
    assert np.array_equal(s.spins, host2)
    with pytest.raises(m.McxError):
        m.BlumeCapel([32, 8]).spin_bits = np.zeros(32, dtype=np.uint8)


def test_bits_rejects_unsupported_shapes(m):
    with pytest.raises(m.McxError):
        m.Ising([48, 8], storage="bit")            # Lx % 32 != 0
    with pytest.raises(m.McxError):
        m.Ising([64], storage="bit")               # 1-D
    with pytest.raises(ValueError):
        m.Ising([64, 8], storage="nibble")


def test_bits_parallel_tempering_matches_int8(m):
    """replica exchange on bit planes: same labels, same energies, same acceptance as on int8 planes"""
    outs = []
    for storage in ("int8", "bit"):
        betas = m.set_betas(8, 0.3, 0.6)
        pt = m.ParallelTempering(betas, seed=5, backend=m.GPUBackend())
        reps = m.Ising([64, 64], nchains=8, storage=storage)
        pt.attach(reps)
        reps.init_("random", rng=m.PhiloxRNG(5, 0))
        for _ in range(6):
            m.sweep_(reps, pt, 3)
            m.update_(pt)
        outs.append((list(pt.index()), reps.spins.copy(), np.array(reps.pair_sum()), np.array(pt.accepted)))
    assert outs[0][0] == outs[1][0]
    assert all(np.array_equal(a, b) for a, b in zip(outs[0][1:], outs[1][1:]))


def test_bits_checkpoint_round_trip(m):
    import pickle
    s = m.Ising([64, 32], storage="bit")
    alg = _alg(m, 0, BETA_C, 3, 0)
    s.init_("random", rng=alg.rng)
    m.sweep_(s, alg, 4)
    blob = pickle.dumps(s)
    m.sweep_(s, alg, 4)
    r = pickle.loads(blob)
    assert r.storage == "bit"
    m.sweep_(r, alg, 4)
    assert np.array_equal(r.spins, s.spins) and r.pair_sum() == s.pair_sum()
