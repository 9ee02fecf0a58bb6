"""Seeded cases behind tests/golden/trajectories.json (RNG layout v1).

`oracle_result(case)` runs a case on the oracle, `device_result(case, m)` on the GPU through the host mirror;
both return the same small dict (sha256 of the int8 spins in the reference's site order, the integer sums, the
accepted count, sha256 of the Float64 tables where there are any).  The committed file was written by
tests/golden/gen_trajectories.py from the oracle; tests/test_oracle_known_answers.py checks the oracle still
reproduces it (no GPU needed) and tests/test_gpu_parity.py requires the same of the device, without touching
the oracle at run time."""
import hashlib

import numpy as np

BETA_C = 0.440686793509772


def cases():
    out = []
    for rule in (0, 1, 2):
        for dims in ([8, 8], [64, 64], [256, 128], [96, 40]):
            out.append(dict(kind="canonical", model=0, dims=dims, rule=rule, beta=BETA_C, J=1, h=0, D=0,
                            seed=42, chain=rule, nsweeps=12))
        for dims in ([32, 8, 4], [6, 4, 8]):
            out.append(dict(kind="canonical", model=0, dims=dims, rule=rule, beta=0.2216544, J=1, h=0, D=0,
                            seed=43, chain=7, nsweeps=8))
        for dims in ([64, 32], [10, 6], [32, 4, 4]):
            out.append(dict(kind="canonical", model=1, dims=dims, rule=rule, beta=0.9, J=1, h=0, D=0.5,
                            seed=44, chain=1, nsweeps=8))
    out.append(dict(kind="canonical", model=0, dims=[64, 32], rule=0, beta=0.35, J=1, h=0.25, D=0, seed=45, chain=0,
                    nsweeps=10))
    out.append(dict(kind="canonical", model=0, dims=[64, 64], rule=1, beta=-0.3, J=1, h=0, D=0, seed=46, chain=3,
                    nsweeps=10))                                     # beta < 0: the drive of the upper windows
    out.append(dict(kind="muca", dims=[8, 8], seed=1000, chain=0, nsweeps=50))
    out.append(dict(kind="wl", dims=[8, 8], seed=77, chain=5, nsweeps=40, logf=1.0))
    out.append(dict(kind="wl", dims=[4, 4, 4], seed=78, chain=2, nsweeps=30, logf=0.25))
    return out


def name_of(c):
    parts = [c["kind"], "x".join(str(d) for d in c["dims"])]
    for k in ("model", "rule", "beta", "h", "D", "seed", "chain", "nsweeps"):
        if k in c:
            parts.append("%s=%s" % (k, c[k]))
    return " ".join(parts)


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _bins(dims):
    n, d = int(np.prod(dims)), len(dims)
    return range(-d * n, d * n + 1, 4)


def _muca_lw0(nbins):
    return np.linspace(0.0, 1.5, nbins) ** 2


def oracle_result(c):
    from oracle import oracle
    if c["kind"] == "canonical":
        s = oracle.System(c["model"], c["dims"], J=float(c["J"]), h=float(c["h"]), D=float(c["D"]))
        s.init_random(c["seed"], c["chain"])
        init_sha = _sha(s.spins)
        a = oracle.Alg(c["rule"], c["beta"])
        s.sweep_checkerboard(a, c["seed"], c["chain"], 0, c["nsweeps"])
        return dict(init=init_sha, spins=_sha(s.spins), pair=s.pair_count(), mag=s.magnetization(full=True),
                    spin2=s.spin2_sum() if c["model"] == 1 else None,
                    accepted=int(a.accepted) if c["rule"] != 2 else None, energy=float(s.energy(full=True)))
    bins = _bins(c["dims"])
    s = oracle.System(oracle.ISING, c["dims"])
    s.init_random(c["seed"], c["chain"])
    a = oracle.Alg(0, 0.0)
    f = oracle.Flat(bins[0], 4, len(bins), logf=c.get("logf", 1.0))
    if c["kind"] == "muca":
        f.logweight[:] = _muca_lw0(len(bins))
    assert s.flat_sweep(a, f, 0 if c["kind"] == "muca" else 1, 0, 0.0, c["seed"], c["chain"], 0, c["nsweeps"]) == 0
    return dict(spins=_sha(s.spins), accepted=int(a.accepted), energy=float(s.energy(full=True)),
                table=_sha(f.histogram if c["kind"] == "muca" else f.logweight))


def device_result(c, m):
    if c["kind"] == "canonical":
        sys_ = (m.Ising(c["dims"], J=c["J"], h=c["h"]) if c["model"] == 0
                else m.BlumeCapel(c["dims"], J=c["J"], D=c["D"], h=c["h"]))
        rng = m.PhiloxRNG(c["seed"], c["chain"])
        alg = [m.Metropolis, m.Glauber, m.HeatBath][c["rule"]](rng, beta=c["beta"])
        sys_.init_("random", rng=rng)
        init_sha = _sha(sys_.spins)
        m.sweep_(sys_, alg, c["nsweeps"])
        return dict(init=init_sha, spins=_sha(sys_.spins), pair=int(sys_.pair_sum()), mag=int(sys_.magnetization()),
                    spin2=int(sys_.spin2_sum()) if c["model"] == 1 else None,
                    accepted=int(alg.accepted) if c["rule"] != 2 else None, energy=float(sys_.energy()))
    bins = _bins(c["dims"])
    sys_ = m.Ising(c["dims"])
    rng = m.PhiloxRNG(c["seed"], c["chain"])
    sys_.init_("random", rng=rng)
    if c["kind"] == "muca":
        alg = m.Multicanonical(rng, bins)
        alg.ensemble.logweight_table.values[:] = _muca_lw0(len(bins))
    else:
        alg = m.WangLandau(rng, bins, logf=c["logf"])
    m.sweep_(sys_, alg, c["nsweeps"])
    table = alg.ensemble.histogram.values if c["kind"] == "muca" else alg.ensemble.logweight_table.values
    return dict(spins=_sha(sys_.spins), accepted=int(alg.accepted), energy=float(sys_.energy()), table=_sha(table))
