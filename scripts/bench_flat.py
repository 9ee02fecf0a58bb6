#!/usr/bin/env python
"""Auxiliary timings for the flat-histogram path (BASELINE.json configs[3], configs[4]) and the
canonical generic kernels.  Prints one JSON object per line; not the headline bench."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcx_b200 as m  # noqa: E402
from mcx_b200._lib import check, lib  # noqa: E402


def timed(fn, sync):
    sync()
    t0 = time.perf_counter()
    fn()
    sync()
    return time.perf_counter() - t0


def c4(nsweeps=2, L=512, nch=1024):
    N = L * L
    sys_ = m.BlumeCapel([L, L], nchains=nch)
    h = C.c_void_p()
    sys_.set_rng(42)
    check(lib().mcx_flat_create(sys_.h_lat, m._lib.FLAT_MUCA, m._lib.OBS_SPIN2_WITH_PAIR_BOLTZMANN, 0, 1, N + 1, 1 / 0.9, 0, C.byref(h)))
    check(lib().mcx_flat_sweep(h, 1))
    dt = timed(lambda: check(lib().mcx_flat_sweep(h, nsweeps)), sys_.sync)
    check(lib().mcx_flat_destroy(h))
    return {"config": "C4 Blume-Capel L=%d muca(sum s^2), %d chains" % (L, nch), "attempts_per_ns": nsweeps * nch * N / dt / 1e9,
            "sweeps_per_s": nsweeps / dt, "seconds": dt}


def c5(nsweeps=1, L=256, nch=32):
    N = L ** 3
    sys_ = m.Ising([L, L, L], nchains=nch)
    sys_.init_("random", rng=m.PhiloxRNG(42, 0))
    sys_.set_rng(42)
    h = C.c_void_p()
    check(lib().mcx_flat_create(sys_.h_lat, m._lib.FLAT_WANG_LANDAU, m._lib.OBS_ENERGY, -(1 << 20), 4, (1 << 19) + 1, 0.0, 1, C.byref(h)))
    check(lib().mcx_flat_set_logf(h, 1.0))
    dt = timed(lambda: check(lib().mcx_flat_sweep(h, nsweeps)), sys_.sync)
    check(lib().mcx_flat_destroy(h))
    return {"config": "C5 3-D Ising L=%d Wang-Landau window, %d walkers" % (L, nch), "attempts_per_ns": nsweeps * nch * N / dt / 1e9,
            "seconds": dt}


def wl2d(nsweeps=20, L=64, nch=2048, init="up"):
    """Wang-Landau walkers with private tables on small lattices; init "up" starts at the ground state, where
    few proposals are accepted (the regime in which deciding many attempts at once pays most)"""
    N = L * L
    sys_ = m.Ising([L, L], nchains=nch)
    if init == "random":
        sys_.init_("random", rng=m.PhiloxRNG(42, 0))
    sys_.set_rng(42)
    h = C.c_void_p()
    check(lib().mcx_flat_create(sys_.h_lat, m._lib.FLAT_WANG_LANDAU, m._lib.OBS_ENERGY, -2 * N, 4, N + 1, 0.0, 0, C.byref(h)))
    check(lib().mcx_flat_set_logf(h, 1.0))
    check(lib().mcx_flat_sweep(h, 5))
    before = np.atleast_1d(sys_.accepted()).sum()
    dt = timed(lambda: check(lib().mcx_flat_sweep(h, nsweeps)), sys_.sync)
    acc = (np.atleast_1d(sys_.accepted()).sum() - before) / (nsweeps * nch * N)
    check(lib().mcx_flat_destroy(h))
    return {"config": "2-D Ising L=%d Wang-Landau(E), %d walkers, init %s" % (L, nch, init),
            "attempts_per_ns": nsweeps * nch * N / dt / 1e9, "acceptance": float(acc), "seconds": dt}


def wl_low(dims, nch, frac=0.2, nsweeps=20, spec=None):
    """Wang-Landau walkers confined to the lowest `frac` of the energy range (out_of_range_policy 1), timed at
    logf = 2^-8 after four coarser stages: the low-acceptance regime of the low-energy windows"""
    N, d = int(np.prod(dims)), len(dims)
    sys_ = m.Ising(dims, nchains=nch)              # all up: E = -d N, inside the window
    sys_.set_rng(42)
    nbins = int(frac * d * N / 2) + 1
    h = C.c_void_p()
    check(lib().mcx_flat_create(sys_.h_lat, m._lib.FLAT_WANG_LANDAU, m._lib.OBS_ENERGY, -d * N, 4, nbins, 0.0, 1, C.byref(h)))
    os.environ.pop("MCX_WL_SPEC", None)
    for logf, n in ((1.0, 20), (0.25, 20), (2.0 ** -4, 20), (2.0 ** -8, 50)):
        check(lib().mcx_flat_set_logf(h, logf))
        check(lib().mcx_flat_sweep(h, n))
    if spec is not None:
        os.environ["MCX_WL_SPEC"] = spec
    before = np.atleast_1d(sys_.accepted()).sum()
    dt = timed(lambda: check(lib().mcx_flat_sweep(h, nsweeps)), sys_.sync)
    acc = (np.atleast_1d(sys_.accepted()).sum() - before) / (nsweeps * nch * N)
    check(lib().mcx_flat_destroy(h))
    return {"config": "Ising %s Wang-Landau(E), %d walkers, window = lowest %.0f %% of the range, logf 2^-8" % (dims, nch, 100 * frac),
            "attempts_per_ns": nsweeps * nch * N / dt / 1e9, "acceptance": float(acc), "seconds": dt}


def muca2d(nsweeps=200, L=64, nch=4096):
    N = L * L
    sys_ = m.Ising([L, L], nchains=nch)
    sys_.init_("random", rng=m.PhiloxRNG(42, 0))
    sys_.set_rng(42)
    h = C.c_void_p()
    check(lib().mcx_flat_create(sys_.h_lat, m._lib.FLAT_MUCA, m._lib.OBS_ENERGY, -2 * N, 4, N + 1, 0.0, 0, C.byref(h)))
    check(lib().mcx_flat_sweep(h, 5))
    dt = timed(lambda: check(lib().mcx_flat_sweep(h, nsweeps)), sys_.sync)
    check(lib().mcx_flat_destroy(h))
    return {"config": "2-D Ising L=%d muca(E), %d chains" % (L, nch), "attempts_per_ns": nsweeps * nch * N / dt / 1e9, "seconds": dt}


def canonical(model, dims, nch, nsweeps, rule=0):
    sys_ = (m.Ising if model == 0 else m.BlumeCapel)(dims, nchains=nch)
    rng = m.PhiloxRNG(42, 0)
    alg = m.Metropolis(rng, beta=0.3)
    sys_._bind_alg(alg)
    sys_.init_("random", rng=rng)
    check(lib().mcx_sweep(sys_.h_lat, 2))
    dt = timed(lambda: check(lib().mcx_sweep(sys_.h_lat, nsweeps)), sys_.sync)
    n = int(np.prod(dims)) * nch
    return {"config": "canonical %s %s x %d chains" % ("Ising" if model == 0 else "Blume-Capel", dims, nch),
            "attempts_per_ns": nsweeps * n / dt / 1e9, "seconds": dt}


if __name__ == "__main__":
    which = sys.argv[1:] or ["c4", "c5", "muca2d", "gen"]
    if "c4" in which:
        print(json.dumps(c4()))
    if "c5" in which:
        print(json.dumps(c5()))
    if "wlscan" in which:      # Wang-Landau attempts decided at once: serial loop / 8 / 32 / adaptive
        for spec in ("0", "8", "32", None):
            if spec is None:
                os.environ.pop("MCX_WL_SPEC", None)
            else:
                os.environ["MCX_WL_SPEC"] = spec
            for r in (wl_low([64, 64], 2048, 0.2, spec=spec), wl_low([64, 64], 2048, 0.05, spec=spec),
                      wl_low([32, 32, 32], 512, 0.1, spec=spec), wl2d(init="random")):
                r["MCX_WL_SPEC"] = spec or "adaptive"
                print(json.dumps(r), flush=True)
    if "muca2d" in which:
        print(json.dumps(muca2d()))
    if "gen" in which:
        print(json.dumps(canonical(1, [8192, 8192], 1, 10)))
        print(json.dumps(canonical(0, [512, 512, 512], 1, 10)))
        print(json.dumps(canonical(1, [256, 256, 256], 1, 10)))
        print(json.dumps(canonical(0, [8176, 8176], 1, 10)))
        print(json.dumps(canonical(0, [1000, 1000], 64, 10)))
        print(json.dumps(canonical(0, [64, 64], 16384, 20)))
        print(json.dumps(canonical(0, [128, 128], 4096, 20)))
