"""Measurement helpers on the path's callers: integrated autocorrelation time
(src/measurements/autocorrelations.jl:28-110), the tau_int-driven retuning of post-exchange sweep
counts (src/algorithms/parallel_tempering.jl:56-143), and the on-device measurement series that
replaces `measure!(measurements, sys, i)` once per sweep (src/measurements/measurements.jl:192-200)."""
import math
import statistics

import numpy as np

from ._lib import check, lib


def integrated_autocorrelation_time(samples, max_lag=None, c=5.0):
    """autocorrelations.jl:28-65: tau_int = 1/2 + sum_{t>=1} C(t), self-consistent window lag > c*tau,
    truncated at the first non-positive C(t); 0.5 for zero-variance input."""
    x = np.asarray(samples, dtype=np.float64)
    n = x.size
    if n < 2:
        raise ValueError("integrated_autocorrelation_time requires at least 2 samples")
    if not c > 0:
        raise ValueError("c must be positive")
    lag_cap_max = n // 2
    lag_cap = lag_cap_max if max_lag is None else int(max_lag)
    if not 1 <= lag_cap <= lag_cap_max:
        raise ValueError("max_lag must satisfy 1 <= max_lag <= floor(length(samples)/2)")
    centered = x - x.sum() / n
    C0 = float(np.dot(centered, centered)) / n
    if not C0 > 0:
        return 0.5
    tau = 0.5
    for lag in range(1, lag_cap + 1):
        C = float(np.dot(centered[:n - lag], centered[lag:])) / (n - lag) / C0
        if C <= 0:
            break
        tau_next = tau + C
        if lag > c * tau_next:
            break
        tau = tau_next
    return max(0.5, tau)


def integrated_autocorrelation_times(traces, min_points=2, max_lag=None, c=5.0):
    """autocorrelations.jl:73-101"""
    if min_points < 2:
        raise ValueError("min_points must be >= 2")
    taus = [float("nan")] * len(traces)
    for i, trace in enumerate(traces):
        n = len(trace)
        if n < min_points:
            continue
        local = None if max_lag is None else min(int(max_lag), n // 2)
        if local is not None and local < 1:
            continue
        taus[i] = integrated_autocorrelation_time(trace, max_lag=local, c=c)
    return taus


tau_int = integrated_autocorrelation_time


def _group_samples(samples, n):
    """parallel_tempering.jl:62-73"""
    if len(samples) and isinstance(samples[0], tuple):
        grouped = [[] for _ in range(n)]
        for idx, e in samples:
            if not 1 <= idx <= n:
                raise ValueError("sample index %d out of bounds for %d ladders" % (idx, n))
            grouped[idx - 1].append(float(e))
        return grouped
    return samples


def _retune_exchange_sweeps_(sweeps_after_exchange, taus, base_sweeps, min_sweeps, max_sweeps):
    """parallel_tempering.jl:124-137"""
    finite = [t for t in taus if math.isfinite(t)]
    tau_ref = statistics.median(finite) if finite else 1.0
    for i in range(len(sweeps_after_exchange)):
        scale = taus[i] / tau_ref if math.isfinite(taus[i]) else 1.0
        x = base_sweeps * scale
        target = int(np.rint(x))          # round(Int, x): ties to even, like Julia
        sweeps_after_exchange[i] = min(max(target, min_sweeps), max_sweeps)
    return sweeps_after_exchange


def optimize_exchange_interval_(pt, local_samples, sweeps_after_exchange, base_sweeps, min_sweeps=1,
                                max_sweeps=2 ** 62, min_points=400, max_lag=200):
    """parallel_tempering.jl:83-122 (single-process branch; with several ranks the caller gathers the
    samples first, as the MPI branch does)."""
    n = pt.size
    if len(sweeps_after_exchange) != n:
        raise ValueError("sweeps_after_exchange must have length size(pt)")
    grouped = _group_samples(local_samples, n)
    taus = integrated_autocorrelation_times(grouped, min_points=min_points, max_lag=max_lag)
    _retune_exchange_sweeps_(sweeps_after_exchange, taus, base_sweeps, min_sweeps, max_sweeps)
    idx = pt.index()
    return sweeps_after_exchange[int(idx[0]) - 1] if pt.size else None


def sweep_series_(sys, alg, nmeasure, interval=1):
    """nmeasure x (interval sweeps + snapshot) entirely on the device; returns a dict of arrays
    [nmeasure, nchains]: energy, magnetization, pair_sum, spin2_sum, accepted (cumulative)."""
    sys._bind_alg(alg)
    before = int(sys._sums()[3].sum())
    out = np.empty((int(nmeasure), sys.nchains, 4), dtype=np.int64)
    check(lib().mcx_sweep_series(sys.h_lat, int(nmeasure), int(interval), out.ctypes.data))
    alg.steps += int(nmeasure) * int(interval) * sys.N * sys.nchains     # summed over the chains, like alg.accepted
    pair, spin, spin2, acc = (out[:, :, k] for k in range(4))
    if hasattr(alg, "accepted") and nmeasure:
        alg.accepted += int(acc[-1].sum()) - before
    return {"energy": sys._energy_from(pair, spin, spin2), "magnetization": spin, "pair_sum": pair,
            "spin2_sum": spin2, "accepted": acc}


_TAU_OBS = {"energy": 0, "magnetization": 1, "abs_magnetization": 2}


def series_tau_int_(sys, nmeasure, observable="energy", max_lag=None, c=5.0):
    """integrated_autocorrelation_time (autocorrelations.jl:28-65) of the series `sweep_series_` just left on the device,
    per chain, computed there (mcx_series_tau_int): float64[nchains].  Agrees with `integrated_autocorrelation_time` on the
    copied-back series to rounding."""
    if observable not in _TAU_OBS:
        raise ValueError("observable must be one of %s" % sorted(_TAU_OBS))
    out = np.empty(sys.nchains, dtype=np.float64)
    check(lib().mcx_series_tau_int(sys.h_lat, int(nmeasure), _TAU_OBS[observable], 0 if max_lag is None else int(max_lag), float(c),
                                   out.ctypes.data))
    return out
