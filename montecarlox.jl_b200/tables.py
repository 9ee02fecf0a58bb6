"""Host-built integer rule tables (include/mcx_b200.h "canonical update rule").

For a draw u = m * 2^-32 and a Float64 p:  u < p  <=>  m < ceil(p * 2^32), and p * 2^32 is exact.
Every entry is obtained by evaluating the reference's own float expression for one local
configuration -- the generalisation of the reference's TableMetropolis example
(docs/src/examples/spin_systems/importance_Ising2D.jl:74-92) -- so the device never calls exp."""
import math

import numpy as np

from ._lib import BLUME_CAPEL, GLAUBER, HEATBATH, ISING, METROPOLIS
from .algorithms import logistic

TWO32 = 4294967296


def _thr_from_p(p):
    if not (p > 0):
        return 0
    x = math.ceil(p * 4294967296.0)
    return TWO32 if x >= TWO32 else int(x)


def _thr_accept(rule, log_ratio):
    if rule == GLAUBER:                                   # metropolis.jl:124
        return _thr_from_p(logistic(log_ratio))
    if log_ratio > 0:                                     # importance_sampling.jl:82
        return TWO32
    return _thr_from_p(math.exp(log_ratio))


def _thr_scaled(z, w):
    """first m in [0, 2^32] with not (m*2^-32*z < w): threshold of blume_capel.jl:75-76"""
    lo, hi = 0, TWO32
    while lo < hi:
        mid = (lo + hi) >> 1
        if (mid / 4294967296.0) * z < w:
            lo = mid + 1
        else:
            hi = mid
    return lo


def table_len(model, rule, ndim):
    nn = 2 * ndim
    if model == ISING:
        return 2 * (nn + 1)
    return 2 * (2 * nn + 1) if rule == HEATBATH else 6 * (2 * nn + 1)


def propose_state(u, s_old):
    """_propose_state (blume_capel.jl:21-30)"""
    if s_old == -1:
        return 0 if u else 1
    if s_old == 0:
        return -1 if u else 1
    return -1 if u else 0


def build_table(model, rule, ndim, beta, J=1, h=0, D=0):
    nn = 2 * ndim
    T = np.zeros(table_len(model, rule, ndim), dtype=np.uint64)
    if model == ISING:
        for sb in (0, 1):
            for nup in range(nn + 1):
                s = 1 if sb else -1
                lpi = s * (2 * nup - nn)                  # local_pair_interactions abstractions.jl:41-48
                dpair = -2 * J * lpi                      # flip_changes ising.jl:187-192
                dspin = -2 * s
                dE = -dpair - h * dspin                   # delta_energy ising.jl:198
                if rule == HEATBATH:                      # ising.jl:49-50
                    t = _thr_from_p(logistic(beta * float(s) * dE))
                else:
                    t = _thr_accept(rule, -beta * dE)     # boltzmann.jl:28
                T[sb * (nn + 1) + nup] = t
    elif rule == HEATBATH:
        for nsum in range(-nn, nn + 1):                   # blume_capel.jl:61-76
            coupling = float(J) * nsum
            h_i = float(h)
            e1 = -(-1) * coupling - h_i * (-1) + D
            e2 = 0.0
            e3 = -(1) * coupling - h_i * (1) + D
            w1, w2, w3 = math.exp(-beta * e1), math.exp(-beta * e2), math.exp(-beta * e3)
            z = w1 + w2 + w3
            T[nsum + nn] = _thr_scaled(z, w1)
            T[(2 * nn + 1) + nsum + nn] = _thr_scaled(z, w1 + w2)
    else:
        for so in range(3):
            for b in (0, 1):
                for nsum in range(-nn, nn + 1):           # blume_capel.jl:52-59, :235-248
                    s_old = so - 1
                    s_new = propose_state(b, s_old)
                    dspin = s_new - s_old
                    dspin2 = s_new * s_new - s_old * s_old
                    dpair = dspin * (float(J) * nsum)
                    dE = -dpair - (h * dspin if h != 0 else 0.0) + D * dspin2
                    T[(so * 2 + b) * (2 * nn + 1) + nsum + nn] = _thr_accept(rule, -beta * dE)
    return T


def rule_of(alg):
    return {"metropolis": METROPOLIS, "glauber": GLAUBER, "heatbath": HEATBATH}[alg.kind]


def beta_of(alg):
    if alg.kind == "heatbath":
        return alg.beta
    ens = alg.ensemble
    if not hasattr(ens, "beta"):
        raise ValueError("checkerboard sweeps need a BoltzmannEnsemble (local acceptance); got %s" % type(ens).__name__)
    return ens.beta
