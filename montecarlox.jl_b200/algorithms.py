"""Acceptance algorithms (src/algorithms/{importance_sampling,metropolis,heat_bath,multicanonical,
wang_landau}.jl).  Host-side mirror: the same fields (`rng`, `ensemble`, `steps`, `accepted`) and the
same scalar `accept_` semantics; the lattice sweeps run them on the device through tables."""
import math

from .ensembles import (BoltzmannEnsemble, MulticanonicalEnsemble, WangLandauEnsemble, _as_ensemble)
from .rng import PhiloxRNG


def logistic(x):
    """Numerically stable sigmoid (src/infrastructure/utils.jl:82-88)."""
    if x >= 0:
        return 1.0 / (1.0 + math.exp(-x))
    ex = math.exp(x)
    return ex / (1.0 + ex)


class ImportanceSampling:
    """importance_sampling.jl:23-30"""
    kind = "importance"

    def __init__(self, rng, ensemble=None, beta=None):
        if ensemble is None:
            ensemble = BoltzmannEnsemble(beta=beta)
        self.rng = rng
        self.ensemble = _as_ensemble(ensemble)
        self.steps = 0
        self.accepted = 0

    # accept!(alg, x_new, x_old) (importance_sampling.jl:69-78)
    def accept_(self, x_new, x_old=None):
        if x_old is None:
            return self._accept_delta(x_new)
        ens = self.ensemble
        if isinstance(ens, WangLandauEnsemble):        # algorithms/wang_landau.jl:29-37
            lw = ens.logweight_table
            log_ratio = lw(x_new) - lw(x_old)
            accepted = self._accept(log_ratio)
            x_vis = x_new if accepted else x_old
            lw[x_vis] = lw[x_vis] - ens.logf
            return accepted
        log_ratio = ens.logweight(x_new) - ens.logweight(x_old)
        accepted = self._accept(log_ratio)
        if ens.should_record_visit:
            ens.record_visit_(x_new if accepted else x_old)
        return accepted

    def _accept_delta(self, delta_state):
        raise TypeError("accept!(alg, delta) is defined for AbstractMetropolis only")

    # _accept! (importance_sampling.jl:80-85)
    def _accept(self, log_ratio):
        self.steps += 1
        accepted = (log_ratio > 0) or (self.rng.rand() < math.exp(log_ratio))
        self.accepted += int(accepted)
        return accepted

    def acceptance_rate(self):
        return self.accepted / self.steps if self.steps > 0 else 0.0

    def reset_(self):
        if isinstance(self.ensemble, MulticanonicalEnsemble):     # algorithms/multicanonical.jl:27-33
            self.ensemble.histogram.values[...] = 0
        self.steps = 0
        self.accepted = 0


class Metropolis(ImportanceSampling):
    """metropolis.jl:63-96; accept!(alg, dE) = _accept!(logweight(ensemble, dE)) (:14-17)."""
    kind = "metropolis"

    def _accept_delta(self, delta_state):
        return self._accept(self.ensemble.logweight(delta_state))


class Glauber(Metropolis):
    """metropolis.jl:108-127: accepted = rand < logistic(log_ratio), always one draw."""
    kind = "glauber"

    def _accept_delta(self, delta_state):
        log_ratio = self.ensemble.logweight(delta_state)
        self.steps += 1
        accepted = self.rng.rand() < logistic(log_ratio)
        self.accepted += int(accepted)
        return accepted


class HeatBath:
    """heat_bath.jl:11-17: fields rng, beta, steps (no `accepted`)."""
    kind = "heatbath"

    def __init__(self, rng, beta):
        self.rng = rng
        self.beta = beta
        self.steps = 0

    def reset_(self):
        self.steps = 0


def Multicanonical(rng, bins, init=0.0):
    """algorithms/multicanonical.jl:9-25"""
    ens = bins if isinstance(bins, MulticanonicalEnsemble) else MulticanonicalEnsemble(bins, init=init)
    return ImportanceSampling(rng, ens)


def WangLandau(rng, bins_or_logweight, init=0.0, logf=1.0):
    """algorithms/wang_landau.jl:10-18"""
    ens = bins_or_logweight if isinstance(bins_or_logweight, WangLandauEnsemble) else \
        WangLandauEnsemble(bins_or_logweight, init=init, logf=logf)
    return ImportanceSampling(rng, ens)


def acceptance_rate(alg):
    return alg.acceptance_rate()


def reset_(alg):
    return alg.reset_()


def default_rng(seed=0, chain=0):
    return PhiloxRNG(seed, chain)
