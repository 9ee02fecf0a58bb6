"""Ensembles (src/ensembles/*.jl): log-weights used by the acceptance rules."""
import math

import numpy as np

from .binned_object import BinnedObject


class AbstractEnsemble:
    def __eq__(self, other):                      # abstract_ensemble.jl:8-10: field-wise ==
        return type(self) is type(other) and self.__dict__.keys() == other.__dict__.keys() and all(
            _eq(self.__dict__[k], other.__dict__[k]) for k in self.__dict__)

    def logweight(self, x):
        raise ValueError("logweight not implemented for ensemble type %s" % type(self).__name__)

    def update_(self, *a, **k):
        raise ValueError("update! not implemented for ensemble type %s" % type(self).__name__)

    should_record_visit = False

    def record_visit_(self, x_vis):
        return None


def _eq(a, b):
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        return np.array_equal(a, b)
    return a == b


class BoltzmannEnsemble(AbstractEnsemble):
    """logweight(E) = -beta * E (boltzmann.jl:7-28)."""

    def __init__(self, beta=None, T=None):
        if (beta is None) == (T is None):
            raise ValueError("Specify exactly one of `beta`/`β` or `T`")
        self.beta = beta if beta is not None else 1.0 / T

    def logweight(self, E):
        if isinstance(E, (list, tuple, np.ndarray)):
            return -self.beta * sum(E)
        return -self.beta * E


class FunctionEnsemble(AbstractEnsemble):
    def __init__(self, f):
        self.f = f

    def logweight(self, x):
        return self.f(x)


def _as_ensemble(e):
    return e if isinstance(e, AbstractEnsemble) else FunctionEnsemble(e)


class MulticanonicalEnsemble(AbstractEnsemble):
    """Tabulated log-weight + visit histogram (ensembles/multicanonical.jl:1-44)."""

    def __init__(self, bins, init=0.0, histogram=None, record_visits=True):
        self.logweight_table = bins if isinstance(bins, BinnedObject) else BinnedObject(bins, float(init))
        self.histogram = histogram if histogram is not None else self.logweight_table.zero()
        if self.histogram.bins != self.logweight_table.bins:
            raise AssertionError("BinnedObject objects must have the same bins")
        self.record_visits = bool(record_visits)

    @property
    def should_record_visit(self):
        return self.record_visits

    def logweight(self, x):
        return self.logweight_table(x) if not isinstance(x, tuple) else self.logweight_table(*x)

    def record_visit_(self, x_vis):
        self.histogram[x_vis] = self.histogram[x_vis] + 1       # h[x] += 1 (:25-30)

    def update_(self, mode="simple"):
        if mode != "simple":
            raise ValueError("unsupported mode=%s, currently only :simple" % mode)
        h = self.histogram.values
        logh = np.zeros_like(h)
        pos = h > 0
        logh[pos] = [math.log(v) for v in h[pos]]
        self.logweight_table.values -= logh                    # lw -= (h > 0 ? log(h) : 0) (:32-44)


class WangLandauEnsemble(AbstractEnsemble):
    """Tabulated log-weight + modification factor logf (ensembles/wang_landau.jl:1-23)."""

    def __init__(self, bins, init=0.0, logf=1.0):
        self.logweight_table = bins if isinstance(bins, BinnedObject) else BinnedObject(bins, float(init))
        self.logf = float(logf)

    def logweight(self, x):
        return self.logweight_table(x)

    def update_(self, power=0.5):
        self.logf *= power


def logweight(ens, x):
    return _as_ensemble(ens).logweight(x)
