"""BinnedObject (src/infrastructure/binned_object.jl): an N-D table indexed by coordinate.
The device path uses the discrete 1-D integer case; the host mirror keeps the reference's
indexing rules for discrete (start, step, num) and continuous (edges) bins."""
import bisect

import numpy as np


class DiscreteBinning:
    """binned_object.jl:13-17"""

    def __init__(self, start, step, num):
        self.start, self.step, self.num = start, step, int(num)

    def index(self, x):
        # Integer bins: div(x - start, step) + 1 (:22-24); otherwise Int(round(...)) + 1 (:18-20)
        if isinstance(self.start, (int, np.integer)) and isinstance(x, (int, np.integer)):
            d = int(x) - int(self.start)
            q = abs(d) // abs(int(self.step))
            q = q if (d >= 0) == (self.step > 0) else -q     # Julia div truncates toward zero
            return q + 1
        return int(np.rint((x - self.start) / self.step)) + 1

    def centers(self):
        return [self.start + self.step * k for k in range(self.num)]

    def __eq__(self, other):
        return isinstance(other, DiscreteBinning) and (self.start, self.step, self.num) == (other.start, other.step, other.num)


class ContinuousBinning:
    """binned_object.jl:31-38"""

    def __init__(self, edges):
        self.edges = [float(e) for e in edges]
        self.cent = [(a + b) * 0.5 for a, b in zip(self.edges[:-1], self.edges[1:])]
        self.num = len(self.edges) - 1

    def index(self, x):
        return bisect.bisect_right(self.edges, x)      # searchsortedlast

    def centers(self):
        return self.cent

    def __eq__(self, other):
        return isinstance(other, ContinuousBinning) and self.edges == other.edges


def _bin_from_domain(d, interpretation):
    d = list(d) if not isinstance(d, range) else d
    if len(d) < 2:
        raise ValueError("Cannot create bins from a single value.")
    is_int = all(isinstance(v, (int, np.integer)) for v in (d[0], d[1], d[-1]))
    if interpretation == "auto":
        interpretation = "discrete" if is_int else "continuous"
    if interpretation == "discrete":
        steps = [b - a for a, b in zip(d[:-1], d[1:])] if not isinstance(d, range) else [d.step]
        if any(s != steps[0] for s in steps):
            raise ValueError("Non-equidistant discrete bins not supported without ExplicitBinning.")
        return DiscreteBinning(d[0], steps[0], len(d))
    if interpretation == "continuous":
        return ContinuousBinning(d)
    raise ValueError("Invalid interpretation=%s. Use :auto, :discrete, or :continuous." % interpretation)


class BinnedObject:
    """BinnedObject(domain, init) (binned_object.jl:98-122); 1-D or a tuple of domains."""

    def __init__(self, domain, init=0.0, interpretation="auto"):
        domains = domain if isinstance(domain, tuple) else (domain,)
        self.bins = tuple(_bin_from_domain(d, interpretation) for d in domains)
        if any(type(b) is not type(self.bins[0]) for b in self.bins):
            raise AssertionError("All bins must be of the same type for NTuple type stability.")
        self.values = np.full(tuple(b.num for b in self.bins), init, dtype=np.float64)

    def _idx(self, xs):
        xs = xs if isinstance(xs, tuple) else (xs,)
        idx = tuple(b.index(x) - 1 for b, x in zip(self.bins, xs))
        for i, b in zip(idx, self.bins):
            if i < 0 or i >= b.num:                     # no clamp: BoundsError (test_multicanonical.jl:39-43)
                raise IndexError("BoundsError: attempt to access BinnedObject at %r" % (xs,))
        return idx

    def __call__(self, *xs):
        return float(self.values[self._idx(tuple(xs))])

    def __getitem__(self, xs):
        return float(self.values[self._idx(xs)])

    def __setitem__(self, xs, v):
        self.values[self._idx(xs)] = v

    @property
    def size(self):
        return self.values.shape

    def zero(self):
        out = BinnedObject.__new__(BinnedObject)
        out.bins = self.bins
        out.values = np.zeros_like(self.values)
        return out

    def copy(self):
        out = self.zero()
        out.values[...] = self.values
        return out

    def __eq__(self, other):
        return isinstance(other, BinnedObject) and self.bins == other.bins and np.array_equal(self.values, other.values, equal_nan=False)


def get_centers(bo, dim=1):
    return bo.bins[dim - 1].centers()


def get_values(bo):
    return bo.values
