"""General topologies: IsingGraph / IsingMatrix with site fields (SpinSystems/src/ising.jl:86-417) on the device.

`IsingGraph(edges, n, J; h)` mirrors `Ising(graph::SimpleGraph, J::Real; h)`, `IsingMatrix(J; h)` mirrors
`Ising(J::SparseMatrixCSC; h)`, `IsingGraph(edges, n, Jvec)` with one coupling per edge mirrors `Ising(graph, J::Vector)`
(which the reference turns into an IsingMatrix, ising.jl:383-404), and `grid_graph(dims, periodic)` is
`Graphs.SimpleGraphs.grid` -- so `Ising(dims; periodic=false)` is `IsingGraph(*grid_graph(dims, False), J)`.
All compute goes through libmcx_b200 (mcx_graph_*); nothing here falls back to the CPU."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib
from .rng import PhiloxRNG
from .tables import beta_of, rule_of

_INIT = {"up": _lib.INIT_UP, "down": _lib.INIT_DOWN, "random": _lib.INIT_RANDOM}


def grid_graph(dims, periodic=True):
    """(edges [m, 2] 0-based with i < j, n) of Graphs.SimpleGraphs.grid(dims; periodic): site i = x + Lx*(y + Ly*z);
    a periodic dimension of length 2 contributes one edge per pair, as a SimpleGraph stores it."""
    dims = [int(d) for d in dims]
    n = int(np.prod(dims))
    idx = np.arange(n, dtype=np.int64)
    strides = np.cumprod([1] + dims[:-1])
    pairs = []
    for L, st in zip(dims, strides):
        x = (idx // st) % L
        if periodic and L > 2:
            j = idx + (((x + 1) % L) - x) * st
            pairs.append(np.stack([idx, j], axis=1))
        elif L > 1:
            keep = x + 1 < L
            pairs.append(np.stack([idx[keep], idx[keep] + st], axis=1))
    e = np.concatenate(pairs) if pairs else np.zeros((0, 2), dtype=np.int64)
    e = np.sort(e, axis=1)
    e = np.unique(e, axis=0)
    return e, n


def _csr_from_edges(edges, n, Jvec=None):
    """symmetric CSR (rowptr, col, val) with ascending neighbours per row -- Graphs.jl adjacency lists / CSC columns"""
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    if e.size and (e.min() < 0 or e.max() >= n):
        raise IndexError("edge endpoint outside [0, n)")
    src = np.concatenate([e[:, 0], e[:, 1]])
    dst = np.concatenate([e[:, 1], e[:, 0]])
    order = np.lexsort((dst, src))
    src, dst = src[order], dst[order]
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, src + 1, 1)
    rowptr = np.cumsum(rowptr)
    val = None
    if Jvec is not None:
        Jv = np.asarray(Jvec, dtype=np.float64)
        if Jv.size != len(e):
            raise AssertionError("Length of J vector must equal number of graph edges")
        val = np.concatenate([Jv, Jv])[order]
    return rowptr, dst.astype(np.int64), val


class AbstractGraphIsing:
    """AbstractIsing on a general topology; `nchains` independent copies."""

    def __init__(self, rowptr, col, val, J, h, nchains=1, ctx=None):
        from .spin_systems import default_context
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        self.col = np.ascontiguousarray(col, dtype=np.int64)
        self.val = None if val is None else np.ascontiguousarray(val, dtype=np.float64)
        self.N = len(self.rowptr) - 1
        self.nchains = int(nchains)
        self.J = J
        if np.ndim(h) == 0:
            self.h = h
            mode, hs, hv = (0 if h == 0 else 1), float(h), None
        else:
            hv = np.ascontiguousarray(h, dtype=np.float64)
            if hv.size != self.N:
                raise AssertionError("Field vector length must match number of spins")
            self.h, mode, hs = hv, 2, 0.0
        self.ctx = ctx or default_context()
        hd = C.c_void_p()
        check(lib().mcx_graph_create(self.ctx.h, self.N, self.rowptr.ctypes.data, self.col.ctypes.data,
                                     None if self.val is None else self.val.ctypes.data, float(J) if self.val is None else 0.0,
                                     mode, hs, None if hv is None else hv.ctypes.data, self.nchains, C.byref(hd)))
        self.h_graph = hd
        self._rule_key = None

    def __del__(self):
        try:
            lib().mcx_graph_destroy(self.h_graph)
        except Exception:
            pass

    def _shape(self, a):
        return a.reshape(self.N) if self.nchains == 1 else a.reshape(self.nchains, self.N)

    @property
    def spins(self):
        out = np.empty(self.nchains * self.N, dtype=np.int8)
        check(lib().mcx_graph_download(self.h_graph, out.ctypes.data))
        return self._shape(out)

    @spins.setter
    def spins(self, v):
        v = np.ascontiguousarray(v, dtype=np.int8).reshape(-1)
        if v.size != self.nchains * self.N:
            raise ValueError("spins must have nchains*N = %d entries" % (self.nchains * self.N))
        check(lib().mcx_graph_upload(self.h_graph, v.ctypes.data))

    def colours(self):
        """(ncolours, colour of every site) of the greedy colouring the sweeps follow"""
        n = C.c_int32()
        col = np.empty(self.N, dtype=np.int32)
        check(lib().mcx_graph_colours(self.h_graph, C.byref(n), col.ctypes.data))
        return n.value, col

    def _sums(self):
        k = self.nchains
        pair, field = np.empty(k), np.empty(k)
        spin, acc, steps = (np.empty(k, dtype=np.int64) for _ in range(3))
        check(lib().mcx_graph_observables(self.h_graph, pair.ctypes.data, spin.ctypes.data, field.ctypes.data, acc.ctypes.data,
                                          steps.ctypes.data))
        return pair, spin, field, acc, steps

    def _scalar(self, a):
        return a[0].item() if self.nchains == 1 else a

    def energy(self, full=False):
        """-sum_pair_interactions - sum_field_interactions (ising.jl:181-185); sums are always formed from the spins"""
        e = np.empty(self.nchains)
        check(lib().mcx_graph_energies(self.h_graph, e.ctypes.data))
        return self._scalar(e)

    def magnetization(self, full=False):
        return self._scalar(self._sums()[1])

    def accepted(self):
        return self._scalar(self._sums()[3])

    def init_(self, type, rng=None):
        if type not in _INIT:
            raise RuntimeError("Unknown initialization type: %s" % type)
        seed = 0
        if type == "random":
            assert rng is not None, "Random initialization requires rng"
            if not isinstance(rng, PhiloxRNG):
                raise ValueError("device init needs a PhiloxRNG (counter-based); got %s" % type(rng).__name__)
            seed = rng.seed
            check(lib().mcx_graph_set_rng(self.h_graph, rng.seed, self.sweep_index, rng.chain))
        check(lib().mcx_graph_init(self.h_graph, _INIT[type], seed))
        return self

    @property
    def sweep_index(self):
        s, n = C.c_uint64(), C.c_uint64()
        check(lib().mcx_graph_get_rng(self.h_graph, C.byref(s), C.byref(n)))
        return n.value

    def sync(self):
        self.ctx.sync()

    def _bind_alg(self, alg):
        rng = alg.rng
        if not isinstance(rng, PhiloxRNG):
            raise ValueError("coloured sweeps need alg.rng::PhiloxRNG (counter-based); got %s" % type(rng).__name__)
        key = (alg.kind, beta_of(alg), rng.seed, rng.chain)
        if key != self._rule_key:
            check(lib().mcx_graph_set_rule(self.h_graph, rule_of(alg), float(beta_of(alg))))
            check(lib().mcx_graph_set_rng(self.h_graph, rng.seed, self.sweep_index, rng.chain))
            self._rule_key = key

    def _graph_sweep(self, alg, nsweeps):
        self._bind_alg(alg)
        before = self._sums()[3].copy() if hasattr(alg, "accepted") else None
        check(lib().mcx_graph_sweep(self.h_graph, int(nsweeps)))
        alg.steps += int(nsweeps) * self.N * self.nchains
        if before is not None:
            alg.accepted += int((self._sums()[3] - before).sum())


class IsingGraph(AbstractGraphIsing):
    """Ising(graph, J::Real; h=0) (ising.jl:117-139): `edges` [m, 2] 0-based, `n` sites.  A vector J (one coupling per
    edge, in the order of `edges`) gives the IsingMatrix the reference builds for it (ising.jl:383-404)."""

    def __init__(self, edges, n, J=1, h=0, nchains=1, ctx=None):
        if np.ndim(J) == 0:
            rowptr, col, _ = _csr_from_edges(edges, n)
            super().__init__(rowptr, col, None, J, h, nchains, ctx)
        else:
            rowptr, col, val = _csr_from_edges(edges, n, J)
            super().__init__(rowptr, col, val, 0.0, h, nchains, ctx)


class IsingMatrix(AbstractGraphIsing):
    """Ising(J::SparseMatrixCSC; h=0) (ising.jl:264-293): J a scipy sparse matrix, a dense 2-D array (zeros = no coupling)
    or a (rowptr, col, val) CSR triple; must be square and symmetric (AssertionError otherwise)."""

    def __init__(self, J, h=0, nchains=1, ctx=None):
        if isinstance(J, tuple):
            rowptr, col, val = J
        else:
            try:
                import scipy.sparse as sp
                M = sp.csr_matrix(J)
            except ImportError:                       # dense fallback of the conversion only
                A = np.asarray(J, dtype=np.float64)
                M = None
            if M is not None:
                if M.shape[0] != M.shape[1]:
                    raise AssertionError("Sparse J must be square")
                M.sort_indices()
                rowptr, col, val = M.indptr, M.indices, M.data
            else:
                if A.ndim != 2 or A.shape[0] != A.shape[1]:
                    raise AssertionError("Sparse J must be square")
                rows, cols = np.nonzero(A)
                rowptr = np.zeros(A.shape[0] + 1, dtype=np.int64)
                np.add.at(rowptr, rows + 1, 1)
                rowptr, col, val = np.cumsum(rowptr), cols, A[rows, cols]
        super().__init__(rowptr, col, np.asarray(val, dtype=np.float64), 0.0, h, nchains, ctx)
