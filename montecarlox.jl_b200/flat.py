"""Flat-histogram sampling on the device: multicanonical and Wang-Landau chains.

Mirrors Multicanonical / WangLandau (src/algorithms/{multicanonical,wang_landau}.jl), their
ensembles' update! / reset! / record_visit! (src/ensembles/{multicanonical,wang_landau}.jl) and the
parallel variant (src/algorithms/parallel_multicanonical.jl: merge_histograms!,
distribute_logweight!).  The host ensemble objects stay the source of truth for user code
(`ens.logweight_table.values`, `ens.histogram.values` are read by the examples,
docs/src/examples/spin_systems/muca_Ising2D.jl:89-90); DeviceFlat mirrors them on the GPU and
copies back after every call."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib
from .binned_object import DiscreteBinning
from .ensembles import AbstractEnsemble, BoltzmannEnsemble, MulticanonicalEnsemble, WangLandauEnsemble
from .rng import PhiloxRNG


class PairBoltzmannSpin2Ensemble(AbstractEnsemble):
    """The CustomEnsemble of docs/src/examples/spin_systems/muca_BlumeCapel.jl:43-66:
    logweight((H1, H2)) = logweight(pair, H1) + logweight(spin2, H2), visits recorded on H2."""

    def __init__(self, pair, spin2, record_visits=True):
        assert isinstance(pair, BoltzmannEnsemble) and isinstance(spin2, MulticanonicalEnsemble)
        self.pair, self.spin2, self.record_visits = pair, spin2, bool(record_visits)

    @property
    def should_record_visit(self):
        return self.record_visits

    def logweight(self, H):
        return self.pair.logweight(H[0]) + self.spin2.logweight(H[1])

    def record_visit_(self, H_vis):
        self.spin2.record_visit_(H_vis[1])

    def update_(self, **kw):
        self.spin2.update_(**kw)


def _table_ensemble(ens):
    return ens.spin2 if isinstance(ens, PairBoltzmannSpin2Ensemble) else ens


class DeviceFlat:
    """Device mirror of one flat-histogram ensemble bound to a (batched) lattice."""

    def __init__(self, sys, alg, policy=0):
        ens = alg.ensemble
        tab = _table_ensemble(ens)
        if isinstance(tab, MulticanonicalEnsemble):
            kind = _lib.FLAT_MUCA
        elif isinstance(tab, WangLandauEnsemble):
            kind = _lib.FLAT_WANG_LANDAU
        else:
            raise ValueError("flat-histogram sweeps need a Multicanonical or WangLandau ensemble")
        b = tab.logweight_table.bins
        if len(b) != 1 or not isinstance(b[0], DiscreteBinning) or not isinstance(b[0].start, (int, np.integer)):
            raise ValueError("device flat-histogram tables need 1-D integer DiscreteBinning (start:step:stop)")
        if kind == _lib.FLAT_WANG_LANDAU and sys.nchains != 1:
            raise ValueError("Wang-Landau keeps one table per chain; bind one chain per ensemble")
        if not isinstance(alg.rng, PhiloxRNG):
            raise ValueError("device sweeps need alg.rng::PhiloxRNG")
        if isinstance(ens, PairBoltzmannSpin2Ensemble):
            obs, beta_pair = _lib.OBS_SPIN2_WITH_PAIR_BOLTZMANN, float(ens.pair.beta)
        else:
            obs, beta_pair = _lib.OBS_ENERGY, 0.0
        self.sys, self.alg, self.tab, self.kind = sys, alg, tab, kind
        h = C.c_void_p()
        check(lib().mcx_flat_create(sys.h_lat, kind, obs, int(b[0].start), int(b[0].step), int(b[0].num), beta_pair,
                                    int(policy), C.byref(h)))
        self.h = h
        check(lib().mcx_lattice_set_first_chain_id(sys.h_lat, alg.rng.chain))
        sys._first_chain = alg.rng.chain       # what a checkpoint of the lattice must restore
        sys.set_rng(alg.rng.seed)

    def close(self):
        """Free the device mirror (idempotent).  Must happen before the lattice is destroyed; the
        owning system calls this from its own finaliser."""
        h, self.h = self.h, None
        if h is not None:
            try:
                lib().mcx_flat_destroy(h)
            except Exception:
                pass

    def __del__(self):
        self.close()

    def push(self):
        lw = np.ascontiguousarray(self.tab.logweight_table.values, dtype=np.float64)
        check(lib().mcx_flat_set_logweight(self.h, lw.ctypes.data))
        if self.kind == _lib.FLAT_WANG_LANDAU:
            check(lib().mcx_flat_set_logf(self.h, float(self.tab.logf)))

    def pull(self):
        lw = self.tab.logweight_table.values
        check(lib().mcx_flat_get_logweight(self.h, lw.ctypes.data))
        if self.kind == _lib.FLAT_MUCA:
            hv = self.tab.histogram.values
            check(lib().mcx_flat_get_histogram(self.h, hv.ctypes.data))

    def sweep_(self, nsweeps):
        before = self.sys._sums()[3].copy()
        self.push()
        if self.kind == _lib.FLAT_MUCA and not np.any(self.tab.histogram.values):
            check(lib().mcx_flat_reset_histogram(self.h))
        check(lib().mcx_flat_sweep(self.h, int(nsweeps)))
        self.pull()
        self.alg.steps += int(nsweeps) * self.sys.N * self.sys.nchains
        self.alg.accepted += int((self.sys._sums()[3] - before).sum())

    def device_histogram(self):
        """zero-copy int64 torch view of the device histogram (for the NCCL all-reduce)."""
        import torch
        from .parallel import _as_torch
        p, n = C.c_void_p(), C.c_int64()
        check(lib().mcx_flat_device_histogram(self.h, C.byref(p), C.byref(n)))
        return _as_torch(p.value, n.value, torch.int64, self.sys.ctx.device)


def flat_for(sys, alg, policy=0):
    key = id(alg)
    cache = sys.__dict__.setdefault("_flat_cache", {})
    if key not in cache or cache[key].alg is not alg:
        cache[key] = DeviceFlat(sys, alg, policy)
    return cache[key]


# --------------------------------------------------------------------------- parallel multicanonical
def ParallelMulticanonical(backend, alg):
    """parallel_multicanonical.jl:21-30"""
    from .parallel import ParallelChains
    algs = alg if isinstance(alg, (list, tuple)) else [alg]
    for a in algs:
        if not isinstance(_table_ensemble(a.ensemble), MulticanonicalEnsemble):
            raise TypeError("ParallelMulticanonical needs MulticanonicalEnsemble algorithms")
    return ParallelChains(backend, algs)


def merge_histograms_(pc, flat=None):
    """merge_histograms! (parallel_multicanonical.jl:38-52).  Threads-style list of algorithms: sum
    into the root chain.  GPUBackend with a DeviceFlat: the chains of one rank already share one
    device histogram; ranks are summed with one NCCL all-reduce (every rank gets the total, which
    also replaces the Bcast of distribute_logweight!)."""
    from .parallel import GPUBackend
    if flat is not None and isinstance(pc.backend, GPUBackend):
        if pc.backend.size > 1:
            pc.backend.all_reduce_sum(flat.device_histogram())
        flat.pull()
        return None
    r = pc.root_chain()
    h_root = _table_ensemble(pc.algorithm(r).ensemble).histogram.values
    for i, a in enumerate(pc.algs):
        if i != r:
            h_root += _table_ensemble(a.ensemble).histogram.values
    return None


def distribute_logweight_(pc, flat=None):
    """distribute_logweight! (parallel_multicanonical.jl:59-73)"""
    r = pc.root_chain()
    root_lw = _table_ensemble(pc.algorithm(r).ensemble).logweight_table.values
    for i, a in enumerate(pc.algs):
        if i != r:
            _table_ensemble(a.ensemble).logweight_table.values[...] = root_lw
    return None
