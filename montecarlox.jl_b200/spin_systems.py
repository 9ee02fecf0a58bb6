"""SpinSystems mirror (SpinSystems/src/{ising,blume_capel,abstractions}.jl) backed by device lattices.

Same names and observable semantics as the reference (`Ising(dims)`, `BlumeCapel(dims)`,
`IsingLatticeOptim(Lx, Ly)`, `init!`, `energy`, `magnetization`, field `spins`), plus the one new
verb the checkerboard needs: `sweep_(sys, alg, nsweeps)` = nsweeps * N single-spin attempts.
All compute goes through libmcx_b200 (C ABI); nothing here falls back to the CPU."""
import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import check, lib
from .rng import PhiloxRNG
from .tables import beta_of, build_table, rule_of

_contexts = {}


class Context:
    """One CUDA device + stream (mcx_ctx).  `stream` may be a raw cudaStream_t integer, e.g.
    torch.cuda.current_stream().cuda_stream, so the host runtime and the kernels share a stream."""

    def __init__(self, device=0, stream=None):
        h = C.c_void_p()
        check(lib().mcx_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self.h, self.device = h, int(device)

    def sync(self):
        check(lib().mcx_ctx_sync(self.h))

    def set_stream(self, stream):
        check(lib().mcx_ctx_set_stream(self.h, C.c_void_p(stream) if stream else None))

    def info(self):
        sm, ma, mi, mem = C.c_int32(), C.c_int32(), C.c_int32(), C.c_uint64()
        check(lib().mcx_ctx_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "total_mem": mem.value}

    def launch_count(self):
        n = C.c_uint64()
        check(lib().mcx_ctx_launch_count(self.h, C.byref(n)))
        return n.value

    def close(self):
        """Explicit teardown (parallel_backends.jl:104 finalize!).  Not done implicitly: lattices keep
        raw pointers to their context, and interpreter shutdown destroys objects in no fixed order."""
        h, self.h = self.h, None
        if h is not None:
            check(lib().mcx_ctx_destroy(h))


def default_context(device=0):
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]


_STORAGE = {"int8": _lib.STORAGE_INT8, "bit": _lib.STORAGE_BIT}
_INIT = {"up": _lib.INIT_UP, "down": _lib.INIT_DOWN, "zero": _lib.INIT_ZERO, "random": _lib.INIT_RANDOM}


class AbstractSpinSystem:
    """abstractions.jl:12; a batch of `nchains` independent lattices of identical shape."""
    model = None

    def __init__(self, dims, J=1, h=0, D=0, nchains=1, ctx=None, periodic=True, storage="int8"):
        if not periodic:
            raise ValueError("the checkerboard backend supports periodic lattices only")
        if storage not in _STORAGE:
            raise ValueError("storage must be 'int8' or 'bit' (got %r)" % (storage,))
        self.storage = storage
        self.dims = [int(d) for d in dims]
        self.J, self.h, self.D = J, h, D
        self.nchains = int(nchains)
        self.N = int(np.prod(self.dims))
        self.ctx = ctx or default_context()
        hd = C.c_void_p()
        d = (C.c_int32 * len(self.dims))(*self.dims)
        check(lib().mcx_lattice_create(self.ctx.h, self.model, len(self.dims), d, self.nchains, _STORAGE[storage],
                                       C.byref(hd)))
        self.h_lat = hd
        check(lib().mcx_lattice_set_couplings(hd, float(J), float(h), float(D)))
        self._rule_key = None
        self._attached = weakref.WeakSet()

    def __del__(self):
        try:
            for f in self.__dict__.get("_flat_cache", {}).values():
                f.close()                      # device mirrors first: they point into the lattice
            lib().mcx_lattice_destroy(self.h_lat)
        except Exception:
            pass

    # ---- serialisation (checkpoint!/restore_checkpoint, checkpointing.jl:48-101)
    def __getstate__(self):
        seed, nxt = C.c_uint64(), C.c_uint64()
        check(lib().mcx_get_rng(self.h_lat, C.byref(seed), C.byref(nxt)))
        pair, spin, spin2, acc, steps = self._sums()
        return {"cls_dims": self.dims, "J": self.J, "h": self.h, "D": self.D, "nchains": self.nchains,
                "device": self.ctx.device, "spins": self.spins.copy(), "seed": seed.value, "next_sweep": nxt.value,
                "labels": self.get_labels(), "rule": getattr(self, "_rule_tables", None),
                "first_chain": getattr(self, "_first_chain", 0), "storage": self.storage,
                "accepted": np.atleast_1d(acc).astype(np.int64), "steps": int(np.atleast_1d(steps)[0])}

    def __setstate__(self, st):
        AbstractSpinSystem.__init__(self, st["cls_dims"], st["J"], st["h"], st["D"], st["nchains"],
                                    default_context(st["device"]), storage=st.get("storage", "int8"))
        self.spins = st["spins"]
        if st["rule"] is not None:
            self.set_rule(*st["rule"])
            self.set_labels(st["labels"])
        check(lib().mcx_lattice_set_first_chain_id(self.h_lat, st["first_chain"]))
        self._first_chain = st["first_chain"]
        check(lib().mcx_set_rng(self.h_lat, st["seed"], st["next_sweep"]))
        if "accepted" in st:        # the lattice-side counters of the bound algorithm (importance_sampling.jl:26-27)
            acc = np.ascontiguousarray(st["accepted"], dtype=np.int64)
            check(lib().mcx_set_counters(self.h_lat, acc.ctypes.data, int(st["steps"])))

    # ---- sys.spins
    def _shape(self, a):
        return a.reshape(self.N) if self.nchains == 1 else a.reshape(self.nchains, self.N)

    @property
    def spins(self):
        out = np.empty(self.nchains * self.N, dtype=np.int8)
        check(lib().mcx_lattice_download(self.h_lat, out.ctypes.data))
        return self._shape(out)

    @spins.setter
    def spins(self, v):
        v = np.ascontiguousarray(v, dtype=np.int8).reshape(-1)
        if v.size != self.nchains * self.N:
            raise ValueError("spins must have nchains*N = %d entries" % (self.nchains * self.N))
        check(lib().mcx_lattice_upload(self.h_lat, v.ctypes.data))

    # ---- the same field at one bit per spin on the host side (site i = bit i & 7 of byte i >> 3, 1 = up): an eighth of
    # the PCIe traffic; numpy's packbits(bitorder="little") of (spins > 0) is this format
    @property
    def spin_bits(self):
        out = np.empty(self.nchains * self.N // 8, dtype=np.uint8)
        check(lib().mcx_lattice_download_bits(self.h_lat, out.ctypes.data))
        return out if self.nchains == 1 else out.reshape(self.nchains, self.N // 8)

    @spin_bits.setter
    def spin_bits(self, v):
        v = np.ascontiguousarray(v, dtype=np.uint8).reshape(-1)
        if v.size * 8 != self.nchains * self.N:
            raise ValueError("spin_bits must have nchains*N/8 = %d bytes" % (self.nchains * self.N // 8))
        check(lib().mcx_lattice_upload_bits(self.h_lat, v.ctypes.data))

    def upload_bits_begin(self, host_ptr):
        """`upload_begin` for a bit buffer (raw pointer); `upload_commit` makes it the lattice."""
        check(lib().mcx_lattice_upload_bits_begin(self.h_lat, C.c_void_p(host_ptr)))

    def upload_from(self, host_ptr):
        """Upload from a raw host pointer (e.g. pinned memory) without touching numpy."""
        check(lib().mcx_lattice_upload(self.h_lat, C.c_void_p(host_ptr)))

    def upload_begin(self, host_ptr):
        """Start copying a host buffer (raw pointer, pinned memory for the copy to overlap) towards the device
        without touching the lattice: sweeps already queued keep running.  `upload_commit` makes it the lattice."""
        check(lib().mcx_lattice_upload_begin(self.h_lat, C.c_void_p(host_ptr)))

    def upload_commit(self):
        check(lib().mcx_lattice_upload_commit(self.h_lat))

    # ---- observables (ising.jl:17-18, blume_capel.jl:18-19)
    def _sums(self):
        n = self.nchains
        a = [np.empty(n, dtype=np.int64) for _ in range(5)]
        check(lib().mcx_observables(self.h_lat, *[x.ctypes.data for x in a]))
        return a

    def _scalar(self, arr):
        return arr[0].item() if self.nchains == 1 else arr

    def energy(self, full=False):
        if full:
            check(lib().mcx_recompute(self.h_lat))
        pair, spin, spin2, _, _ = self._sums()
        return self._scalar(self._energy_from(pair, spin, spin2))

    def _energy_from(self, pair, spin, spin2):
        raise NotImplementedError

    def magnetization(self, full=False):
        if full:
            check(lib().mcx_recompute(self.h_lat))
        return self._scalar(self._sums()[1])

    def pair_sum(self):
        return self._scalar(self._sums()[0])

    def spin2_sum(self):
        return self._scalar(self._sums()[2])

    def accepted(self):
        return self._scalar(self._sums()[3])

    # ---- init!(sys, type; rng) (ising.jl:74-78, blume_capel.jl:106-110)
    def init_(self, type, rng=None):
        if type not in _INIT:
            raise RuntimeError("Unknown initialization type: %s" % type)
        seed = 0
        if type == "random":
            assert rng is not None, "Random initialization requires rng"
            if not isinstance(rng, PhiloxRNG):
                raise ValueError("device init needs a PhiloxRNG (counter-based); got %s" % type(rng).__name__)
            seed = rng.seed
            check(lib().mcx_lattice_set_first_chain_id(self.h_lat, rng.chain))
            self._first_chain = rng.chain
        check(lib().mcx_lattice_init(self.h_lat, _INIT[type], seed))
        return self

    # ---- rule + rng plumbing
    def set_rule(self, rule, tables):
        """tables: uint64 [n_labels, table_len]"""
        t = np.ascontiguousarray(tables, dtype=np.uint64)
        if t.ndim == 1:
            t = t[None, :]
        check(lib().mcx_set_rule(self.h_lat, rule, t.ctypes.data, t.shape[0], t.shape[1]))
        self._rule_key = None
        self._rule_tables = (rule, t.copy())

    def set_labels(self, labels):
        a = np.ascontiguousarray(labels, dtype=np.int32)
        check(lib().mcx_set_labels(self.h_lat, a.ctypes.data))

    def get_labels(self):
        a = np.empty(self.nchains, dtype=np.int32)
        check(lib().mcx_get_labels(self.h_lat, a.ctypes.data))
        return a

    def set_rng(self, seed, next_sweep=None):
        if next_sweep is None:
            next_sweep = self.sweep_index
        check(lib().mcx_set_rng(self.h_lat, int(seed), int(next_sweep)))

    @property
    def sweep_index(self):
        s, n = C.c_uint64(), C.c_uint64()
        check(lib().mcx_get_rng(self.h_lat, C.byref(s), C.byref(n)))
        return n.value

    def set_tracking(self, on):
        check(lib().mcx_set_tracking(self.h_lat, int(bool(on))))

    def sync(self):
        self.ctx.sync()

    def _bind_alg(self, alg):
        """Install alg's rule table and RNG stream on the lattice (cached)."""
        rng = alg.rng
        if not isinstance(rng, PhiloxRNG):
            raise ValueError("checkerboard sweeps need alg.rng::PhiloxRNG (counter-based); got %s" % type(rng).__name__)
        key = (alg.kind, beta_of(alg), self.J, self.h, self.D, rng.seed, rng.chain)
        if key != self._rule_key:
            T = build_table(self.model, rule_of(alg), len(self.dims), beta_of(alg), self.J, self.h, self.D)
            self.set_rule(rule_of(alg), T)
            check(lib().mcx_lattice_set_first_chain_id(self.h_lat, rng.chain))
            self._first_chain = rng.chain
            self.set_rng(rng.seed)
            self._rule_key = key


class AbstractIsing(AbstractSpinSystem):
    model = _lib.ISING

    def _energy_from(self, pair, spin, spin2):
        # -sum_pair_interactions - sum_field_interactions (ising.jl:175-177)
        e = -(self.J * pair)
        if self.h != 0:
            e = e - self.h * spin
        return e


class Ising(AbstractIsing):
    """Ising(dims; J=1, periodic=true, h=0) (ising.jl:406-417) on a periodic grid.  `storage="bit"` keeps the spins at
    one bit each on the device (2-D / 3-D, Lx % 32 == 0); trajectories and observables do not depend on it.
    Open boundaries (`periodic=False`), a per-site field vector `h` or one coupling per edge (vector `J`) give the grid
    graph of the reference (Graphs.SimpleGraphs.grid) on the general-topology path (graph_systems.IsingGraph)."""

    def __new__(cls, dims=None, J=1, h=0, D=0, nchains=1, ctx=None, periodic=True, storage="int8"):
        if dims is not None and (not periodic or np.ndim(h) != 0 or np.ndim(J) != 0):   # dims None: unpickling
            from .graph_systems import IsingGraph, grid_graph
            edges, n = grid_graph(dims, periodic)
            return IsingGraph(edges, n, J=J, h=h, nchains=nchains, ctx=ctx)      # not an instance of cls: __init__ is skipped
        return super().__new__(cls)


class IsingLatticeOptim(AbstractIsing):
    """IsingLatticeOptim(Lx, Ly) (ising.jl:430-461): 2-D, J=1, h=0."""

    def __init__(self, Lx, Ly, nchains=1, ctx=None, storage="int8"):
        super().__init__([Lx, Ly], 1, 0, 0, nchains, ctx, storage=storage)


class AbstractBlumeCapel(AbstractSpinSystem):
    model = _lib.BLUME_CAPEL

    def _energy_from(self, pair, spin, spin2):
        # -sum_pair - sum_field + D*sum_spins2 (blume_capel.jl:222-224); sum_pair is Float64 there
        e = -(float(self.J) * pair)
        if self.h != 0:
            e = e - self.h * spin
        return e + self.D * spin2


class BlumeCapel(AbstractBlumeCapel):
    """BlumeCapel(dims; J=1, D=0, periodic=true, h=0) (blume_capel.jl:500-511)."""

    def __init__(self, dims, J=1, D=0, h=0, nchains=1, ctx=None, periodic=True):
        super().__init__(dims, J, h, D, nchains, ctx, periodic)


def energy(sys, full=False):
    return sys.energy(full=full)


def magnetization(sys, full=False):
    return sys.magnetization(full=full)


def init_(sys, type, rng=None):
    return sys.init_(type, rng=rng)


def sweep_(sys, alg, nsweeps=1):
    """nsweeps * N attempts in checkerboard order = `for _ in 1:N*nsweeps; spin_flip!(sys, alg); end`
    (pt_Ising2D.jl:52-57) with the update order changed from random-site to checkerboard.
    Asynchronous; `alg.steps` / `alg.accepted` follow the reference's counters
    (importance_sampling.jl:80-85) and are read back lazily."""
    if hasattr(sys, "_graph_sweep"):           # general topology (graph_systems.py): coloured sweeps
        return sys._graph_sweep(alg, nsweeps)
    if hasattr(alg, "sweep_system_"):          # ReplicaExchange: every replica with its own label
        return alg.sweep_system_(sys, nsweeps)
    ens = getattr(alg, "ensemble", None)
    if ens is not None and not hasattr(ens, "beta"):
        from .flat import flat_for             # Multicanonical / WangLandau: serial chains, FLAT stream
        return flat_for(sys, alg).sweep_(nsweeps)
    sys._bind_alg(alg)
    before = None
    if getattr(alg, "_track_counters", True):
        before = sys._sums()[3].copy()
    check(lib().mcx_sweep(sys.h_lat, int(nsweeps)))
    # one algorithm object drives every chain of a batch: its counters are sums over the chains, so that
    # acceptance_rate(alg) = accepted / steps stays a rate (importance_sampling.jl:95-101)
    alg.steps += int(nsweeps) * sys.N * sys.nchains
    if before is not None and hasattr(alg, "accepted"):
        alg.accepted += int((sys._sums()[3] - before).sum())
    return None
