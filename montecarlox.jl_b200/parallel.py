"""Parallel coordination mirror (src/infrastructure/{parallel_backends,parallel_chains}.jl,
src/algorithms/{replica_exchange,parallel_tempering}.jl).

The reference's two backends are ThreadsBackend (a vector of algorithms on one host) and MPIBackend
(one chain per rank).  GPUBackend is the third: one process per GPU (torch.distributed, NCCL),
each rank holding a contiguous block of replica slots inside ONE batched device lattice.  Only the
per-replica energies cross NVLink (one all-gather per exchange); the swap decision is evaluated
identically on every rank from a counter-based u, and labels move, not lattices
(replica_exchange.jl:133)."""
import ctypes as C
import os
import math
from fractions import Fraction

import numpy as np

from . import _lib
from ._lib import check, lib
from .algorithms import ImportanceSampling, Metropolis
from .rng import PhiloxRNG, exchange_u
from .tables import build_table, rule_of, beta_of


# --------------------------------------------------------------------------- backends
class ThreadsBackend:
    """parallel_backends.jl:25-33"""

    def __init__(self, n):
        self.n = int(n)

    rank = 0
    is_root = True

    @property
    def size(self):
        return self.n


class GPUBackend:
    """One process per GPU.  `group` is a torch.distributed process group (None: default group if
    torch.distributed is initialised, else a single rank).  Mirrors rank/size/is_root of
    parallel_backends.jl:31-33,68-70."""

    def __init__(self, group=None, root=0):
        self.group, self.root = group, root
        try:
            import torch.distributed as dist
            self._dist = dist if dist.is_available() and dist.is_initialized() else None
        except Exception:
            self._dist = None

    @property
    def rank(self):
        return self._dist.get_rank(self.group) if self._dist else 0

    @property
    def size(self):
        return self._dist.get_world_size(self.group) if self._dist else 1

    @property
    def is_root(self):
        return self.rank == self.root

    def slots(self, n_global):
        """[first, first + count) replica slots owned by this rank (contiguous, equal blocks)."""
        return partition_slots(n_global, self.size, self.rank)

    def all_gather_inplace(self, tensor, first, count):
        """all-gather equal slices of a 1-D tensor in place (each rank owns [first, first+count))."""
        if not self._dist or self.size == 1:
            return
        # in place: the input is this rank's slice of the output (NCCL's in-place all-gather layout)
        if tensor.is_cuda:
            self._dist.all_gather_into_tensor(tensor, tensor[first:first + count], group=self.group)
        else:   # gloo (CPU tests) wants distinct buffers
            self._dist.all_gather_into_tensor(tensor, tensor[first:first + count].clone(), group=self.group)

    def all_reduce_sum(self, tensor):
        if not self._dist or self.size == 1:
            return
        self._dist.all_reduce(tensor, group=self.group)

    def all_reduce_max(self, tensor):
        if not self._dist or self.size == 1:
            return
        self._dist.all_reduce(tensor, op=self._dist.ReduceOp.MAX, group=self.group)

    def barrier(self):
        if self._dist and self.size > 1:
            self._dist.barrier(group=self.group)


def partition_slots(n_global, nranks, rank):
    if n_global % nranks != 0:
        raise ValueError("replicas (%d) must divide evenly over ranks (%d)" % (n_global, nranks))
    per = n_global // nranks
    return rank * per, per


# --------------------------------------------------------------------------- ParallelChains
class ParallelChains:
    """parallel_chains.jl:14-128 (host-side container; `algs` is a list for Threads/GPU backends)."""

    def __init__(self, backend, algs):
        self.backend, self.algs = backend, list(algs) if isinstance(algs, (list, tuple)) else [algs]

    @property
    def size(self):
        return len(self.algs) if not isinstance(self.backend, GPUBackend) else self._n_global()

    def _n_global(self):
        return len(self.algs)

    @property
    def rank(self):
        return self.backend.rank

    @property
    def is_root(self):
        return self.backend.is_root

    def algorithm(self, i=0):
        return self.algs[i]

    def root_chain(self):
        return 0

    def on_root(self, f):
        if self.is_root:
            try:
                return f(self.root_chain())
            except TypeError:
                return f()
        return None

    def with_parallel(self, f):
        return [f(i, a) for i, a in enumerate(self.algs)]

    def merge_(self, buf):
        """merge!(pc, buf): element-wise sum of a user array over ranks (parallel_chains.jl:121-128;
        identity for one rank, MPI.Allreduce! otherwise).  `buf` is a torch tensor or numpy array."""
        if isinstance(self.backend, GPUBackend) and self.backend.size > 1:
            import torch
            t = buf if isinstance(buf, torch.Tensor) else torch.from_numpy(buf)
            self.backend.all_reduce_sum(t)
        return buf


# --------------------------------------------------------------------------- replica exchange
def exchange_log_ratio(ens_i, ens_j, x_i, x_j):
    """replica_exchange.jl:110-113"""
    return (ens_i.logweight(x_j) - ens_i.logweight(x_i)) + (ens_j.logweight(x_i) - ens_j.logweight(x_j))


def _accept_exchange(log_ratio, u):
    """replica_exchange.jl:115"""
    return (log_ratio > 0) or (u < math.exp(log_ratio))


def attempt_exchange_pair_(alg_i, alg_j, x_i, x_j, u):
    """replica_exchange.jl:124-136"""
    if not math.isfinite(u):
        raise ValueError("shared random number `u` must be finite")
    accepted = _accept_exchange(exchange_log_ratio(alg_i.ensemble, alg_j.ensemble, x_i, x_j), u)
    if accepted:
        alg_i.ensemble, alg_j.ensemble = alg_j.ensemble, alg_i.ensemble
    return accepted


def _resolve_pair(my_index, stage, nranks):
    """replica_exchange.jl:138-149 -> (active, pair_id, partner_index)"""
    first = 1 if stage % 2 == 0 else 2
    offset = my_index - first
    if offset >= 0 and offset % 2 == 0 and my_index < nranks:
        return True, my_index, my_index + 1
    if offset > 0 and offset % 2 == 1 and my_index - 1 >= first:
        return True, my_index - 1, my_index - 1
    return False, 0, 0


class ReplicaExchange:
    """replica_exchange.jl:13-19.  indices[r] = 1-based ladder position held by slot r."""

    def __init__(self, backend, algs):
        algs = list(algs)
        n = len(algs)
        if n < 2:
            raise ValueError("need at least 2 algorithms for replica exchange")
        self.replica = ParallelChains(backend, algs)
        self.backend = backend
        self.stage = 0
        self.indices = np.arange(1, n + 1, dtype=np.int64)
        self.steps = np.zeros(n - 1, dtype=np.int64)
        self.accepted = np.zeros(n - 1, dtype=np.int64)
        self.round = 0
        self._pt = None          # device handle once attached to a lattice
        self._sys = None

    # -- reference accessors
    @property
    def size(self):
        return len(self.replica.algs)

    @property
    def rank(self):
        return self.backend.rank

    @property
    def is_root(self):
        return self.backend.is_root

    def algorithm(self, i=0):
        return self.replica.algs[i]

    def index(self, i=None):
        self._pull()
        return self.indices if i is None else int(self.indices[i])

    def acceptance_rates(self):
        self._pull()
        return [a / s if s > 0 else 0.0 for s, a in zip(self.steps, self.accepted)]

    def acceptance_rate(self):
        self._pull()
        tot = int(self.steps.sum())
        return float(self.accepted.sum()) / tot if tot > 0 else 0.0

    def reset_(self):
        self.stage = 0
        self.indices[:] = np.arange(1, self.size + 1)
        self.steps[:] = 0
        self.accepted[:] = 0
        self.round = 0
        if self._pt is not None:
            check(lib().mcx_pt_reset(self._pt))
        return self

    # -- host update!(rx, xs) (ThreadsBackend semantics, :158-178); also the oracle-checkable path
    def update_(self, xs=None):
        if xs is None:
            return self._update_device()
        if len(xs) != self.size:
            raise ValueError("xs must have length size(rx)")
        first = 1 if self.stage % 2 == 0 else 2
        for pair_id in range(first, self.size, 2):
            ri = int(np.nonzero(self.indices == pair_id)[0][0])
            rj = int(np.nonzero(self.indices == pair_id + 1)[0][0])
            self.steps[pair_id - 1] += 1
            rng = self.algorithm(ri).rng
            if isinstance(rng, PhiloxRNG):
                u = exchange_u(rng.seed, rng.chain, self.round)
            else:
                u = rng.rand()
            if attempt_exchange_pair_(self.algorithm(ri), self.algorithm(rj), xs[ri], xs[rj], u):
                self.accepted[pair_id - 1] += 1
                self.indices[ri], self.indices[rj] = self.indices[rj], self.indices[ri]
        self.stage = 1 - self.stage
        self.round += 1
        return None

    # -- device path
    def attach(self, sys):
        """Bind the ladder to a batched device lattice holding this rank's replica slots."""
        n = self.size
        first, count = self.backend.slots(n) if isinstance(self.backend, GPUBackend) else (0, n)
        if sys.nchains != count:
            raise ValueError("lattice holds %d chains but this rank owns %d replica slots" % (sys.nchains, count))
        algs = self.replica.algs
        rng0 = algs[0].rng
        if not isinstance(rng0, PhiloxRNG):
            raise ValueError("device replica exchange needs PhiloxRNG streams")
        for r, a in enumerate(algs):
            if a.rng.seed != rng0.seed or a.rng.chain != r:
                raise ValueError("replica r must carry PhiloxRNG(seed, chain=r); use philox_family(seed)")
        betas = np.array([beta_of(a) for a in algs], dtype=np.float64)
        tables = np.stack([build_table(sys.model, rule_of(a), len(sys.dims), beta_of(a), sys.J, sys.h, sys.D)
                           for a in algs])
        sys.set_rule(rule_of(algs[0]), tables)
        sys.set_rng(rng0.seed)
        h = C.c_void_p()
        check(lib().mcx_pt_create(sys.h_lat, n, first, betas.ctypes.data, C.byref(h)))
        self._pt, self._sys, self._first, self._count = h, sys, first, count
        self._betas = betas
        self._xbuf = None
        self._peers = False
        self._acc_seen = np.atleast_1d(sys._sums()[3]).astype(np.int64).copy()   # per-chain accepted already credited
        if isinstance(self.backend, GPUBackend) and self.backend.size > 1 and os.environ.get("MCX_PT_P2P", "1") != "0":
            # energies travel by peer stores over NVLink (CUDA IPC), no collective call per exchange
            buf = C.create_string_buffer(128)
            check(lib().mcx_pt_export(h, buf))
            tokens = b"".join(_all_gather_bytes(self.backend, buf.raw))
            check(lib().mcx_pt_attach_peers(h, self.backend.size, self.backend.rank, tokens))
            self._peers = True
            self.backend.barrier()
        return self

    def sweep_system_(self, sys, nsweeps):
        if self._pt is None or sys is not self._sys:
            self.attach(sys)
        # exchanging every sweep: keep the energy sums current per flip; with longer intervals it is
        # cheaper to sweep without bookkeeping and recompute the sums once, when they are published
        sys.set_tracking(int(nsweeps) < 3)
        check(lib().mcx_sweep(sys.h_lat, int(nsweeps)))
        for a in self.replica.algs[self._first:self._first + self._count]:
            a.steps += int(nsweeps) * sys.N

    def run_(self, sys, nrounds, sweeps_per_round=1):
        """`nrounds` x (`sweeps_per_round` sweeps of every replica, then update_) -- the loop of
        pt_Ising2D.jl:52-57 -- queued by the library in one call when the energies reach all ranks without
        the host (one rank, or peer stores); otherwise the same loop from here."""
        if self._pt is None or sys is not self._sys:
            self.attach(sys)
        nrounds, k = int(nrounds), int(sweeps_per_round)
        if self._peers or not (isinstance(self.backend, GPUBackend) and self.backend.size > 1):
            check(lib().mcx_pt_run(self._pt, nrounds, k))
            for a in self.replica.algs[self._first:self._first + self._count]:
                a.steps += nrounds * k * sys.N
            self._dirty = True
            return None
        for _ in range(nrounds):
            self.sweep_system_(sys, k)
            self._update_device()
        return None

    def _x_tensor(self):
        if self._xbuf is None:
            import torch
            p = C.c_void_p()
            check(lib().mcx_pt_energy_buffer(self._pt, C.byref(p)))
            self._xbuf = _as_torch(p.value, self.size, torch.float64, self._sys.ctx.device)
        return self._xbuf

    def _update_device(self):
        if self._pt is None:
            raise AssertionError("update_(rx) without energies needs rx.attach(sys) first")
        check(lib().mcx_pt_publish(self._pt))
        if isinstance(self.backend, GPUBackend) and self.backend.size > 1 and not self._peers:
            self.backend.all_gather_inplace(self._x_tensor(), self._first, self._count)
        check(lib().mcx_pt_exchange(self._pt))
        self._dirty = True
        return None

    def sync_counters(self):
        """credit the accepted moves of this rank's lattices to their replicas' algorithms (alg.accepted,
        importance_sampling.jl:80-85): the device counts them per chain; read lazily, never inside a sweep loop"""
        if self._pt is None:
            return
        now = np.atleast_1d(self._sys._sums()[3]).astype(np.int64)
        seen = np.where(now < self._acc_seen, 0, self._acc_seen)      # counters were reset in between
        for c, a in enumerate(self.replica.algs[self._first:self._first + self._count]):
            a.accepted += int(now[c] - seen[c])
        self._acc_seen = now.copy()

    def _pull(self):
        if self._pt is None or not getattr(self, "_dirty", False):
            return
        self.sync_counters()
        st, rd = C.c_int64(), C.c_int64()
        check(lib().mcx_pt_state(self._pt, self.indices.ctypes.data, self.steps.ctypes.data,
                                 self.accepted.ctypes.data, C.byref(st), C.byref(rd)))
        self.stage, self.round = st.value, rd.value
        # ensembles follow the labels (replica_exchange.jl:133)
        ens = {i + 1: type(self.replica.algs[0].ensemble)(beta=b) for i, b in enumerate(self._betas)}
        for r, a in enumerate(self.replica.algs):
            a.ensemble = ens[int(self.indices[r])]
        self._dirty = False

    def energies(self):
        """per-slot energies as last published (all ranks)."""
        self._sys.sync()
        if self._peers:        # two buffers by round parity; the last published round is round - 1
            import torch
            p = C.c_void_p()
            check(lib().mcx_pt_energy_buffer(self._pt, C.byref(p)))
            both = _as_torch(p.value, 2 * self.size, torch.float64, self._sys.ctx.device)
            self._pull_round()
            off = ((self.round - 1) & 1) * self.size
            return both[off:off + self.size].cpu().numpy().copy()
        return self._x_tensor().cpu().numpy().copy()

    def _pull_round(self):
        st, rd = C.c_int64(), C.c_int64()
        check(lib().mcx_pt_state(self._pt, None, None, None, C.byref(st), C.byref(rd)))
        self.stage, self.round = st.value, rd.value

    def peer_status(self):
        t = C.c_int32()
        check(lib().mcx_pt_peer_status(self._pt, C.byref(t)))
        return t.value

    def close(self):
        h, self._pt = self._pt, None
        if h is not None:
            try:
                lib().mcx_pt_destroy(h)
            except Exception:
                pass

    # -- serialisation (checkpoint!/restore_checkpoint, checkpointing.jl:48-101): the ladder (indices, per-edge
    # counters, stage, exchange round), the algorithms, and the lattice it is attached to
    def __getstate__(self):
        self._dirty = self._pt is not None
        self._pull()
        return {"backend_kind": type(self.backend).__name__, "algs": self.replica.algs, "stage": int(self.stage),
                "indices": self.indices.copy(), "steps": self.steps.copy(), "accepted": self.accepted.copy(),
                "round": int(self.round), "sys": self._sys if self._pt is not None else None,
                "betas": [float(b) for b in self._betas] if self._pt is not None else None}

    def __setstate__(self, st):
        algs = st["algs"]
        backend = GPUBackend() if st["backend_kind"] == "GPUBackend" else ThreadsBackend(len(algs))
        ReplicaExchange.__init__(self, backend, algs)
        self.stage, self.round = st["stage"], st["round"]
        self.indices[:] = st["indices"]
        self.steps[:] = st["steps"]
        self.accepted[:] = st["accepted"]
        if st["sys"] is not None:
            # attach() reads the ladder off the algorithms in slot order: hand it the ladder, then restore the permutation
            ens_type = type(algs[0].ensemble)
            for r, a in enumerate(algs):
                a.ensemble = ens_type(beta=st["betas"][r])
            self.attach(st["sys"])
            check(lib().mcx_pt_set_state(self._pt, self.indices.ctypes.data, self.steps.ctypes.data,
                                         self.accepted.ctypes.data, int(self.stage), int(self.round)))
            self._dirty = True
            self._pull()            # the ensembles follow the restored labels again (replica_exchange.jl:133)


class _CudaArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def _as_torch(ptr, n, dtype, device):
    """Zero-copy torch view of a device buffer owned by libmcx_b200 (plumbing for collectives)."""
    import torch
    typestr = {torch.float64: "<f8", torch.int64: "<i8", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_CudaArray(ptr, n, typestr), device="cuda:%d" % device)


def philox_family(seed):
    """`rng=` argument for ParallelTempering: maps the reference's `rng(seed + i)` call
    (parallel_tempering.jl:38,44,51) to PhiloxRNG(seed, chain=i-1), one stream per replica slot."""
    return lambda s: PhiloxRNG(seed, chain=int(s) - int(seed) - 1)


def ParallelTempering(betas, seed=1000, rng=None, backend=None):
    """parallel_tempering.jl:24-53"""
    betas = [float(b) for b in betas]
    n = len(betas)
    if n < 2:
        raise ValueError("need at least 2 replicas")
    rng = rng or philox_family(seed)
    if backend is None:
        backend = ThreadsBackend(n)
    if isinstance(backend, ThreadsBackend) and backend.size != n:
        raise ValueError("size(backend) (=%d) must equal length(betas) (=%d)" % (backend.size, n))
    algs = [Metropolis(rng(seed + i), beta=betas[i - 1]) for i in range(1, n + 1)]
    return ReplicaExchange(backend, algs)


def set_betas(nreplicas, bmin, bmax, mode="uniform"):
    """parallel_tempering.jl:146-162: range(bmax, bmin, length=n) (optionally in log space).
    Julia's float ranges are computed in twice precision; the exact rational lerp rounded once
    reproduces them for the ladders the reference tests pin (test_parallel_ensembles.jl:152-158)."""
    n = int(nreplicas)
    if n < 2:
        raise ValueError("nreplicas must be >= 2")
    if mode == "uniform":
        a, b = Fraction(float(bmax)), Fraction(float(bmin))
        return [float((a * (n - 1 - i) + b * i) / (n - 1)) for i in range(n)]
    if mode == "geometric":
        a, b = Fraction(math.log(float(bmax))), Fraction(math.log(float(bmin)))
        return [math.exp(float((a * (n - 1 - i) + b * i) / (n - 1))) for i in range(n)]
    raise ValueError("unknown beta mode %s; use :uniform or :geometric" % mode)


def update_(obj, *args, **kwargs):
    return obj.update_(*args, **kwargs)


# ---------------------------------------------------------------------------------------------------
# slab decomposition: one lattice over several GPUs (beyond the reference, whose
# IsingLatticeOptim (SpinSystems/src/ising.jl:430-461) is one Vector{Int8} in one address space)
# ---------------------------------------------------------------------------------------------------
def _all_gather_bytes(backend, raw):
    """equal-length byte strings of all ranks, in rank order (handle exchange; NCCL or gloo)"""
    import torch
    dist = backend._dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(backend.group) == "nccl" else torch.device("cpu")
    mine = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
    out = torch.empty(backend.size * len(raw), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, mine, group=backend.group)
    flat = out.cpu().numpy().tobytes()
    return [flat[i * len(raw):(i + 1) * len(raw)] for i in range(backend.size)]


class SlabIsing:
    """A 2-D Ising lattice `dims = [Lx, Ly]` (J = 1, h = 0) split by rows into equal slabs.

    With a multi-rank `GPUBackend` every rank holds one slab on its GPU; the half-sweep kernel reads the
    neighbour rows straight from the neighbour GPU's memory (CUDA IPC mapping, NVLink) and device flags
    order the half-sweeps, so `sweep_` never synchronises the host.  With a single rank, `nslabs`
    handles on one GPU advance in lockstep (what the single-GPU parity test drives).  Either way the
    trajectory is bit-identical to `Ising(dims)` on one GPU with the same PhiloxRNG."""

    def __init__(self, dims, backend=None, nslabs=None, ctx=None):
        import ctypes as C
        from ._lib import check, lib
        from .spin_systems import Ising
        self.backend = backend or GPUBackend()
        self.dims = [int(d) for d in dims]
        Lx, Ly = self.dims
        self.N = Lx * Ly
        self.remote = self.backend.size > 1
        n = self.backend.size if self.remote else int(nslabs or 1)
        if Ly % (2 * n) != 0:
            raise ValueError("Ly = %d does not split into %d slabs of an even number of rows" % (Ly, n))
        self.nslabs, self.rows = n, Ly // n
        mine = [self.backend.rank] if self.remote else list(range(n))
        self.parts = [Ising([Lx, self.rows], ctx=ctx) for _ in mine]
        for p, k in zip(self.parts, mine):
            check(lib().mcx_slab_configure(p.h_lat, Ly, k * self.rows))
        if self.remote:
            buf = C.create_string_buffer(128)
            check(lib().mcx_slab_export(self.parts[0].h_lat, buf))
            handles = self._all_gather_bytes(buf.raw)
            r = self.backend.rank
            up, dn = handles[(r - 1) % n], handles[(r + 1) % n]
            check(lib().mcx_slab_attach_ipc(self.parts[0].h_lat, up, dn))
            self.backend.barrier()
        else:
            for k, p in enumerate(self.parts):
                check(lib().mcx_slab_attach_local(p.h_lat, self.parts[(k - 1) % n].h_lat, self.parts[(k + 1) % n].h_lat))

    def _all_gather_bytes(self, raw):
        return _all_gather_bytes(self.backend, raw)

    def _sync_all(self):
        for p in self.parts:
            p.sync()
        self.backend.barrier()

    def _recompute(self):
        from ._lib import check, lib
        self._sync_all()                       # every slab's rows are final before anyone reads halo rows
        for p in self.parts:
            check(lib().mcx_recompute(p.h_lat))
        self._sync_all()

    def init_(self, type, rng=None):
        for p in self.parts:
            p.init_(type, rng=rng)
        self._recompute()
        return self

    def set_tracking(self, on):
        self._tracking = bool(on)
        for p in self.parts:
            p.set_tracking(on)

    def _fresh_sums(self):
        """untracked sums are recomputed from the spins, which reads halo rows: only between sweeps of ALL slabs"""
        if not getattr(self, "_tracking", True):
            self._recompute()

    def sweep_(self, alg, nsweeps=1):
        from ._lib import check, lib
        for p in self.parts:
            p._bind_alg(alg)
        self._fresh_sums()
        before = sum(int(p._sums()[3].sum()) for p in self.parts)
        if self.remote:
            check(lib().mcx_sweep(self.parts[0].h_lat, int(nsweeps)))
        else:
            for _ in range(2 * int(nsweeps)):
                for p in self.parts:
                    check(lib().mcx_slab_half_sweep(p.h_lat))
        alg.steps += int(nsweeps) * self.N
        if hasattr(alg, "accepted"):
            self._fresh_sums()
            alg.accepted += self._reduce(sum(int(p._sums()[3].sum()) for p in self.parts) - before)

    def _reduce(self, value):
        if not self.remote:
            return int(value)
        import torch
        dist = self.backend._dist
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.backend.group) == "nccl" else torch.device("cpu")
        t = torch.tensor([int(value)], dtype=torch.int64, device=dev)
        self.backend.all_reduce_sum(t)
        return int(t.item())

    def status(self):
        """(timed_out, half_sweeps_done) of this rank's slab(s); synchronises."""
        import ctypes as C
        from ._lib import check, lib
        out = []
        for p in self.parts:
            t, e = C.c_int32(), C.c_uint64()
            check(lib().mcx_slab_status(p.h_lat, C.byref(t), C.byref(e)))
            out.append((t.value, e.value))
        return out

    def pair_sum(self):
        self._fresh_sums()
        return self._reduce(sum(int(p._sums()[0].sum()) for p in self.parts))

    def magnetization(self, full=False):
        if full:
            self._recompute()
        self._fresh_sums()
        return self._reduce(sum(int(p._sums()[1].sum()) for p in self.parts))

    def energy(self, full=False):
        if full:
            self._recompute()
        return -self.pair_sum()

    @property
    def local_spins(self):
        """this rank's rows (reference site order inside the slab)"""
        import numpy as np
        return np.concatenate([p.spins.reshape(-1) for p in self.parts])

    @property
    def spins(self):
        """the whole lattice on every rank (gathers over ranks)"""
        import numpy as np
        loc = self.local_spins
        if not self.remote:
            return loc
        parts = self._all_gather_bytes(loc.tobytes())
        return np.concatenate([np.frombuffer(b, dtype=np.int8) for b in parts])

    def close(self):
        for p in self.parts:
            p.sync()
