"""ctypes binding of the CPU ORACLE (oracle/mcx_oracle.c).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.  The product
package (montecarlox.jl_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmcx_oracle.so")

ISING, BLUME_CAPEL = 0, 1
METROPOLIS, GLAUBER, HEATBATH = 0, 1, 2
TAG_SWEEP, TAG_EXCHANGE, TAG_INIT, TAG_FLAT = 0, 1, 2, 3


def build(force=False):
    src = os.path.join(_HERE, "mcx_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class _Rng(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("chain", C.c_uint32), ("tag", C.c_uint32),
                ("t", C.c_uint64), ("q", C.c_uint64), ("draw", C.c_uint32)]


class _Alg(C.Structure):
    _fields_ = [("rule", C.c_int), ("beta", C.c_double), ("steps", C.c_int64), ("accepted", C.c_int64)]


class _Flat(C.Structure):
    _fields_ = [("start", C.c_int64), ("step", C.c_int64), ("num", C.c_int64),
                ("logweight", C.POINTER(C.c_double)), ("histogram", C.POINTER(C.c_double)),
                ("logf", C.c_double)]


class _Xo(C.Structure):
    _fields_ = [("s", C.c_uint64 * 4)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    vp, i64, u64, u32, dbl, ci = C.c_void_p, C.c_int64, C.c_uint64, C.c_uint32, C.c_double, C.c_int
    pd, pi64 = C.POINTER(C.c_double), C.POINTER(C.c_int64)
    L.mcxo_philox4x32_10.argtypes = [C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
    L.mcxo_rng_position.argtypes = [C.POINTER(_Rng), u32, u64, u64]
    L.mcxo_rng_lane16.argtypes = [C.POINTER(_Rng), u32]
    L.mcxo_rng_lane16.restype = u32
    L.mcxo_rand_f64.argtypes = [C.POINTER(_Rng)]
    L.mcxo_rand_f64.restype = dbl
    L.mcxo_rand_bool.argtypes = [C.POINTER(_Rng)]
    L.mcxo_rand_bool.restype = ci
    L.mcxo_exchange_u.argtypes = [u64, u32, u64]
    L.mcxo_exchange_u.restype = dbl
    L.mcxo_system_create.argtypes = [ci, ci, pi64, dbl, dbl, dbl]
    L.mcxo_system_create.restype = vp
    L.mcxo_system_destroy.argtypes = [vp]
    L.mcxo_system_set_spins.argtypes = [vp, vp]
    L.mcxo_system_get_spins.argtypes = [vp, vp]
    L.mcxo_system_init_random.argtypes = [vp, u64, u32]
    L.mcxo_recompute.argtypes = [vp]
    L.mcxo_energy.argtypes = [vp, ci]
    L.mcxo_energy.restype = dbl
    L.mcxo_magnetization.argtypes = [vp, ci]
    L.mcxo_magnetization.restype = i64
    L.mcxo_pair_count.argtypes = [vp]
    L.mcxo_pair_count.restype = i64
    L.mcxo_spin2_sum.argtypes = [vp]
    L.mcxo_spin2_sum.restype = i64
    L.mcxo_local_pair_interactions.argtypes = [vp, i64]
    L.mcxo_local_pair_interactions.restype = i64
    L.mcxo_delta_energy_flip.argtypes = [vp, i64]
    L.mcxo_delta_energy_flip.restype = dbl
    L.mcxo_delta_energy_bc.argtypes = [vp, i64, ci]
    L.mcxo_delta_energy_bc.restype = dbl
    L.mcxo_logistic.argtypes = [dbl]
    L.mcxo_logistic.restype = dbl
    L.mcxo_propose_state.argtypes = [ci, ci]
    L.mcxo_propose_state.restype = ci
    L.mcxo_attempt_at.argtypes = [vp, C.POINTER(_Alg), i64, C.POINTER(_Rng)]
    L.mcxo_sweep_checkerboard.argtypes = [vp, C.POINTER(_Alg), u64, u32, u64, i64]
    L.mcxo_sweep_random_site.argtypes = [vp, C.POINTER(_Alg), C.POINTER(_Xo), i64, ci]
    L.mcxo_xoshiro_seed.argtypes = [C.POINTER(_Xo), u64]
    L.mcxo_baseline_random_site.argtypes = [ci, dbl, ci, i64, ci, ci, u64, pd, pd]
    L.mcxo_baseline_random_site.restype = dbl
    L.mcxo_stats_random_site.argtypes = [ci, dbl, ci, i64, i64, i64, ci, u64, pd]
    L.mcxo_baseline_lean.argtypes = [ci, dbl, ci, i64, ci, u64, pd]
    L.mcxo_baseline_lean.restype = dbl
    L.mcxo_table_len.argtypes = [ci, ci, ci]
    L.mcxo_table_len.restype = ci
    L.mcxo_build_table.argtypes = [ci, ci, ci, dbl, dbl, dbl, dbl, C.POINTER(u64)]
    L.mcxo_binindex.argtypes = [i64, i64, i64]
    L.mcxo_binindex.restype = i64
    L.mcxo_binindex_f.argtypes = [dbl, dbl, dbl]
    L.mcxo_binindex_f.restype = i64
    L.mcxo_muca_update.argtypes = [pd, pd, i64]
    L.mcxo_flat_sweep.argtypes = [vp, C.POINTER(_Alg), C.POINTER(_Flat), ci, ci, dbl, u64, u32, u64, i64]
    L.mcxo_flat_sweep.restype = ci
    L.mcxo_flat_sweep_policy.argtypes = [vp, C.POINTER(_Alg), C.POINTER(_Flat), ci, ci, dbl, u64, u32, u64, i64, ci]
    L.mcxo_flat_sweep_policy.restype = ci
    L.mcxo_flat_accept.argtypes = [C.POINTER(_Alg), C.POINTER(_Flat), ci, i64, i64, dbl]
    L.mcxo_flat_accept.restype = ci
    L.mcxo_exchange_log_ratio.argtypes = [dbl, dbl, dbl, dbl]
    L.mcxo_exchange_log_ratio.restype = dbl
    L.mcxo_accept_exchange.argtypes = [dbl, dbl]
    L.mcxo_accept_exchange.restype = ci
    L.mcxo_resolve_pair.argtypes = [i64, i64, i64, pi64]
    L.mcxo_set_betas.argtypes = [i64, dbl, dbl, ci, pd]
    L.mcxo_rx_update.argtypes = [i64, pi64, pi64, pi64, pi64, pd, pd, pd]
    # general topologies
    L.mcxo_graph_create.argtypes = [i64, vp, vp, vp, dbl, ci, dbl, vp]
    L.mcxo_graph_create.restype = vp
    L.mcxo_graph_destroy.argtypes = [vp]
    L.mcxo_graph_set_spins.argtypes = [vp, vp]
    L.mcxo_graph_get_spins.argtypes = [vp, vp]
    L.mcxo_graph_init_random.argtypes = [vp, u64, u32]
    L.mcxo_graph_recompute.argtypes = [vp]
    L.mcxo_graph_local_pair.argtypes = [vp, i64]
    L.mcxo_graph_local_pair.restype = dbl
    L.mcxo_graph_energy.argtypes = [vp, ci]
    L.mcxo_graph_energy.restype = dbl
    L.mcxo_graph_magnetization.argtypes = [vp]
    L.mcxo_graph_magnetization.restype = i64
    L.mcxo_graph_delta_energy.argtypes = [vp, i64]
    L.mcxo_graph_delta_energy.restype = dbl
    L.mcxo_graph_flip.argtypes = [vp, i64]
    L.mcxo_graph_attempt_at.argtypes = [vp, C.POINTER(_Alg), i64, C.POINTER(_Rng)]
    L.mcxo_graph_colour.argtypes = [vp, vp]
    L.mcxo_graph_colour.restype = ci
    L.mcxo_graph_sweep_coloured.argtypes = [vp, C.POINTER(_Alg), u64, u32, u64, i64, vp, ci]
    _lib = L
    return L


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().mcxo_philox4x32_10(c, k, o)
    return [int(v) for v in o]


class Rng:
    """Positioned Philox stream: what the Julia shim's PhiloxRNG <: AbstractRNG is."""

    def __init__(self, seed, chain=0):
        self.r = _Rng(seed, chain, 0, 0, 0, 0)

    def position(self, tag, t, q):
        lib().mcxo_rng_position(C.byref(self.r), tag, t, q)
        return self

    def rand(self):
        return lib().mcxo_rand_f64(C.byref(self.r))

    def rand_bool(self):
        return bool(lib().mcxo_rand_bool(C.byref(self.r)))

    def lane16(self, plane):
        return int(lib().mcxo_rng_lane16(C.byref(self.r), plane))


class Alg:
    def __init__(self, rule, beta):
        self.a = _Alg(rule, beta, 0, 0)

    @property
    def steps(self):
        return self.a.steps

    @property
    def accepted(self):
        return self.a.accepted

    @property
    def beta(self):
        return self.a.beta

    def reset(self):
        self.a.steps = 0
        self.a.accepted = 0


class System:
    """Ising / Blume-Capel on a periodic grid (restated IsingLatticeOptim / graph systems)."""

    def __init__(self, model, dims, J=1.0, h=0.0, D=0.0):
        self.model, self.dims = model, list(dims)
        d = (C.c_int64 * 3)(*(list(dims) + [1] * (3 - len(dims))))
        self.p = lib().mcxo_system_create(model, len(dims), d, J, h, D)
        self.N = int(np.prod(dims))

    def __del__(self):
        try:
            lib().mcxo_system_destroy(self.p)
        except Exception:
            pass

    @property
    def spins(self):
        out = np.empty(self.N, dtype=np.int8)
        lib().mcxo_system_get_spins(self.p, out.ctypes.data)
        return out

    @spins.setter
    def spins(self, v):
        v = np.ascontiguousarray(v, dtype=np.int8).reshape(-1)
        assert v.size == self.N
        lib().mcxo_system_set_spins(self.p, v.ctypes.data)

    def init_random(self, seed, chain=0):
        lib().mcxo_system_init_random(self.p, seed, chain)

    def energy(self, full=False):
        return lib().mcxo_energy(self.p, int(full))

    def magnetization(self, full=False):
        return int(lib().mcxo_magnetization(self.p, int(full)))

    def pair_count(self):
        return int(lib().mcxo_pair_count(self.p))

    def spin2_sum(self):
        return int(lib().mcxo_spin2_sum(self.p))

    def local_pair_interactions(self, i):
        return int(lib().mcxo_local_pair_interactions(self.p, i))

    def delta_energy(self, i, s_new=None):
        if self.model == ISING:
            return lib().mcxo_delta_energy_flip(self.p, i)
        return lib().mcxo_delta_energy_bc(self.p, i, s_new)

    def attempt_at(self, alg, i, rng):
        lib().mcxo_attempt_at(self.p, C.byref(alg.a), i, C.byref(rng.r))

    def sweep_checkerboard(self, alg, seed, chain, sweep0, nsweeps):
        lib().mcxo_sweep_checkerboard(self.p, C.byref(alg.a), seed, chain, sweep0, nsweeps)

    def sweep_random_site(self, alg, xo, nattempts, use_table=False):
        lib().mcxo_sweep_random_site(self.p, C.byref(alg.a), C.byref(xo), nattempts, int(use_table))

    def flat_sweep(self, alg, flat, kind, observable, beta_pair, seed, chain, sweep0, nsweeps, policy=0):
        return lib().mcxo_flat_sweep_policy(self.p, C.byref(alg.a), C.byref(flat.f), kind, observable, beta_pair,
                                            seed, chain, sweep0, nsweeps, policy)


def xoshiro(seed):
    x = _Xo()
    lib().mcxo_xoshiro_seed(C.byref(x), seed)
    return x


class Flat:
    """Discrete 1-D BinnedObject pair (logweight, histogram) + logf."""

    def __init__(self, start, step, num, logf=1.0):
        self.logweight = np.zeros(num, dtype=np.float64)
        self.histogram = np.zeros(num, dtype=np.float64)
        self.f = _Flat(start, step, num, self.logweight.ctypes.data_as(C.POINTER(C.c_double)),
                       self.histogram.ctypes.data_as(C.POINTER(C.c_double)), logf)

    def accept(self, alg, kind, x_new, x_old, u):
        return lib().mcxo_flat_accept(C.byref(alg.a), C.byref(self.f), kind, x_new, x_old, u)

    def muca_update(self):
        pd = C.POINTER(C.c_double)
        lib().mcxo_muca_update(self.logweight.ctypes.data_as(pd), self.histogram.ctypes.data_as(pd),
                               self.logweight.size)


def build_table(model, rule, ndim, beta, J=1.0, h=0.0, D=0.0):
    n = lib().mcxo_table_len(model, rule, ndim)
    T = (C.c_uint64 * n)()
    lib().mcxo_build_table(model, rule, ndim, beta, J, h, D, T)
    return np.array(list(T), dtype=np.uint64)


def resolve_pair(my_index, stage, nranks):
    out = (C.c_int64 * 3)()
    lib().mcxo_resolve_pair(my_index, stage, nranks, out)
    return bool(out[0]), int(out[1]), int(out[2])


def set_betas(n, bmin, bmax, mode="uniform"):
    out = np.empty(n, dtype=np.float64)
    lib().mcxo_set_betas(n, bmin, bmax, int(mode == "geometric"), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def rx_update(stage, indices, steps, accepted, beta_of_slot, xs, u_of_slot):
    """In-place update!(rx, xs); returns new stage."""
    n = len(indices)
    st = C.c_int64(stage)
    pi, pd = C.POINTER(C.c_int64), C.POINTER(C.c_double)
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    us = np.ascontiguousarray(u_of_slot, dtype=np.float64)
    lib().mcxo_rx_update(n, C.byref(st), indices.ctypes.data_as(pi), steps.ctypes.data_as(pi),
                         accepted.ctypes.data_as(pi), beta_of_slot.ctypes.data_as(pd),
                         xs.ctypes.data_as(pd), us.ctypes.data_as(pd))
    return int(st.value)


def baseline_random_site(L, beta, nchains, sweeps, nthreads, use_table=False, seed=42):
    m, e = C.c_double(), C.c_double()
    secs = lib().mcxo_baseline_random_site(L, beta, nchains, sweeps, nthreads, int(use_table), seed,
                                           C.byref(m), C.byref(e))
    return secs, m.value, e.value


def baseline_lean(L, beta, nchains, nattempts, use_table=False, seed=42):
    """(seconds, acceptance rate) of the reference's random-site loop, one chain per thread."""
    r = C.c_double()
    secs = lib().mcxo_baseline_lean(L, beta, nchains, nattempts, int(use_table), seed, C.byref(r))
    return secs, r.value


def stats_random_site(L, beta, nchains, therm, sweeps, interval, nthreads=8, seed=42):
    """[nchains, 4] per-chain averages of (e, |m|, m^2, m^4) from the reference's random-site loop."""
    out = np.zeros((nchains, 4), dtype=np.float64)
    lib().mcxo_stats_random_site(L, beta, nchains, therm, sweeps, interval, nthreads, seed,
                                 out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


class Graph:
    """IsingGraph (val=None: global J) / IsingMatrix (val: J_ij per CSR entry) with h = 0, a scalar or a vector
    (SpinSystems/src/ising.jl:86-360).  rowptr / col: 0-based CSR neighbour lists in ascending neighbour order."""

    def __init__(self, rowptr, col, val=None, J=1.0, h=0):
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        self.col = np.ascontiguousarray(col, dtype=np.int64)
        self.val = None if val is None else np.ascontiguousarray(val, dtype=np.float64)
        self.n = len(self.rowptr) - 1
        if np.ndim(h) == 0:
            hmode, hs, hv = (0 if h == 0 else 1), float(h), None
        else:
            hv = np.ascontiguousarray(h, dtype=np.float64)
            hmode, hs = 2, 0.0
        self.hvec = hv
        self.p = lib().mcxo_graph_create(self.n, self.rowptr.ctypes.data, self.col.ctypes.data,
                                         None if self.val is None else self.val.ctypes.data, float(J), hmode, hs,
                                         None if hv is None else hv.ctypes.data)

    def __del__(self):
        try:
            lib().mcxo_graph_destroy(self.p)
        except Exception:
            pass

    @property
    def spins(self):
        out = np.empty(self.n, dtype=np.int8)
        lib().mcxo_graph_get_spins(self.p, out.ctypes.data)
        return out

    @spins.setter
    def spins(self, v):
        v = np.ascontiguousarray(v, dtype=np.int8)
        lib().mcxo_graph_set_spins(self.p, v.ctypes.data)

    def init_random(self, seed, chain=0):
        lib().mcxo_graph_init_random(self.p, seed, chain)

    def energy(self, full=False):
        return lib().mcxo_graph_energy(self.p, int(full))

    def magnetization(self):
        return lib().mcxo_graph_magnetization(self.p)

    def local_pair_interactions(self, i):
        return lib().mcxo_graph_local_pair(self.p, i)

    def delta_energy(self, i):
        return lib().mcxo_graph_delta_energy(self.p, i)

    def flip(self, i):
        lib().mcxo_graph_flip(self.p, i)

    def colour(self):
        c = np.empty(self.n, dtype=np.int32)
        n = lib().mcxo_graph_colour(self.p, c.ctypes.data)
        return c, n

    def sweep_coloured(self, alg, seed, chain, sweep0, nsweeps, colour=None, ncolours=None):
        if colour is None:
            colour, ncolours = self.colour()
        colour = np.ascontiguousarray(colour, dtype=np.int32)
        lib().mcxo_graph_sweep_coloured(self.p, C.byref(alg.a), seed, chain, sweep0, nsweeps, colour.ctypes.data, int(ncolours))


def grid_csr(dims, periodic=True):
    """CSR neighbour lists of Graphs.SimpleGraphs.grid(dims; periodic): site i = x + Lx*(y + Ly*z), neighbours ascending,
    duplicate edges of length-2 periodic dimensions merged as a SimpleGraph does"""
    dims = [int(d) for d in dims]
    n = int(np.prod(dims))
    strides = np.cumprod([1] + dims[:-1])
    rows = [set() for _ in range(n)]
    for i in range(n):
        for ax, (L, st) in enumerate(zip(dims, strides)):
            x = (i // st) % L
            for dx in (-1, 1):
                y = x + dx
                if periodic:
                    y %= L
                elif y < 0 or y >= L:
                    continue
                j = i + (y - x) * st
                if j != i:
                    rows[i].add(j)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    col = []
    for i, r in enumerate(rows):
        col.extend(sorted(r))
        rowptr[i + 1] = len(col)
    return rowptr, np.array(col, dtype=np.int64)
