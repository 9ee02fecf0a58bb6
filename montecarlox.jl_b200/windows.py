"""Wang-Landau density of states with the energy range split into overlapping windows
(BASELINE.json configs[4]: "3D Ising L=256 WangLandau density-of-states, energy windows sharded over
8 B200"; SURVEY.md section 8e, row "C5 WL windows").

The reference has Wang-Landau (`WangLandau(rng, bins; logf)` algorithms/wang_landau.jl:10-18, its
`accept!` :29-37, `update!(ens; power)` ensembles/wang_landau.jl:23) but no energy windows: a lookup
outside the binned range is a `BoundsError` (test/test_multicanonical.jl:39-43).  Everything a window
adds is therefore defined here, on top of the reference's per-attempt rule:

* a window is a contiguous block of the bins of ONE global `BinnedObject(start:step:stop)`; adjacent
  windows overlap (`overlap` = shared fraction of a window);
* inside a window a walker is the reference's Wang-Landau chain; a proposal that would leave the
  window is a rejected attempt that visits the current bin (`out_of_range_policy = 1` of
  `mcx_flat_create`, include/mcx_b200.h);
* `walkers` chains per window keep one table each (one `WangLandauEnsemble` per algorithm object);
  `update_` averages them and halves `logf`, like `update!` on every ensemble;
* windows are independent units: they are dealt to the ranks in contiguous blocks with NO data-path
  collective while sampling; `logdos()` all-gathers the window pieces (a few KB to MB, once) and every
  rank joins them at the overlaps on the host;
* optionally (`exchange_`) walkers of neighbouring windows swap configurations -- the reference's replica
  exchange (`exchange_log_ratio` / `_accept_exchange`, replica_exchange.jl:110-115, even / odd pair stages
  :158-178) with the two windows' Wang-Landau tables as the two ensembles, attempted only when both energies
  lie in both windows; across ranks only the walkers' energies, two table differences per walker pair and the
  accepted configurations move (point-to-point), and `u` is the counter-based EXCHANGE stream, so every rank
  takes the same decisions;
* random streams are keyed by the GLOBAL walker number (window * walkers + walker), so the joined result
  does not depend on how many ranks the windows were dealt to.

Before a window can be sampled its walkers have to sit inside it.  `prepare_` drives them there with the
canonical checkerboard sweep (Glauber, the same kernels as `sweep_`): starting from the ground state on
the window's side of E = 0, heat at |beta| found by bisection and keep the first configuration whose
energy lies in the window.

All sampling goes through libmcx_b200 (C ABI); the windows of one rank live on separate contexts
(streams), so their serial chains run concurrently on the GPU."""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import McxError, check, lib
from .binned_object import BinnedObject
from .parallel import GPUBackend, _accept_exchange, partition_slots
from .rng import exchange_u
from .tables import build_table


# --------------------------------------------------------------------------- window geometry
def partition_windows(nbins, nwindows, overlap=0.5):
    """[(first_bin, width)] * nwindows: equal-width windows covering bins 0..nbins-1, adjacent windows
    sharing about `overlap` of a window (0 < overlap < 1).  0-based bin indices."""
    nbins, nwindows = int(nbins), int(nwindows)
    if nwindows < 1:
        raise ValueError("nwindows must be >= 1")
    if nwindows == 1:
        return [(0, nbins)]
    if not (0.0 < overlap < 1.0):
        raise ValueError("overlap must lie in (0, 1)")
    width = int(math.ceil(nbins / (1.0 + (nwindows - 1) * (1.0 - overlap))))
    width = min(max(width, 2), nbins)
    stride = (nbins - width) / (nwindows - 1)
    firsts = [int(round(k * stride)) for k in range(nwindows)]
    firsts[-1] = nbins - width
    for a, b in zip(firsts[:-1], firsts[1:]):
        if not (a < b < a + width):
            raise ValueError("%d bins cannot be split into %d overlapping windows of %d bins"
                             % (nbins, nwindows, width))
    return [(f, width) for f in firsts]


def join_logdos(pieces, firsts, nbins):
    """Join window estimates of log g into one table of `nbins` entries.

    pieces[k][i] estimates log g at global bin firsts[k] + i up to a per-window constant; an entry that is
    exactly 0 was never visited (a Wang-Landau visit always adds logf > 0).  Window k is shifted by the mean
    difference over the bins it shares with what is already joined, and cross-faded linearly over them.
    Unvisited bins come out as NaN (cf. the NaN rows of logdos_exact_ising2D, ising2d_exact.jl:24-38)."""
    out = np.full(int(nbins), np.nan, dtype=np.float64)
    prev_end = 0
    for k, (piece, a) in enumerate(zip(pieces, firsts)):
        piece = np.asarray(piece, dtype=np.float64)
        a = int(a)
        e = a + piece.size
        visited = piece != 0.0
        if k == 0:
            out[a:e][visited] = piece[visited]
            prev_end = e
            continue
        seg = out[a:e]
        both = np.isfinite(seg) & visited
        if not both.any():
            raise ValueError("windows %d and %d share no visited bin: more sweeps or more overlap" % (k - 1, k))
        shift = float(np.mean(seg[both] - piece[both]))
        span = max(min(prev_end, e) - a, 1)
        t = np.clip((np.arange(piece.size) + 0.5) / span, 0.0, 1.0)
        new = piece + shift
        seg[both] = (1.0 - t[both]) * seg[both] + t[both] * new[both]
        only = visited & ~both
        seg[only] = new[only]
        prev_end = max(prev_end, e)
    return out


# --------------------------------------------------------------------------- one window on the device
class DeviceWindow:
    """The walkers of one window: a batched Ising lattice (one chain per walker) on its own context.

    The driver below only uses the methods of this class, which is what lets the test-suite run the same
    driver against a CPU restatement and compare every table bit for bit."""

    def __init__(self, dims, nwalkers, seed, first_chain, device=0):
        from .spin_systems import Context, Ising
        self.dims, self.k = [int(d) for d in dims], int(nwalkers)
        self.N = int(np.prod(self.dims))
        self.ctx = Context(device)              # own stream: windows of one rank overlap on the GPU
        self.sys = Ising(self.dims, nchains=self.k, ctx=self.ctx)
        check(lib().mcx_lattice_set_first_chain_id(self.sys.h_lat, int(first_chain)))
        self.sys._first_chain = int(first_chain)
        self.sys_seed = int(seed)
        self.sys.set_rng(self.sys_seed, 0)
        self.h_flat = None
        self.nbins = 0

    # ---- configurations and energies
    def set_spins(self, spins):
        self.sys.spins = np.ascontiguousarray(spins, dtype=np.int8).reshape(-1)

    def spins(self):
        return np.asarray(self.sys.spins).reshape(self.k, self.N)

    def energies(self):
        """E = -sum_pair_interactions per walker (ising.jl:175-177 with J = 1, h = 0), int64[k]"""
        return -self.sys._sums()[0]

    # ---- position of the random streams: one sweep counter per lattice, shared by both kinds of sweep
    def get_sweep(self):
        return int(self.sys.sweep_index)

    def set_sweep(self, n):
        self.sys.set_rng(self.sys_seed, int(n))

    # ---- canonical checkerboard sweeps (the drive into the window)
    def canonical_(self, rule, beta, nsweeps):
        T = build_table(_lib.ISING, rule, len(self.dims), float(beta), 1, 0, 0)
        self.sys.set_rule(rule, T)
        self.sys.set_labels(np.zeros(self.k, dtype=np.int32))
        check(lib().mcx_sweep(self.sys.h_lat, int(nsweeps)))

    # ---- the window's Wang-Landau tables
    def open_window(self, start, step, nbins):
        self.close_window()
        h = C.c_void_p()
        check(lib().mcx_flat_create(self.sys.h_lat, _lib.FLAT_WANG_LANDAU, _lib.OBS_ENERGY, int(start), int(step),
                                    int(nbins), 0.0, 1, C.byref(h)))
        self.h_flat, self.nbins = h, int(nbins)

    def wl_sweep_(self, nsweeps, logf):
        """asynchronous: returns once the sweeps are queued on the window's stream"""
        check(lib().mcx_flat_set_logf(self.h_flat, float(logf)))
        check(lib().mcx_flat_sweep(self.h_flat, int(nsweeps)))

    def logweight(self):
        """[k, nbins] Float64 tables; waits for the queued sweeps (IndexError if a walker sits outside)"""
        lw = np.empty((self.k, self.nbins), dtype=np.float64)
        check(lib().mcx_flat_get_logweight(self.h_flat, lw.ctypes.data))
        return lw

    def set_logweight(self, lw):
        lw = np.ascontiguousarray(lw, dtype=np.float64).reshape(self.k, self.nbins)
        check(lib().mcx_flat_set_logweight(self.h_flat, lw.ctypes.data))

    def close_window(self):
        h, self.h_flat = self.h_flat, None
        if h is not None:
            check(lib().mcx_flat_destroy(h))

    def close(self):
        self.close_window()
        sys_, self.sys = self.sys, None
        if sys_ is not None:
            h_lat, sys_.h_lat = sys_.h_lat, None
            check(lib().mcx_lattice_destroy(h_lat))
        ctx, self.ctx = self.ctx, None
        if ctx is not None:
            ctx.close()

    def __del__(self):
        try:                      # tables before the lattice, the lattice before its context
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------------- the driver
class WangLandauWindows:
    """Windowed Wang-Landau for periodic Ising lattices (J = 1, h = 0; 2-D or 3-D, every dimension even).

        wl = WangLandauWindows([L, L, L], nwindows=8, walkers=4, backend=GPUBackend())
        wl.prepare_()                       # walkers into their windows
        while wl.logf > 1e-6:
            wl.sweep_(1000)                 # every walker: 1000 * N Wang-Landau attempts
            wl.update_()                    # update!(ens): logf *= 0.5
        g = wl.logdos()                     # BinnedObject over the full energy range, on every rank
    """

    def __init__(self, dims, nwindows, walkers=1, overlap=0.5, seed=42, logf=1.0, bins=None, backend=None,
                 device=None, window_factory=None, windows=None, replicate_seed=False):
        self.dims = [int(d) for d in dims]
        if len(self.dims) not in (2, 3) or any(d % 2 or d < 4 for d in self.dims):
            raise ValueError("windowed Wang-Landau needs a 2-D or 3-D lattice with even dimensions >= 4")
        self.N = int(np.prod(self.dims))
        d = len(self.dims)
        if bins is None:
            bins = range(-d * self.N, d * self.N + 1, 4)          # every energy of the periodic J = 1 lattice
        if not isinstance(bins, range) or bins.step <= 0 or len(bins) < 1:
            raise ValueError("bins must be an integer range start:step:stop (DiscreteBinning, binned_object.jl:13-24)")
        self.bins = bins
        self.nwindows, self.walkers = int(nwindows), int(walkers)
        if self.walkers < 1:
            raise ValueError("walkers must be >= 1")
        self.windows = [tuple(w) for w in windows] if windows is not None else partition_windows(len(bins), self.nwindows, overlap)
        if len(self.windows) != self.nwindows or len({n for _, n in self.windows}) != 1:
            raise ValueError("need nwindows windows of equal width")
        self.width = self.windows[0][1]
        self.seed, self.logf = int(seed), float(logf)
        if not (self.logf > 0):
            raise ValueError("logf must be > 0")
        self.backend = backend if backend is not None else GPUBackend()
        self.first, self.count = partition_slots(self.nwindows, self.backend.size, self.backend.rank)
        if device is None:
            import os
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = int(device)
        make = window_factory or (lambda dims_, k, seed_, first_chain: DeviceWindow(dims_, k, seed_, first_chain,
                                                                                   self.device))
        # global walker number = window * walkers + walker: the random streams do not depend on the rank count
        self.local = [make(self.dims, self.walkers, self.seed, w * self.walkers)
                      for w in range(self.first, self.first + self.count)]
        # replicate_seed: drive ONE walker per window into it and start all the window's walkers from that
        # configuration (their random streams differ, so they separate at once); the drive then costs one walker per
        # window instead of `walkers` -- what makes thousands of walkers per GPU practical
        self._make, self.replicate_seed = make, bool(replicate_seed)
        self._lw = [None] * self.count          # last tables read back, [walkers, width] per local window
        self._lw_stage = [None] * self.count    # tables at the start of the current logf stage
        self.steps = 0
        self.prepared = False
        # neighbour-window exchange: stage parity and round as in ReplicaExchange (replica_exchange.jl:13-19);
        # attempts / acceptances per window pair (w, w + 1), counted on the rank that owns window w
        self.exchange_stage, self.exchange_round = 0, 0
        self.exchange_steps = np.zeros(max(self.nwindows - 1, 0), dtype=np.int64)
        self.exchange_accepted = np.zeros(max(self.nwindows - 1, 0), dtype=np.int64)
        # drive parameters
        self.beta_max, self.drive_sweeps, self.drive_check, self.drive_trials = 2.0, 24, 1, 48

    # ---- geometry helpers
    def window_energies(self, w):
        """(E_lo, E_hi) of global window w, both inside"""
        f, n = self.windows[w]
        return self.bins.start + f * self.bins.step, self.bins.start + (f + n - 1) * self.bins.step

    def _ground_state(self, antiferro):
        if not antiferro:
            return np.ones(self.N, dtype=np.int8)
        par = np.zeros(self.dims[::-1], dtype=np.int8)           # site i = x + Lx*(y + Ly*z) (ising.jl:444-456)
        for ax, n in enumerate(self.dims[::-1]):
            shape = [1] * len(self.dims)
            shape[ax] = n
            par = par ^ (np.arange(n, dtype=np.int8) & 1).reshape(shape)
        return (1 - 2 * par).astype(np.int8).reshape(-1)

    # ---- seeding
    def prepare_(self):
        """Put every walker inside its window and open the window's tables."""
        for j, eng in enumerate(self.local):
            w = self.first + j
            lo, hi = self.window_energies(w)
            if self.replicate_seed and self.walkers > 1:
                one = self._make(self.dims, 1, self.seed, w * self.walkers)
                self._drive(one, lo, hi)
                first = np.asarray(one.spins()).reshape(1, self.N)
                one.close()
                eng.set_spins(np.repeat(first, self.walkers, axis=0))
            else:
                self._drive(eng, lo, hi)
            eng.open_window(lo, self.bins.step, self.width)
            self._lw[j] = np.zeros((self.walkers, self.width))
            self._lw_stage[j] = self._lw[j].copy()
        self.prepared = True
        return self

    def _drive(self, eng, lo, hi):
        k = eng.k
        centre = 0.5 * (lo + hi)
        sign = 1.0 if centre <= 0 else -1.0                       # E > 0 is reached with beta < 0
        ground = np.tile(self._ground_state(sign < 0), (k, 1))
        snap = [None] * k

        def take():
            E = np.asarray(eng.energies())
            inside = [c for c in range(k) if snap[c] is None and lo <= E[c] <= hi]
            if inside:
                s = eng.spins()
                for c in inside:
                    snap[c] = s[c].copy()
            return E

        b_lo, b_hi = 0.0, self.beta_max
        for _ in range(self.drive_trials):
            eng.set_spins(ground)
            beta = 0.5 * (b_lo + b_hi)
            E = take()
            t = 0
            while t < self.drive_sweeps and any(s is None for s in snap):
                eng.canonical_(_lib.GLAUBER, sign * beta, self.drive_check)
                t += self.drive_check
                E = take()
            if all(s is not None for s in snap):
                break
            # too hot (too far from the ground state): larger |beta|; too cold: smaller
            if float(np.mean(sign * E)) > sign * centre:
                b_lo = beta
            else:
                b_hi = beta
        else:
            raise McxError("could not place the walkers of window [%d, %d] inside it" % (lo, hi))
        eng.set_spins(np.stack(snap))

    # ---- sampling
    def sweep_(self, nsweeps=1):
        """nsweeps * N Wang-Landau attempts per walker (the reference's `for _ in 1:N; spin_flip!(sys, alg); end`
        per sweep, ising.jl:25-33 + wang_landau.jl:29-37), all local windows concurrently."""
        if not self.prepared:
            raise AssertionError("call prepare_() first")
        for eng in self.local:
            eng.wl_sweep_(nsweeps, self.logf)
        for j, eng in enumerate(self.local):
            self._lw[j] = eng.logweight()
        self.steps += int(nsweeps) * self.N * self.walkers * self.count
        return None

    def sweep_device_(self, nsweeps=1):
        """`sweep_` without reading the tables back: queues the sweeps of every local window and returns; `sync_()`
        waits for them.  The host copies of the tables go stale until the next `sweep_` / `refresh_()`."""
        if not self.prepared:
            raise AssertionError("call prepare_() first")
        for eng in self.local:
            eng.wl_sweep_(nsweeps, self.logf)
        self.steps += int(nsweeps) * self.N * self.walkers * self.count
        return None

    def sync_(self):
        for eng in self.local:
            eng.ctx.sync()

    def refresh_(self):
        for j, eng in enumerate(self.local):
            self._lw[j] = eng.logweight()

    def visits(self):
        """[count][walkers, width] visits per bin since the last update_ (the histogram the reference's
        WangLandauEnsemble does not keep: every visit lowers lw by logf)"""
        return [np.rint((s - a) / self.logf) for s, a in zip(self._lw_stage, self._lw)]

    def flatness(self):
        """min / mean of the window's visit histogram over the bins it has ever visited; one value per local
        window (1 = perfectly flat)"""
        out = []
        for v, lw in zip(self.visits(), self._lw):
            h = v.sum(axis=0)
            sup = (lw != 0.0).any(axis=0)
            out.append(float(h[sup].min() / h[sup].mean()) if sup.any() and h[sup].mean() > 0 else 0.0)
        return out

    def update_(self, power=0.5, average=True):
        """update!(ens::WangLandauEnsemble; power) (ensembles/wang_landau.jl:23) on every ensemble: logf *= power.
        With several walkers per window their tables are first replaced by the window mean."""
        for j, eng in enumerate(self.local):
            if average and self.walkers > 1:
                mean = self._lw[j].mean(axis=0)
                self._lw[j] = np.tile(mean, (self.walkers, 1))
                eng.set_logweight(self._lw[j])
            self._lw_stage[j] = self._lw[j].copy()
        self.logf *= power
        return None

    def run_(self, logf_final, sweeps_per_stage, flatness=None, max_checks=20, exchange_every=None):
        """Stages of `sweeps_per_stage` sweeps until logf <= logf_final.  With `flatness` set a stage is
        extended (up to max_checks times) until every window of every rank reaches it.  With `exchange_every`
        set the sweeps of a stage come in blocks of that many, each followed by one exchange stage."""
        while self.logf > logf_final:
            for _ in range(max_checks):
                if exchange_every:
                    done = 0
                    while done < sweeps_per_stage:
                        n = min(int(exchange_every), sweeps_per_stage - done)
                        self.sweep_(n)
                        self.exchange_()
                        done += n
                else:
                    self.sweep_(sweeps_per_stage)
                if flatness is None or self._all_min(min(self.flatness())) >= flatness:
                    break
            self.update_()
        return self

    # ---- replica exchange between neighbouring windows
    def _owner(self, w):
        return w // self.count if self.count else 0

    def _table_terms(self, j, E_mine, E_other):
        """per walker c of local window j: lw(E_other[c]) - lw(E_mine[c]) on the walker's own table, or NaN when
        E_other[c] is outside the window (then the pair is not attempted)"""
        lo, _ = self.window_energies(self.first + j)
        step = self.bins.step
        out = np.full(self.walkers, np.nan)
        for c in range(self.walkers):
            d = int(E_other[c]) - lo
            if d >= 0 and d % step == 0 and d // step < self.width:
                out[c] = self._lw[j][c, d // step] - self._lw[j][c, (int(E_mine[c]) - lo) // step]
        return out

    def _p2p(self, peer, send, recv):
        """exchange two equally shaped numpy arrays with rank `peer` (tensors on the device under NCCL)"""
        import torch
        dist = self.backend._dist
        on_gpu = dist.get_backend(self.backend.group) == "nccl"
        dev = "cuda:%d" % self.device if on_gpu else "cpu"
        t_out = torch.from_numpy(np.ascontiguousarray(send)).to(dev)
        t_in = torch.empty_like(t_out)
        ops = [dist.P2POp(dist.isend, t_out, peer, group=self.backend.group),
               dist.P2POp(dist.irecv, t_in, peer, group=self.backend.group)]
        if self.backend.rank > peer:
            ops.reverse()
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        recv[...] = t_in.cpu().numpy()

    def exchange_(self):
        """One exchange stage: window pairs (w, w + 1) with w = stage (mod 2), walker c with walker c.
        log_ratio = [lw_w(E') - lw_w(E)] + [lw_w+1(E) - lw_w+1(E')] (exchange_log_ratio, replica_exchange.jl:110-113);
        accepted = log_ratio > 0 || u < exp(log_ratio) (:115); on accept the two CONFIGURATIONS change places (the
        tables belong to the windows).  Call between sweep_ calls; the tables are not touched."""
        if not self.prepared:
            raise AssertionError("call prepare_() first")
        k, me = self.walkers, self.backend.rank
        energies = {j: np.asarray(self.local[j].energies(), dtype=np.int64) for j in range(self.count)}
        spins_cache = {}

        def spins_of(j):
            if j not in spins_cache:
                spins_cache[j] = self.local[j].spins().copy()
            return spins_cache[j]

        dirty = set()
        for w in range(self.exchange_stage % 2, self.nwindows - 1, 2):
            lo_rank, hi_rank = self._owner(w), self._owner(w + 1)
            if me not in (lo_rank, hi_rank):
                continue
            i_am_lo, i_am_hi = me == lo_rank, me == hi_rank
            j_lo, j_hi = w - self.first, w + 1 - self.first
            # energies of both sides
            E_lo = energies[j_lo] if i_am_lo else np.empty(k, dtype=np.int64)
            E_hi = energies[j_hi] if i_am_hi else np.empty(k, dtype=np.int64)
            if lo_rank != hi_rank:
                if i_am_lo:
                    self._p2p(hi_rank, E_lo, E_hi)
                else:
                    self._p2p(lo_rank, E_hi, E_lo)
            # each side evaluates the difference on its own tables
            d_lo = self._table_terms(j_lo, E_lo, E_hi) if i_am_lo else np.empty(k)
            d_hi = self._table_terms(j_hi, E_hi, E_lo) if i_am_hi else np.empty(k)
            if lo_rank != hi_rank:
                if i_am_lo:
                    self._p2p(hi_rank, d_lo, d_hi)
                else:
                    self._p2p(lo_rank, d_hi, d_lo)
            swap = np.zeros(k, dtype=bool)
            for c in range(k):
                if np.isnan(d_lo[c]) or np.isnan(d_hi[c]):
                    continue                                            # an energy outside the other window: no attempt
                u = exchange_u(self.seed, w * k + c, self.exchange_round)
                swap[c] = _accept_exchange(float(d_lo[c] + d_hi[c]), u)
                if i_am_lo:
                    self.exchange_steps[w] += 1
                    self.exchange_accepted[w] += int(swap[c])
            if not swap.any():
                continue
            if lo_rank == hi_rank:
                a, b = spins_of(j_lo), spins_of(j_hi)
                tmp = a[swap].copy()
                a[swap] = b[swap]
                b[swap] = tmp
                energies[j_lo], energies[j_hi] = np.where(swap, E_hi, E_lo), np.where(swap, E_lo, E_hi)
                dirty.update((j_lo, j_hi))
            else:
                j_mine, peer = (j_lo, hi_rank) if i_am_lo else (j_hi, lo_rank)
                mine = spins_of(j_mine)
                incoming = np.empty_like(mine[swap])
                self._p2p(peer, mine[swap], incoming)
                mine[swap] = incoming
                energies[j_mine] = np.where(swap, E_hi, E_lo) if i_am_lo else np.where(swap, E_lo, E_hi)
                dirty.add(j_mine)
        for j in dirty:
            self.local[j].set_spins(spins_cache[j])
        self.exchange_stage = 1 - self.exchange_stage
        self.exchange_round += 1
        return None

    def exchange_rates(self):
        """acceptance rate per window pair (acceptance_rates(rx), replica_exchange.jl:64-74), on every rank"""
        if self.nwindows < 2:
            return np.zeros(0)
        st, ac = self.exchange_steps.astype(np.float64), self.exchange_accepted.astype(np.float64)
        if self.backend.size > 1:
            import torch
            t = self._tensor(2 * (self.nwindows - 1))
            t[:] = torch.from_numpy(np.concatenate([st, ac])).to(t.device)
            self.backend.all_reduce_sum(t)
            v = t.cpu().numpy()
            st, ac = v[:self.nwindows - 1], v[self.nwindows - 1:]
        return np.where(st > 0, ac / np.maximum(st, 1), 0.0)

    def _all_min(self, value):
        if self.backend.size == 1:
            return value
        t = self._tensor(1)
        t[0] = -value
        self.backend.all_reduce_max(t)
        return -float(t[0].item())

    # ---- result
    def local_pieces(self):
        """-mean over walkers of lw: the window's estimate of log g up to a constant, [count, width]"""
        return np.stack([-lw.mean(axis=0) for lw in self._lw]) if self.count else np.zeros((0, self.width))

    def _tensor(self, n):
        import torch
        dist = self.backend._dist
        on_gpu = dist is not None and self.backend.size > 1 and dist.get_backend(self.backend.group) == "nccl"
        return torch.zeros(n, dtype=torch.float64, device="cuda:%d" % self.device if on_gpu else "cpu")

    def pieces(self):
        """[nwindows, width] on every rank: one all-gather of the window pieces (the only collective)."""
        import torch
        mine = self.local_pieces().reshape(-1)
        if self.backend.size == 1:
            return mine.reshape(self.nwindows, self.width)
        t = self._tensor(self.nwindows * self.width)
        a = self.first * self.width
        t[a:a + mine.size] = torch.from_numpy(mine).to(t.device)
        self.backend.all_gather_inplace(t, a, mine.size)
        return t.cpu().numpy().reshape(self.nwindows, self.width)

    def logdos(self, anchor=None):
        """Joined log g(E) as a BinnedObject over the full bins (NaN where no walker ever was).  `anchor` =
        (E, value) fixes the additive constant, e.g. (E_min, log 2) for the two Ising ground states; default:
        the lowest visited energy is 0."""
        g = join_logdos(self.pieces(), [f for f, _ in self.windows], len(self.bins))
        fin = np.flatnonzero(np.isfinite(g))
        if fin.size:
            if anchor is None:
                g -= g[fin[0]]
            else:
                g += float(anchor[1]) - g[(int(anchor[0]) - self.bins.start) // self.bins.step]
        out = BinnedObject(self.bins, np.nan)
        out.values[...] = g
        return out

    # ---- checkpoint / restart (cf. checkpoint!/restore_checkpoint, src/infrastructure/checkpointing.jl:48-101, and
    #      docs/src/examples/spin_systems/checkpoint_Ising2D.jl:40-106: the restarted run continues the same trajectory)
    def state(self):
        """Everything this rank needs to continue the run: plain numpy / Python values (picklable)."""
        if not self.prepared:
            raise AssertionError("nothing to checkpoint before prepare_()")
        return {"dims": self.dims, "bins": (self.bins.start, self.bins.stop, self.bins.step), "nwindows": self.nwindows,
                "walkers": self.walkers, "windows": list(self.windows), "seed": self.seed, "logf": self.logf,
                "rank": self.backend.rank, "size": self.backend.size, "steps": self.steps,
                "exchange": (self.exchange_stage, self.exchange_round, self.exchange_steps.copy(), self.exchange_accepted.copy()),
                "spins": [eng.spins().copy() for eng in self.local], "sweep": [eng.get_sweep() for eng in self.local],
                "lw": [a.copy() for a in self._lw], "lw_stage": [a.copy() for a in self._lw_stage]}

    @classmethod
    def restore(cls, st, backend=None, device=None, window_factory=None):
        """Rebuild the driver of this rank from `state()`; the same rank layout is required."""
        b = st["bins"]
        self = cls(st["dims"], st["nwindows"], walkers=st["walkers"], seed=st["seed"], logf=st["logf"],
                   bins=range(b[0], b[1], b[2]), backend=backend, device=device, window_factory=window_factory,
                   windows=st["windows"])
        if (self.backend.rank, self.backend.size) != (st["rank"], st["size"]):
            raise ValueError("checkpoint of rank %d / %d restored on rank %d / %d"
                             % (st["rank"], st["size"], self.backend.rank, self.backend.size))
        self.steps = st["steps"]
        self.exchange_stage, self.exchange_round = st["exchange"][0], st["exchange"][1]
        self.exchange_steps[...] = st["exchange"][2]
        self.exchange_accepted[...] = st["exchange"][3]
        for j, eng in enumerate(self.local):
            lo, _ = self.window_energies(self.first + j)
            eng.set_spins(st["spins"][j])
            eng.set_sweep(st["sweep"][j])
            eng.open_window(lo, self.bins.step, self.width)
            eng.set_logweight(st["lw"][j])
            self._lw[j] = st["lw"][j].copy()
            self._lw_stage[j] = st["lw_stage"][j].copy()
        self.prepared = True
        return self

    def spins(self):
        """[count][walkers, N] configurations of the local windows"""
        return [eng.spins() for eng in self.local]

    def energies(self):
        return [np.asarray(eng.energies()) for eng in self.local]

    def close(self):
        for eng in self.local:
            eng.close()
        self.local = []
