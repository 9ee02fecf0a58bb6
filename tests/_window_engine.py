"""CPU stand-in for mcx_b200.windows.DeviceWindow, backed by the oracle (test infrastructure only).

Same methods, same random-stream layout (seed, first_chain + walker, one sweep counter shared by the
canonical and the flat-histogram sweeps, as mcx_lattice keeps it), so WangLandauWindows run against
this class and against the device must produce identical configurations and tables."""
import numpy as np

from oracle import oracle


class OracleWindow:
    def __init__(self, dims, nwalkers, seed, first_chain):
        self.dims, self.k = [int(d) for d in dims], int(nwalkers)
        self.N = int(np.prod(self.dims))
        self.seed, self.first_chain = int(seed), int(first_chain)
        self.sys = [oracle.System(oracle.ISING, self.dims) for _ in range(self.k)]
        for s in self.sys:
            s.spins = np.ones(self.N, dtype=np.int8)
        self.sweep = 0
        self.flat = None
        self.nbins = 0

    def set_spins(self, spins):
        sp = np.ascontiguousarray(spins, dtype=np.int8).reshape(self.k, self.N)
        for s, v in zip(self.sys, sp):
            s.spins = v

    def spins(self):
        return np.stack([s.spins for s in self.sys])

    def energies(self):
        return np.array([int(s.energy(full=True)) for s in self.sys], dtype=np.int64)

    def get_sweep(self):
        return self.sweep

    def set_sweep(self, n):
        self.sweep = int(n)

    def canonical_(self, rule, beta, nsweeps):
        for c, s in enumerate(self.sys):
            s.sweep_checkerboard(oracle.Alg(rule, float(beta)), self.seed, self.first_chain + c, self.sweep, int(nsweeps))
        self.sweep += int(nsweeps)

    def open_window(self, start, step, nbins):
        self.flat = [oracle.Flat(int(start), int(step), int(nbins)) for _ in range(self.k)]
        self.nbins = int(nbins)

    def wl_sweep_(self, nsweeps, logf):
        self._err = 0
        for c, (s, f) in enumerate(zip(self.sys, self.flat)):
            f.f.logf = float(logf)
            self._err |= s.flat_sweep(oracle.Alg(0, 0.0), f, 1, 0, 0.0, self.seed, self.first_chain + c, self.sweep,
                                      int(nsweeps), policy=1) != 0
        self.sweep += int(nsweeps)

    def logweight(self):
        if getattr(self, "_err", 0):
            raise IndexError("BoundsError: a walker sits outside its window")
        return np.stack([f.logweight.copy() for f in self.flat])

    def set_logweight(self, lw):
        lw = np.asarray(lw, dtype=np.float64).reshape(self.k, self.nbins)
        for f, v in zip(self.flat, lw):
            f.logweight[...] = v

    def close(self):
        self.sys, self.flat = [], None
