#!/usr/bin/env python
"""Independent exact enumeration of the density of states g(E) of the periodic L x L Ising model
(row transfer matrix with the energy kept as a polynomial degree), written for this repo.

It regenerates tests/golden/ising2d_{L}x{L}_logdos.csv.  When the reference tree is mounted
(/root/reference, this container only) it also checks the result against the reference's only
golden file, SpinSystems/data/exact_solutions/ising2D_8x8.csv (pinned by
SpinSystems/test/test_ising.jl:174-191), and records the max |difference| in the header line.
Nothing under tests/ reads /root/reference at run time; only this generator does.
"""
import math
import os
import sys

import numpy as np


def exact_dos(L):
    R = 1 << L
    rows = np.arange(R)
    bits = ((rows[:, None] >> np.arange(L)[None, :]) & 1) * 2 - 1          # [R, L] spins +-1
    Eh = -(bits * np.roll(bits, -1, axis=1)).sum(axis=1)                    # horizontal bonds of a row
    pop = np.array([bin(v).count("1") for v in range(R)])
    dist = pop[rows[:, None] ^ rows[None, :]]                               # Hamming distance a,b
    NE = L * L * 2 + 1                                                      # index (E + 2 L^2) / 2
    off = 2 * L * L
    # vec[r0, a, e]: number of row sequences r0 .. a with partial energy e (uint64, exact)
    vec = np.zeros((R, R, NE), dtype=np.uint64)
    vec[rows, rows, (Eh + off) // 2] = 1
    masks = [(dist == d).astype(np.float64) for d in range(L + 1)]
    LIMB = 26
    M = np.uint64((1 << LIMB) - 1)
    for _ in range(L - 1):
        new = np.zeros_like(vec)
        limbs = [((vec >> np.uint64(LIMB * k)) & M).astype(np.float64) for k in range(3)]
        for d in range(L + 1):
            ev = -(L - 2 * d)                                               # vertical bonds a-b
            acc = np.zeros((R, R, NE), dtype=np.uint64)
            for k in range(3):
                # sum over a of limbs[k][r0, a, e] * masks[d][a, b]  -> [r0, e, b]
                part = np.tensordot(limbs[k], masks[d], axes=([1], [0]))
                acc += np.rint(part).astype(np.uint64).transpose(0, 2, 1) << np.uint64(LIMB * k)
            # shift energy by ev + Eh[b]
            for shift in np.unique(Eh):
                sel = np.nonzero(Eh == shift)[0]
                s = (ev + int(shift)) // 2
                if s >= 0:
                    new[:, sel, s:] += acc[:, sel, :NE - s]
                else:
                    new[:, sel, :NE + s] += acc[:, sel, -s:]
        vec = new
    g = [0] * NE
    for d in range(L + 1):
        ev = -(L - 2 * d)
        s = ev // 2
        a_idx, r0_idx = np.nonzero(dist.T == d)                             # pairs (a, r0)
        closed = vec[r0_idx, a_idx, :]                                      # [pairs, NE]
        tot = [int(v) for v in closed.astype(object).sum(axis=0)]
        for e in range(NE):
            if 0 <= e + s < NE:
                g[e + s] += tot[e]
    out = []
    for e in range(NE):
        if g[e]:
            out.append((2 * e - off, g[e]))
    assert sum(c for _, c in out) == 1 << (L * L)
    return out


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dos = exact_dos(L)
    here = os.path.dirname(os.path.abspath(__file__))
    note = "unchecked (reference tree not mounted)"
    ref = "/root/reference/SpinSystems/data/exact_solutions/ising2D_%dx%d.csv" % (L, L)
    if os.path.exists(ref):
        refd = {}
        for line in open(ref).read().strip().splitlines()[1:]:
            e, v = line.split(",")
            refd[int(e)] = float(v)
        assert set(refd) == set(e for e, _ in dos), "accessible energies differ from the reference file"
        err = max(abs(refd[e] - math.log(c)) for e, c in dos)
        assert err < 1e-12, err
        note = "max |log g - reference csv| = %.3g over %d energies" % (err, len(dos))
    path = os.path.join(here, "ising2d_%dx%d_logdos.csv" % (L, L))
    with open(path, "w") as f:
        f.write("# exact periodic %dx%d Ising DOS by transfer-matrix enumeration (gen_exact_dos.py); %s\n" % (L, L, note))
        f.write("energy,count,logdos\n")
        for e, c in dos:
            f.write("%d,%d,%.17g\n" % (e, c, math.log(c)))
    print(path, note)


if __name__ == "__main__":
    main()
