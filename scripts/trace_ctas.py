"""Per-CTA timeline of one k_ising2d half-sweep launch (tuning probe, needs the MCX_OPT_TRACE build):

    make -C montecarlox.jl_b200/csrc ... -DMCX_OPT_TRACE  ->  lib/libmcx_b200_trace.so
    MCX_B200_LIB=montecarlox.jl_b200/lib/libmcx_b200_trace.so python scripts/trace_ctas.py --chains 32

Prints when CTAs start and end relative to the first start, per-SM busy spans and the item histogram.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=1024)
    ap.add_argument("--chains", type=int, default=32)
    ap.add_argument("--sweeps", type=int, default=20)
    args = ap.parse_args()
    import mcx_b200 as m
    from mcx_b200._lib import lib
    sys_ = m.Ising([args.L, args.L], nchains=args.chains)
    rng = m.PhiloxRNG(7)
    m.init_(sys_, "random", rng=rng)
    alg = m.Metropolis(rng, beta=0.44)
    m.sweep_(sys_, alg, args.sweeps)
    n = 888
    buf = np.zeros(12 * n, dtype=np.uint64)
    f = lib().mcx_debug_trace
    f.argtypes = [C.c_void_p, C.c_int]
    rc = f(buf.ctypes.data, n)
    assert rc == 0, rc
    full = buf.reshape(n, 12)
    tr = full[:, :4]
    keep = tr[:, 1] > 0
    full = full[keep]
    tr = tr[keep]
    t0 = tr[:, 0].min()
    st = (tr[:, 0] - t0).astype(np.int64)
    en = (tr[:, 1] - t0).astype(np.int64)
    print("CTAs traced", len(tr), "items histogram", np.bincount(tr[:, 3].astype(int)))
    print("start ns: min %d  p50 %d  p90 %d  max %d" % (st.min(), np.median(st), np.percentile(st, 90), st.max()))
    print("end   ns: min %d  p10 %d  p50 %d  p90 %d  max %d" % (en.min(), np.percentile(en, 10), np.median(en), np.percentile(en, 90), en.max()))
    dur = en - st
    for k in np.unique(tr[:, 3]):
        d = dur[tr[:, 3] == k]
        print("items=%d: n=%d duration ns min %d p50 %d max %d" % (k, len(d), d.min(), np.median(d), d.max()))
    sm = tr[:, 2].astype(int)
    ends = np.array([en[sm == s].max() for s in np.unique(sm)])
    starts = np.array([st[sm == s].min() for s in np.unique(sm)])
    items = np.array([tr[sm == s, 3].sum() for s in np.unique(sm)])
    print("SMs", len(ends), "items/SM min %d max %d" % (items.min(), items.max()),
          "SM first start ns: max %d" % starts.max(), " SM last end ns: min %d p50 %d max %d" % (ends.min(), np.median(ends), ends.max()))
    print("mean SM busy span / kernel span: %.3f" % ((ends - starts).mean() / en.max()))
    # per-item durations: item k of a CTA runs from its start to the next item's start (or the CTA's end)
    durs, sms, order = [], [], []
    for row in full:
        k = int(row[3])
        ts = [int(row[4 + i]) for i in range(min(k, 8))] + [int(row[1])]
        for i in range(len(ts) - 1):
            durs.append(ts[i + 1] - ts[i]); sms.append(int(row[2])); order.append(i)
    durs, sms, order = np.array(durs), np.array(sms), np.array(order)
    for i in np.unique(order):
        d = durs[order == i]
        print("item #%d of a CTA: n=%d duration ns p10 %d p50 %d p90 %d max %d" % (i, len(d), np.percentile(d, 10), np.median(d), np.percentile(d, 90), d.max()))
    per_sm = np.array([durs[sms == s_].mean() for s_ in np.unique(sms)])
    print("mean item duration per SM: min %d p50 %d max %d  (spread %.1f %%)" % (per_sm.min(), np.median(per_sm), per_sm.max(), 100 * (per_sm.max() / per_sm.min() - 1)))
    slow = np.unique(sms)[np.argsort(per_sm)[-8:]]
    fast = np.unique(sms)[np.argsort(per_sm)[:8]]
    print("slowest SMs", slow.tolist(), "fastest SMs", fast.tolist())


if __name__ == "__main__":
    main()
