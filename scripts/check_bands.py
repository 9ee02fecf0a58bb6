"""bit-exactness of banded / grouped launches against plain launches (quick GPU check)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcx_b200 as m

def run(dims, rule, track, nchains, env):
    for k in ("MCX_BANDS", "MCX_GROUPS", "MCX_RESIDENT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    s = m.Ising(dims, nchains=nchains)
    s.set_tracking(track)
    rng = m.PhiloxRNG(11)
    alg = (m.Metropolis, m.Glauber, m.HeatBath)[rule](rng, beta=0.44)
    m.init_(s, "random", rng=rng)
    l0 = s.ctx.launch_count()
    m.sweep_(s, alg, 4)
    m.sweep_(s, alg, 1)
    return s.spins.copy(), np.array(s.pair_sum()), np.array(s.magnetization()), np.array(s.accepted()), s.ctx.launch_count() - l0

ok = True
for dims, nch, envs in (([1024, 1024], 1, [{"MCX_BANDS": "2"}, {"MCX_BANDS": "4"}, {"MCX_BANDS": "8"}]),
                        ([2048, 512], 1, [{"MCX_BANDS": "4"}]),
                        ([512, 512], 6, [{"MCX_GROUPS": "2"}, {"MCX_GROUPS": "4"}, {"MCX_GROUPS": "6"}])):
    for rule in (0, 2):
        for track in (True, False):
            a = run(dims, rule, track, nch, {"MCX_BANDS": "0", "MCX_GROUPS": "0", "MCX_RESIDENT": "0"})
            for env in envs:
                e = dict(env); e["MCX_RESIDENT"] = "0"
                b = run(dims, rule, track, nch, e)
                same = all(np.array_equal(x, y) for x, y in zip(a[:4], b[:4]))
                ok &= same
                print(dims, nch, rule, track, env, "OK" if same else "MISMATCH", a[4], b[4])
print("ALL OK" if ok else "FAILED")
