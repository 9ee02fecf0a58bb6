"""Writes tests/golden/trajectories.json: seeded trajectories of the checkerboard + Philox definition
(RNG layout v1) computed by the CPU oracle.  Run from the repo root:

    python tests/golden/gen_trajectories.py

The reference cannot run in this image (Julia is absent) and none of its tests pins a trajectory, so these
vectors do not come from the reference itself: they freeze the oracle -- which IS pinned to the reference's
per-site known answers (tests/test_oracle_known_answers.py) -- so that neither the oracle nor the kernels can
drift together unnoticed.  A change of the RNG layout is a new layout version and a regenerated file."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import _golden_cases as g  # noqa: E402
from oracle import oracle  # noqa: E402

if __name__ == "__main__":
    oracle.build()
    oracle.lib()
    out = {"rng_layout": 1, "cases": {g.name_of(c): g.oracle_result(c) for c in g.cases()}}
    path = os.path.join(HERE, "trajectories.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
        fh.write("\n")
    print("wrote %d cases to %s" % (len(out["cases"]), path))
