"""Last-seconds GPU sanity run (gpurun): core parity tests and smoke() in ONE interpreter."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)
os.makedirs("gpurun_out", exist_ok=True)
t0 = time.time()
import pytest  # noqa: E402

rc = pytest.main(["tests/test_gpu_parity.py", "tests/test_gpu_queue.py::test_default_policy_takes_the_queue_for_a_pt_rank_share",
                  "-x", "-q", "-m", "gpu", "-k",
                  "ising2d_bit_exact or batched or parallel_tempering or pt_run or split_upload or committed or default_policy"])
t1 = time.time()
import __graft_entry__ as g  # noqa: E402

g.smoke()
with open("gpurun_out/final_check.log", "w") as fh:
    fh.write("pytest rc=%s in %.1f s; smoke ok in %.1f s\n" % (rc, t1 - t0, time.time() - t1))
print(open("gpurun_out/final_check.log").read())
