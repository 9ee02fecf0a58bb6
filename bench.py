#!/usr/bin/env python
"""bench.py -- spin-flip attempts/ns of the checkerboard sweep (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[1], "2D Ising L=16384 single-chain checkerboard
Metropolis at beta_c on 1 B200", one byte per spin (256 MiB, larger than the 126 MB L2).  One STEP is
`sweep!(sys, alg, S)` -- S full sweeps = 2S half-sweep kernel launches -- followed by a read of the
step's result (energy, magnetisation, accepted count).  With N GPUs every rank runs its own
independent chain of the same size (the path shards as replicas only: weak scaling, no data-path
collective).

    python bench.py --gpus 1 --steps 5 --warmup 3            # ours
    python bench.py --impl reference --steps 3 --warmup 1    # the reference's CPU algorithm (oracle port)
    python bench.py --config c1|c3|c4|c5                     # the other BASELINE.json configs as the line's metric
    python bench.py --storage bit                            # configs[1] on one-bit-per-spin planes

The default line (configs[1]) also carries short legs of the other configs under "configs": {"c1", "c4", "c5", "bit"}
(each with its own cpu_baseline on the host cores) and "pt" / "pt_every_sweep" (configs[2]); --no-extras drops them.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BETA_C = 0.440686793509772
METRIC = "spin_flip_attempts_per_ns"
UNIT = "attempts/ns"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--L", type=int, default=16384)
    ap.add_argument("--sweeps-per-step", type=int, default=100)
    ap.add_argument("--rule", default="metropolis", choices=["metropolis", "glauber", "heatbath"])
    ap.add_argument("--track", type=int, default=0, help="1: keep energy/magnetisation sums current per flip")
    ap.add_argument("--no-pt", action="store_true", help="skip the auxiliary parallel-tempering measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-slab", action="store_true", help="skip the auxiliary slab-decomposition measurement (N > 1)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5"],
                    help="which BASELINE.json config the line's metric is measured on (c2 = configs[1], the headline)")
    ap.add_argument("--storage", default="int8", choices=["int8", "bit"], help="device storage of the c2 lattice")
    ap.add_argument("--no-extras", action="store_true", help="c2 only: skip the short legs of the other configs")
    return ap.parse_args()


def workload_config(args, world):
    return {
        "workload": "2D Ising L=%d single-chain checkerboard %s at beta_c (BASELINE.json configs[1])" % (args.L, args.rule),
        "L": args.L, "beta": BETA_C, "rule": args.rule,
        "storage": "int8 (1 byte/spin, 2 colour planes)" if args.storage == "int8" else "bit (1 bit/spin, 2 colour planes)",
        "chains_per_gpu": 1, "sweeps_per_step": args.sweeps_per_step,
        "rng": "Philox4x32-10, 32-bit draws (16-bit high half + lazy low half)",
        "l2_policy": ("input (%d MiB) larger than L2 (126 MB); no flush needed" % (args.L * args.L >> 20)) if args.storage == "int8"
                     else "input (%d MiB) fits the 126 MB L2 by design: one-bit storage is quoted in attempts/ns and L2 GB/s, not as an HBM fraction" % (args.L * args.L >> 23),
        "parallelism": "replicas only: one independent chain per GPU",
        "track_sums": bool(args.track),
    }


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes of one half-sweep of the dominant kernel (all its band launches) from the committed ncu capture, if any"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            pass
    return None


# ------------------------------------------------------------------ CPU legs (oracle: the checker/baseline only)
def cpu_baseline(args, nthreads, seconds):
    """The reference's own algorithm (random-site Metropolis with exp per non-trivial attempt,
    SpinSystems/src/ising.jl:35-41) on the same workload, one chain per thread
    (ThreadsBackend, parallel_chains.jl:99-105), bounded sample."""
    from oracle import oracle
    oracle.build()
    # calibrate on a short burst, then size the sample to ~`seconds`
    t, _ = oracle.baseline_lean(args.L, BETA_C, nthreads, 200000, False, 42)
    rate = 200000 / max(t, 1e-6)
    n = int(max(200000, min(rate * seconds, 5e8)))
    t, acc = oracle.baseline_lean(args.L, BETA_C, nthreads, n, False, 42)
    return {"value": nthreads * n / (t * 1e9), "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": "%d chain(s) x %d random-site Metropolis attempts on L=%d at beta_c (%.1f s; oracle/mcx_oracle.c "
                      "mcxo_baseline_lean: reference algorithm, exp per uphill attempt, xoshiro256++)" % (nthreads, n, args.L, t),
            "acceptance": acc}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config != "c2":
        # the other configs: the same oracle legs that cpu_baseline reports, on all host threads
        leg = {"c1": lambda: cpu_leg_c1(_threads(), 20000), "c3": lambda: cpu_leg_c3(_threads(), args.cpu_seconds),
               "c4": lambda: cpu_leg_c4(_threads()), "c5": lambda: cpu_leg_c5(min(_threads(), 32))}[args.config]()
        print(json.dumps({"impl": "reference", "metric": "pt_sweeps_per_s" if args.config == "c3" else METRIC, "value": leg["value"],
                          "unit": leg["unit"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                          "config": {"workload": "BASELINE.json config %s on the host cores (oracle port)" % args.config},
                          "cpu_baseline": leg, "e2e": {"value": leg["value"], "unit": leg["unit"], "h2d_bytes_per_step": 0,
                                                       "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return
    from oracle import oracle
    oracle.build()
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    per_chain = args.L * args.L
    nthreads = max(1, min(os.cpu_count() or 1, 64, int(avail * 0.5 // per_chain)))
    t, _ = oracle.baseline_lean(args.L, BETA_C, nthreads, 100000, False, 42)
    per_step = int(max(100000, min(100000 / max(t, 1e-6) * 8.0, 2e8)))   # ~8 s per step
    for _ in range(args.warmup):
        oracle.baseline_lean(args.L, BETA_C, nthreads, max(per_step // 8, 10000), False, 7)
    tot = 0.0
    for k in range(args.steps):
        t, _ = oracle.baseline_lean(args.L, BETA_C, nthreads, per_step, False, 100 + k)
        tot += t
    value = nthreads * per_step * args.steps / (tot * 1e9)
    sample = ("%d threads x %d attempts per step, one L=%d chain per thread (ThreadsBackend-style), random-site "
              "Metropolis, exp per uphill attempt (oracle port of SpinSystems/src/ising.jl:35-41; Julia absent)"
              % (nthreads, per_step, args.L))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": workload_config(args, 1),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def numa_bind(local):
    """Run this rank's host thread (and, by first touch, its pinned buffers) on the NUMA node its GPU hangs off: eight
    ranks streaming 256 MiB steps through one socket's memory controllers is what bent the 8-GPU e2e curve."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return {"node": None, "why": "no NUMA affinity reported for %s" % bdf}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"node": node, "why": "no allowed CPU on that node"}
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus), "gpu": bdf}
    except Exception as e:
        return {"node": None, "why": repr(e)}


# ------------------------------------------------------------------ ours
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import mcx_b200 as m
    from mcx_b200._lib import check, lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = numa_bind(local)
    if world > 1:
        import datetime
        # a mismatched collective must fail the run in minutes, not hold the box for NCCL's default ten
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    # an explicit (non-default) stream shared by torch and libmcx_b200, so that torch.cuda.Event
    # brackets exactly the kernels the library launches
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = m.Context(local, stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.config != "c2":
        # another BASELINE.json config as the line's own metric
        if args.config == "c1":
            leg = bench_c1(m, ctx, stream, args, cpu=not args.no_cpu and rank == 0)
        elif args.config == "c3":
            leg = bench_pt(m, ctx, world, rank, barrier, max_over_ranks, stream, rounds=max(args.steps, 1), every=200)
            leg["every_sweep"] = bench_pt(m, ctx, world, rank, barrier, max_over_ranks, stream, rounds=300, every=1)
            leg["config"] = {"workload": "2D Ising L=1024 ParallelTempering, 256 replicas sharded over %d GPU(s), exchange every 200 sweeps "
                                         "(BASELINE.json configs[2])" % world}
            if rank == 0 and not args.no_cpu:
                leg["cpu_baseline"] = cpu_leg_c3(_threads(), args.cpu_seconds)
        elif args.config == "c4":
            leg = bench_c4(m, ctx, stream, world, rank, barrier, max_over_ranks, cpu=not args.no_cpu)
        else:
            leg = bench_c5(m, world, rank, barrier, max_over_ranks, windows_per_gpu=32, walkers=64, cpu=not args.no_cpu, read_tables=True)   # 2048 walkers per GPU
        leg.update({"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
                    "scaling": "strong" if args.config in ("c3", "c4") else "weak", "vs_baseline": None,
                    "dtype": "u8" if args.config in ("c1", "c3") else "f64", "data": "synthetic", "clocks": None})
        if rank == 0:
            print(json.dumps(leg))
        if world > 1:
            dist.destroy_process_group()
        return

    L, S = args.L, args.sweeps_per_step
    N = L * L
    rule = {"metropolis": 0, "glauber": 1, "heatbath": 2}[args.rule]
    bit = args.storage == "bit"
    sys_ = m.Ising([L, L], ctx=ctx, storage=args.storage)
    sys_.set_tracking(bool(args.track))
    rng = m.PhiloxRNG(42, rank)
    alg = (m.Metropolis, m.Glauber, m.HeatBath)[rule](rng, beta=BETA_C)
    sys_._bind_alg(alg)
    sys_.init_("random", rng=rng)
    h = sys_.h_lat
    obs = [np.empty(1, dtype=np.int64) for _ in range(5)]

    def step():
        check(lib().mcx_sweep(h, S))                                   # 2*S half-sweeps
        check(lib().mcx_observables(h, *[o.ctypes.data for o in obs]))  # the step's result (syncs)

    for _ in range(args.warmup):
        step()

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.5)
    launches0 = ctx.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launch_count() - launches0

    # dominant kernel alone (roofline): 2*S half-sweeps of k_ising2d, nothing else in between.  A half-sweep of a
    # big lattice is issued as `bands` launches of the same kernel on as many streams (row bands whose tails
    # overlap), so the unit timed here is the half-sweep = bands launches, N/2 attempts.
    barrier()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nk = 2 * S
    lk0 = ctx.launch_count()
    k0.record(stream)
    check(lib().mcx_sweep(h, S))
    k1.record(stream)
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / nk
    launches_per_half_sweep = (ctx.launch_count() - lk0) / nk
    clocks = sampler.stop()

    value = world * args.steps * S * N / (ms * 1e6)
    peak, peak_src = measured_peak()
    bytes_per_attempt = 0.375 if bit else 3.0                        # 3 x storage bytes per spin
    bytes_per_launch = bytes_per_attempt * (N // 2)                   # x N/2 attempts per half-sweep
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    traffic = None if bit else ncu_traffic()
    roofline = {"bound": "hbm", "kernel": "k_ising2d_bits (half-sweep)" if bit else "k_ising2d (half-sweep)", "achieved": achieved,
                "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src, "bytes_per_attempt": bytes_per_attempt,
                "attempts_per_launch": (N // 2) / launches_per_half_sweep, "launches_per_half_sweep": launches_per_half_sweep,
                "attempts_per_half_sweep": N // 2, "kernel_ms": kernel_ms,
                "kernel_ms_is": "one half-sweep (all its concurrent band launches), CUDA events on the launching stream",
                "kernel_attempts_per_ns": (N // 2) / (kernel_ms * 1e6),
                "traffic": (traffic.get("dram_bytes_per_half_sweep") or traffic.get("dram_bytes_per_launch")) if traffic else None,
                "traffic_source": traffic.get("source") if traffic else None}

    # ---- e2e: the public API with HOST buffers; H2D of the step's input and D2H of its result inside the timed region
    host = torch.empty(N, dtype=torch.int8).pin_memory()
    host.copy_(torch.from_numpy(sys_.spins))

    def e2e_step():
        sys_.upload_from(host.data_ptr())       # H2D N bytes from pinned memory + layout pack + recompute
        check(lib().mcx_sweep(h, S))
        check(lib().mcx_observables(h, *[o.ctypes.data for o in obs]))   # D2H 5 x int64

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    serial_s = max_over_ranks(time.perf_counter() - t0)

    # the same steps through the split upload: every step still copies its own N input bytes from pinned host memory
    # and reads its result back, but the copy of step k+1 crosses PCIe while step k is being swept
    hosts = [host, torch.empty(N, dtype=torch.int8).pin_memory()]
    hosts[1].copy_(host)

    def e2e_pipelined(nsteps):
        sys_.upload_begin(hosts[0].data_ptr())
        for k in range(nsteps):
            sys_.upload_commit()                                         # step k's input: wait for its copy, pack, recompute
            if k + 1 < nsteps:
                sys_.upload_begin(hosts[(k + 1) & 1].data_ptr())         # step k+1's input, behind step k's sweeps
            check(lib().mcx_sweep(h, S))
            check(lib().mcx_observables(h, *[o.ctypes.data for o in obs]))   # D2H 5 x int64 (syncs)

    e2e_pipelined(2)
    barrier()
    t0 = time.perf_counter()
    e2e_pipelined(args.steps)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": world * args.steps * S * N / (e2e_s * 1e9), "unit": UNIT, "h2d_bytes_per_step": N,
           "d2h_bytes_per_step": 40, "ms_per_step": e2e_s / args.steps * 1e3,
           "api": "mcx_lattice_upload_begin/_commit + mcx_sweep + mcx_observables (C ABI, two pinned host buffers; "
                  "the H2D copy of step k+1 overlaps the sweeps of step k)",
           "numa": numa,
           "serial": {"value": world * args.steps * S * N / (serial_s * 1e9), "ms_per_step": serial_s / args.steps * 1e3,
                      "api": "mcx_lattice_upload + mcx_sweep + mcx_observables, nothing overlapped"}}

    # the same pipelined steps with the host side at one bit per spin (N / 8 bytes in), and with the step's result being the
    # whole configuration (N / 8 bytes out) instead of five integers
    hbits = [torch.empty(N // 8, dtype=torch.uint8).pin_memory() for _ in range(2)]
    hbits[0].copy_(torch.from_numpy(np.ascontiguousarray(sys_.spin_bits)))
    hbits[1].copy_(hbits[0])
    out_bits = torch.empty(N // 8, dtype=torch.uint8).pin_memory()

    def e2e_bits(nsteps, download):
        sys_.upload_bits_begin(hbits[0].data_ptr())
        for k in range(nsteps):
            sys_.upload_commit()
            if k + 1 < nsteps and not download:
                sys_.upload_bits_begin(hbits[(k + 1) & 1].data_ptr())        # behind step k's sweeps
            check(lib().mcx_sweep(h, S))
            if download:                                                     # the download shares the handle's bit staging buffer:
                check(lib().mcx_lattice_download_bits(h, out_bits.data_ptr()))   # D2H N / 8 bytes (syncs), then the next upload
                if k + 1 < nsteps:
                    sys_.upload_bits_begin(hbits[(k + 1) & 1].data_ptr())
            check(lib().mcx_observables(h, *[o.ctypes.data for o in obs]))

    for download, key in ((False, "bit_buffers"), (True, "bit_buffers_with_download")):
        e2e_bits(2, download)
        barrier()
        t0 = time.perf_counter()
        e2e_bits(args.steps, download)
        torch.cuda.synchronize()
        sec = max_over_ranks(time.perf_counter() - t0)
        e2e[key] = {"value": world * args.steps * S * N / (sec * 1e9), "ms_per_step": sec / args.steps * 1e3,
                    "h2d_bytes_per_step": N // 8, "d2h_bytes_per_step": 40 + (N // 8 if download else 0),
                    "api": "mcx_lattice_upload_bits_begin/_commit + mcx_sweep" + (" + mcx_lattice_download_bits" if download else "") +
                           " + mcx_observables"}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "u8", "data": "synthetic", "config": workload_config(args, world), "roofline": roofline,
           "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
           "result": {"energy_per_site": -float(obs[0][0]) / N, "abs_m": abs(float(obs[1][0])) / N,
                      "acceptance": float(obs[3][0]) / max(float(obs[4][0]), 1.0)}}

    # ---- auxiliary: parallel tempering sweeps/s (BASELINE.json configs[2]), replicas sharded over ranks
    if not args.no_pt:
        try:   # reference cadence (pt_Ising2D.jl:37: exchange every 200 sweeps) and the worst case (every sweep)
            out["pt"] = bench_pt(m, ctx, world, rank, barrier, max_over_ranks, stream, rounds=5, every=200)
            out["pt_every_sweep"] = bench_pt(m, ctx, world, rank, barrier, max_over_ranks, stream, rounds=300, every=1)
        except Exception as e:   # auxiliary metric must not take the headline down
            out["pt"] = {"error": repr(e)}

    # ---- short legs of the other configs (each with its own cpu_baseline at N = 1), the tracked-sums rate and the bit planes
    if not args.no_extras:
        extras = {}
        try:
            extras["tracked_sums"] = bench_plain(m, ctx, stream, L, S, rule, rank, "int8", True, barrier, max_over_ranks, world)
            extras["bit" if not bit else "int8"] = bench_plain(m, ctx, stream, L, S, rule, rank, "int8" if bit else "bit", False, barrier,
                                                               max_over_ranks, world)
        except Exception as e:
            extras["tracked_sums"] = {"error": repr(e)}
        cpu = world == 1 and not args.no_cpu
        for key, fn in (("c1", lambda: bench_c1(m, ctx, stream, args, cpu=cpu and rank == 0)),
                        ("c4", lambda: bench_c4(m, ctx, stream, world, rank, barrier, max_over_ranks, iterations=2, therm=1, record=4, cpu=cpu)),
                        ("c5", lambda: bench_c5(m, world, rank, barrier, max_over_ranks, windows_per_gpu=16, walkers=32, cpu=cpu))):
            if not (key == "c1" and rank != 0):          # c1 is one small chain: rank 0 only; every rank meets at the barrier
                try:
                    extras[key] = fn()
                except Exception as e:
                    extras[key] = {"error": repr(e)}
            barrier()
        out["configs"] = extras

    # ---- auxiliary: ONE lattice split by rows over the ranks (slab decomposition, halo rows read over NVLink)
    if world > 1 and not args.no_slab:
        try:
            out["slab_strong"] = bench_slab(m, ctx, world, barrier, max_over_ranks, stream, [L, L], 20)
            out["slab_weak"] = bench_slab(m, ctx, world, barrier, max_over_ranks, stream, [L, L * world], 20)
        except Exception as e:
            out["slab_strong"] = {"error": repr(e)}

    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            out["cpu_baseline"] = cpu_baseline(args, 1, args.cpu_seconds)
        except Exception as e:
            out["cpu_baseline"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def bench_plain(m, ctx, stream, L, S, rule, rank, storage, track, barrier, max_over_ranks, world):
    """attempts/ns of the configs[1] lattice in another mode (tracked sums = modify!-like bookkeeping per flip; the other storage)"""
    import torch
    from mcx_b200._lib import check, lib
    s = m.Ising([L, L], ctx=ctx, storage=storage)
    s.set_tracking(track)
    rng = m.PhiloxRNG(42, rank)
    alg = (m.Metropolis, m.Glauber, m.HeatBath)[rule](rng, beta=BETA_C)
    s._bind_alg(alg)
    s.init_("random", rng=rng)
    check(lib().mcx_sweep(s.h_lat, S))        # warm-up with the timed call's own length (a long series builds its CUDA graph once)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    check(lib().mcx_sweep(s.h_lat, S))
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    N = L * L
    bpa = 0.375 if storage == "bit" else 3.0
    peak, _ = measured_peak()
    pair = int(s.pair_sum())
    return {"metric": METRIC, "value": world * S * N / (ms * 1e6), "unit": UNIT, "storage": storage, "track_sums": bool(track),
            "per_gpu_attempts_per_ns": S * N / (ms * 1e6), "algorithmic_GBps": bpa * S * N / (ms * 1e-3) / 1e9,
            "frac_of_hbm_peak": bpa * S * N / (ms * 1e-3) / 1e9 / peak,
            "note": ("%d MiB of planes: L2-resident, bound by instruction issue (Philox + decision), not by memory" % (N >> 22)) if storage == "bit" else None,
            "pair_sum_rank0": pair}


def bench_slab(m, ctx, world, barrier, max_over_ranks, stream, dims, sweeps):
    """attempts/ns of ONE [Lx, Ly] lattice whose rows are split over the ranks; every half-sweep reads two halo
    rows from the neighbour GPUs' memory and is ordered against them by device flags (k_slab.cu)"""
    import torch
    from mcx_b200._lib import check, lib
    s = m.SlabIsing(dims, backend=m.GPUBackend(), ctx=ctx)
    s.set_tracking(False)
    rng = m.PhiloxRNG(42, 0)
    alg = m.Metropolis(rng, beta=BETA_C)
    s.init_("random", rng=rng)
    part = s.parts[0]
    part._bind_alg(alg)
    check(lib().mcx_sweep(part.h_lat, 5))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    check(lib().mcx_sweep(part.h_lat, sweeps))
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    timed_out = max(t for t, _ in s.status())
    n = dims[0] * dims[1]
    return {"metric": METRIC, "value": sweeps * n / (ms * 1e6), "unit": UNIT, "dims": dims, "rows_per_gpu": dims[1] // world,
            "sweeps": sweeps, "us_per_half_sweep": ms * 1e3 / (2 * sweeps), "energy_per_site": s.energy(full=True) / n,
            "parity": {"kind": "identical at every rank count (randomness positioned by global row)", "energy": float(s.energy(full=True))},
            "wait_timed_out": int(timed_out),
            "halo": "2 rows x %d B per half-sweep per GPU, peer loads over NVLink (CUDA IPC), no copies" % (dims[0] // 2)}


def bench_pt(m, ctx, world, rank, barrier, max_over_ranks, stream, L=1024, n=256, rounds=200, every=1):
    import torch
    betas = m.set_betas(n, 1 / 3.0, 1 / 1.5, "uniform")       # T in [1.5, 3], pt_Ising2D.jl:40-41
    backend = m.GPUBackend()
    pt = m.ParallelTempering(betas, seed=42, backend=backend)
    first, count = backend.slots(n)
    reps = m.Ising([L, L], nchains=count, ctx=ctx)
    pt.attach(reps)
    reps.init_("random", rng=m.PhiloxRNG(42, first))

    # rounds x (`every` sweeps of all replicas, then update!(pt)): the loop of pt_Ising2D.jl:52-57, queued by
    # ParallelTempering.run_ (one library call when the energies reach all ranks without the host)
    pt.run_(reps, 3 if every > 1 else 20, every)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    pt.run_(reps, rounds, every)
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    sweeps = rounds * every
    # parity across rank counts: ladder labels and the energies of the last exchange are the same at every N
    # (replicas keyed by global slot, u keyed by round), so their digest must not change along the scaling run
    try:
        parity = {"kind": "identical at every rank count", "labels_and_energies_sha": _sha(__import__("numpy").asarray(pt.index()), pt.energies())}
    except Exception as e:
        parity = {"error": repr(e)}
    return {"metric": "pt_sweeps_per_s", "value": sweeps / (ms * 1e-3), "unit": "PT sweeps/s (all %d replicas swept once)" % n,
            "parity": parity,
            "attempts_per_ns": sweeps * n * L * L / (ms * 1e6), "L": L, "replicas": n, "exchange_every": every,
            "rounds": rounds, "scaling": "strong", "exchange_acceptance": pt.acceptance_rate(),
            "collective": "none (1 rank)" if world == 1 else
                          ("peer stores of %d doubles per exchange over NVLink (CUDA IPC), device-side arrival counters, no collective call" % n
                           if getattr(pt, "_peers", False) else "NCCL all-gather of %d doubles per exchange" % n)}


# ------------------------------------------------------------------ the other BASELINE.json configs
# Each leg returns {"metric", "value", "unit", "config", "roofline", "e2e", "cpu_baseline"?, "parity"?, ...}: the default
# line embeds short runs of them under "configs"; `--config cX` prints one as the line itself.
def _sha(*arrays):
    import hashlib
    import numpy as np
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def _threads():
    return max(1, min(os.cpu_count() or 1, 64))


def _run_threads(fn, n):
    """fn(i) on n Python threads (the oracle's C loops release the GIL); returns wall seconds"""
    th = [threading.Thread(target=fn, args=(i,)) for i in range(n)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    return time.perf_counter() - t0


def _moments(series, N):
    """<E>/N, <|m|>, U4 = 1 - <m^4> / (3 <m^2>^2) from int64 [n, chains, 4] snapshots {pair, spin, spin2, accepted}"""
    import numpy as np
    e = -series[:, :, 0].astype(np.float64) / N
    mm = series[:, :, 1].astype(np.float64) / N
    m2, m4 = (mm ** 2).mean(), (mm ** 4).mean()
    return {"energy_per_site": float(e.mean()), "abs_m": float(np.abs(mm).mean()), "U4": float(1.0 - m4 / (3.0 * m2 * m2))}


def cpu_leg_c1(nthreads, sweeps):
    """the reference's random-site Metropolis loop, L = 64 at beta_c, seed 42,
    one chain per thread (oracle port of SpinSystems/src/ising.jl:35-41 + importance_sampling.jl:80-85)"""
    from oracle import oracle
    oracle.build()
    L, therm, interval = 64, 1000, 10
    t0 = time.perf_counter()
    st = oracle.stats_random_site(L, BETA_C, nthreads, therm, sweeps, interval, nthreads, 42)
    dt = time.perf_counter() - t0
    attempts = nthreads * (therm + sweeps) * L * L
    m2, m4 = st[:, 2].mean(), st[:, 3].mean()
    spread = {}
    if nthreads > 1:        # independent chains: the error of their mean from the chain-to-chain spread
        spread = {"err_energy_per_site": float(st[:, 0].std(ddof=1) / nthreads ** 0.5), "err_abs_m": float(st[:, 1].std(ddof=1) / nthreads ** 0.5)}
    return {**spread, "value": attempts / (dt * 1e9), "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": "%d chain(s) x (%d + %d) random-site sweeps of L=64 at beta_c, seed 42, xoshiro256++ (%.1f s; oracle "
                      "mcxo_stats_random_site)" % (nthreads, therm, sweeps, dt),
            "energy_per_site": float(st[:, 0].mean()), "abs_m": float(st[:, 1].mean()), "U4": float(1.0 - m4 / (3.0 * m2 * m2))}


def bench_c1(m, ctx, stream, args, sweeps=100000, cpu=True):
    """configs[0]: 2-D Ising L = 64 Metropolis at beta_c, 10^5 sweeps, <E>/N, <|m|>, U4 (importance_Ising2D.jl:105-122).
    On the device the lattice lives in shared memory for the whole series (k_ising2d_resident); the measurements are
    snapshots of the tracked sums every `interval` sweeps (mcx_sweep_series)."""
    import numpy as np
    import torch
    from mcx_b200._lib import check, lib
    L, interval, therm = 64, 10, 1000
    N = L * L
    nmeasure = sweeps // interval
    sys_ = m.Ising([L, L], ctx=ctx)
    rng = m.PhiloxRNG(42, 0)
    alg = m.Metropolis(rng, beta=BETA_C)
    sys_._bind_alg(alg)
    sys_.init_("random", rng=rng)
    check(lib().mcx_sweep(sys_.h_lat, therm))
    series = np.empty((nmeasure, 1, 4), dtype=np.int64)
    check(lib().mcx_sweep_series(sys_.h_lat, min(nmeasure, 100), interval, series.ctypes.data))   # warm-up (grows the snapshot buffer)
    check(lib().mcx_sweep_series(sys_.h_lat, nmeasure, interval, series.ctypes.data))
    torch.cuda.synchronize()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    check(lib().mcx_sweep_series(sys_.h_lat, nmeasure, interval, series.ctypes.data))
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    stats = _moments(series, N)
    # error bars from the integrated autocorrelation times, computed on the device from the series it still holds
    # (mcx_series_tau_int; autocorrelations.jl:28-65): err = sqrt(2 tau / n * var)
    try:
        from mcx_b200.measurements import series_tau_int_
        tau_e = float(series_tau_int_(sys_, nmeasure, "energy")[0])
        tau_m = float(series_tau_int_(sys_, nmeasure, "abs_magnetization")[0])
        e = -series[:, 0, 0].astype(np.float64) / N
        am = np.abs(series[:, 0, 1].astype(np.float64)) / N
        stats.update({"tau_int_energy_in_measurements": tau_e, "tau_int_abs_m_in_measurements": tau_m,
                      "err_energy_per_site": float(np.sqrt(2 * tau_e / nmeasure * e.var())),
                      "err_abs_m": float(np.sqrt(2 * tau_m / nmeasure * am.var()))})
    except Exception as ex:
        stats["tau_int_error"] = repr(ex)
    # e2e: host spins in (pinned), the series out
    host = torch.from_numpy(sys_.spins.copy()).pin_memory()
    t0 = time.perf_counter()
    sys_.upload_from(host.data_ptr())
    check(lib().mcx_sweep_series(sys_.h_lat, nmeasure, interval, series.ctypes.data))
    e2e_s = time.perf_counter() - t0
    attempts = nmeasure * interval * N
    peak, _ = measured_peak()
    achieved = 3.0 * attempts / (ms * 1e-3) / 1e9
    out = {"metric": METRIC, "value": attempts / (ms * 1e6), "unit": UNIT, "ms": ms, "gpu_launches": int(launches),
           "config": {"workload": "2D Ising L=64 Metropolis at beta_c, %d sweeps, seed 42 (BASELINE.json configs[0])" % sweeps,
                      "L": L, "sweeps": sweeps, "measure_every": interval, "thermalisation_sweeps": therm},
           "us_per_sweep": ms * 1e3 / (nmeasure * interval),
           "roofline": {"bound": "hbm", "kernel": "k_ising2d_resident (both colour planes in shared memory for %d sweeps per launch)" % interval,
                        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                        "note": "3 B/attempt is the streaming figure; this kernel touches HBM for 2 B/site per launch only, so the "
                                "fraction says how far a 4096-site lattice is from filling the machine, not kernel quality"},
           "e2e": {"value": attempts / (e2e_s * 1e9), "unit": UNIT, "h2d_bytes_per_step": N, "d2h_bytes_per_step": int(series.nbytes),
                   "api": "mcx_lattice_upload + mcx_sweep_series (C ABI)"},
           "result": stats}
    if cpu:
        try:
            out["cpu_baseline"] = cpu_leg_c1(1, sweeps)
            out["cpu_baseline_all_cores"] = cpu_leg_c1(_threads(), max(sweeps // 4, 1000))
            ref = out["cpu_baseline"]
            ref = out["cpu_baseline_all_cores"]           # 16+ independent chains: the better-averaged reference leg
            out["parity"] = {"kind": "statistical (different update order and generator: random-site xoshiro vs checkerboard Philox); "
                                     "differences against the all-cores reference leg, z in units of the combined error",
                             "d_energy_per_site": stats["energy_per_site"] - ref["energy_per_site"],
                             "d_abs_m": stats["abs_m"] - ref["abs_m"], "d_U4": stats["U4"] - ref["U4"],
                             "z_energy": (stats["energy_per_site"] - ref["energy_per_site"]) /
                                         max((stats.get("err_energy_per_site", 0.0) ** 2 + ref.get("err_energy_per_site", 0.0) ** 2) ** 0.5, 1e-12),
                             "z_abs_m": (stats["abs_m"] - ref["abs_m"]) /
                                        max((stats.get("err_abs_m", 0.0) ** 2 + ref.get("err_abs_m", 0.0) ** 2) ** 0.5, 1e-12),
                             "errors": "device: sqrt(2 tau_int var / n) with tau_int from mcx_series_tau_int; reference: spread over its independent chains"}
        except Exception as e:
            out["cpu_baseline"] = {"error": repr(e)}
    return out


def cpu_leg_c3(nthreads, seconds, L=1024, every=200):
    """BASELINE.md section 4 'PT': replicas = host threads, L = 1024, one exchange per `every` sweeps
    (pt_Ising2D.jl:37-57 on ThreadsBackend: random-site sweeps per replica, then update!(rx))"""
    import numpy as np
    from oracle import oracle
    oracle.build()
    n = nthreads
    betas = oracle.set_betas(n, 1 / 3.0, 1 / 1.5)
    sysv = [oracle.System(oracle.ISING, [L, L]) for _ in range(n)]
    algs = [oracle.Alg(oracle.METROPOLIS, float(b)) for b in betas]
    xos = [oracle.xoshiro(42 + i) for i in range(n)]
    for i, s_ in enumerate(sysv):
        s_.init_random(42, i)
    N = L * L
    t = _run_threads(lambda i: sysv[i].sweep_random_site(algs[i], xos[i], N), n)      # calibrate: one sweep each
    sweeps = int(max(1, min(every, seconds / max(t, 1e-6))))
    dt = _run_threads(lambda i: sysv[i].sweep_random_site(algs[i], xos[i], sweeps * N), n)
    t0 = time.perf_counter()
    idx = np.arange(1, n + 1, dtype=np.int64)
    steps, acc = np.zeros(max(n - 1, 1), dtype=np.int64), np.zeros(max(n - 1, 1), dtype=np.int64)
    xs = np.array([s_.energy() for s_ in sysv])
    if n > 1:
        oracle.rx_update(0, idx, steps, acc, betas[idx - 1], xs, np.full(n, 0.5))
    dt += (time.perf_counter() - t0) * sweeps / every          # the exchange's share of `sweeps` sweeps
    return {"value": sweeps / dt, "unit": "PT sweeps/s (all %d replicas swept once)" % n, "cores": nthreads, "kind": "port",
            "attempts_per_ns": sweeps * n * N / (dt * 1e9),
            "sample": "%d replicas of L=%d (one per thread), %d random-site sweeps each + 1/%d of an update!(rx) (%.1f s; oracle "
                      "mcxo_sweep_random_site + mcxo_rx_update)" % (n, L, sweeps, every, dt)}


def cpu_leg_c4(nthreads, sweeps=1, L=512):
    """config 4 on the host: one Blume-Capel L = 512 multicanonical chain per thread, pair term at T = 0.9, visits on
    sum s^2 (muca_BlumeCapel.jl:33-108 through ThreadsBackend); oracle mcxo_flat_sweep_policy"""
    from oracle import oracle
    oracle.build()
    N = L * L
    sysv = [oracle.System(oracle.BLUME_CAPEL, [L, L]) for _ in range(nthreads)]
    flats = [oracle.Flat(0, 1, N + 1) for _ in range(nthreads)]
    algs = [oracle.Alg(0, 0.0) for _ in range(nthreads)]
    dt = _run_threads(lambda i: sysv[i].flat_sweep(algs[i], flats[i], 0, 1, 1 / 0.9, 42, i, 0, sweeps, policy=0), nthreads)
    return {"value": nthreads * sweeps * N / (dt * 1e9), "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": "%d chain(s) x %d multicanonical sweep(s) of Blume-Capel L=%d, one chain per thread (%.1f s)" % (nthreads, sweeps, L, dt)}


def bench_c4(m, ctx, stream, world, rank, barrier, max_over_ranks, iterations=3, therm=2, record=8, L=512, chains=1024, cpu=True):
    """configs[3]: Blume-Capel L = 512, 1024 chains, multicanonical weight iteration in sum s^2 (muca_BlumeCapel.jl:33-108):
    per iteration thermalise, reset!, record, merge_histograms! (all-reduce when the chains are sharded over ranks),
    update!.  Everything stays on the device (C ABI: mcx_flat_*)."""
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    from mcx_b200._lib import check, lib
    N = L * L
    count = chains // world
    first = rank * count
    sys_ = m.BlumeCapel([L, L], nchains=count, ctx=ctx)
    check(lib().mcx_lattice_set_first_chain_id(sys_.h_lat, first))
    sys_.set_rng(42, 0)
    h = C.c_void_p()
    check(lib().mcx_flat_create(sys_.h_lat, m._lib.FLAT_MUCA, m._lib.OBS_SPIN2_WITH_PAIR_BOLTZMANN, 0, 1, N + 1, 1 / 0.9, 0, C.byref(h)))
    hist_ptr, nb = C.c_void_p(), C.c_int64()
    check(lib().mcx_flat_device_histogram(h, C.byref(hist_ptr), C.byref(nb)))
    from mcx_b200.parallel import _as_torch
    hist = _as_torch(hist_ptr.value, nb.value, torch.int64, ctx.device)

    def iteration():
        check(lib().mcx_flat_sweep(h, therm))
        check(lib().mcx_flat_reset_histogram(h))
        check(lib().mcx_flat_sweep(h, record))
        if world > 1:
            dist.all_reduce(hist)                                # merge_histograms! (parallel_multicanonical.jl:38-52)
        check(lib().mcx_flat_update(h))                          # every rank applies the same update!: no broadcast needed

    iteration()                                                  # warm-up (also an iteration of the weights)
    barrier()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iterations):
        iteration()
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launch_count() - l0
    lw = np.empty(N + 1, dtype=np.float64)
    check(lib().mcx_flat_get_logweight(h, lw.ctypes.data))
    hv = np.empty(N + 1, dtype=np.float64)
    check(lib().mcx_flat_get_histogram(h, hv.ctypes.data))
    # e2e: the host ensemble objects stay the source of truth (flat.py DeviceFlat): tables in, histogram + tables out
    t0 = time.perf_counter()
    check(lib().mcx_flat_set_logweight(h, lw.ctypes.data))
    iteration()
    check(lib().mcx_flat_get_logweight(h, lw.ctypes.data))
    check(lib().mcx_flat_get_histogram(h, hv.ctypes.data))
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    sweeps = iterations * (therm + record)
    attempts = sweeps * chains * N
    peak, _ = measured_peak()
    achieved = 2.0 * attempts / world / (ms * 1e-3) / 1e9
    out = {"metric": METRIC, "value": attempts / (ms * 1e6), "unit": UNIT, "iterations_per_s": iterations / (ms * 1e-3), "ms": ms,
           "gpu_launches": int(launches), "n_gpus": world,
           "config": {"workload": "2D Blume-Capel L=%d multicanonical weight iteration in sum s^2, %d chains (BASELINE.json configs[3])" % (L, chains),
                      "iterations": iterations, "therm_sweeps": therm, "record_sweeps": record, "T_pair": 0.9,
                      "parallelism": "chains sharded over the ranks; one all-reduce of the %d-bin histogram per iteration" % (N + 1)},
           "roofline": {"bound": "hbm", "kernel": "k_flat_warp (one warp per chain; serial in the chain's sum s^2)", "achieved": achieved,
                        "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                        "note": "2 B/attempt (read + conditional write of the site); latency-bound by construction (BASELINE.md section 3)"},
           "e2e": {"value": (therm + record) * chains * N / (e2e_s * 1e9), "unit": UNIT, "h2d_bytes_per_step": int(lw.nbytes),
                   "d2h_bytes_per_step": int(lw.nbytes + hv.nbytes), "api": "mcx_flat_set_logweight + iteration + mcx_flat_get_logweight/_histogram"},
           "parity": {"kind": "identical at every rank count (chains keyed by global id, integer histogram)",
                      "histogram_sha": _sha(hv), "logweight_sha": _sha(lw), "visits": float(hv.sum())}}
    check(lib().mcx_flat_destroy(h))
    if cpu and rank == 0:
        try:
            out["cpu_baseline"] = cpu_leg_c4(_threads())
        except Exception as e:
            out["cpu_baseline"] = {"error": repr(e)}
    return out


def cpu_leg_c5(nthreads, L=256):
    """config 5 on the host: one Wang-Landau walker of the 3-D Ising L = 256 lattice per thread, one sweep each, the window
    being the whole upper half of the spectrum around E = 0 (oracle mcxo_flat_sweep_policy, policy 1)"""
    from oracle import oracle
    oracle.build()
    N = L ** 3
    sysv = [oracle.System(oracle.ISING, [L, L, L]) for _ in range(nthreads)]
    for i, s_ in enumerate(sysv):
        s_.init_random(42, i)
    flats = [oracle.Flat(-(1 << 22), 4, (1 << 21) + 1) for _ in range(nthreads)]
    algs = [oracle.Alg(0, 0.0) for _ in range(nthreads)]
    dt = _run_threads(lambda i: sysv[i].flat_sweep(algs[i], flats[i], 1, 0, 0.0, 42, i, 0, 1, policy=1), nthreads)
    return {"value": nthreads * N / (dt * 1e9), "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": "%d walker(s) x 1 Wang-Landau sweep of 3-D Ising L=%d, one walker per thread (%.1f s)" % (nthreads, L, dt)}


def bench_c5(m, world, rank, barrier, max_over_ranks, windows_per_gpu=8, walkers=32, L=256, sweeps=1, cpu=True, read_tables=False):
    """configs[4]: 3-D Ising L = 256 Wang-Landau, energy windows dealt to the ranks, `walkers` walkers per window with one
    table each (windows.py).  Timed: `sweeps` sweeps of every walker, tables left on the device."""
    import numpy as np
    backend = m.GPUBackend()
    nwin = windows_per_gpu * world
    wl = m.WangLandauWindows([L] * 3, nwindows=nwin, walkers=walkers, overlap=0.5, seed=42, backend=backend, replicate_seed=True)
    t0 = time.perf_counter()
    wl.prepare_()
    t_prepare = time.perf_counter() - t0
    l0 = sum(e.ctx.launch_count() for e in wl.local)
    barrier()
    t0 = time.perf_counter()
    wl.sweep_device_(sweeps)
    wl.sync_()
    dt = max_over_ranks(time.perf_counter() - t0)
    launches = sum(e.ctx.launch_count() for e in wl.local) - l0
    barrier()
    energies = np.concatenate([np.asarray(e) for e in wl.energies()])
    inside = all(((wl.window_energies(wl.first + j)[0] <= e) & (e <= wl.window_energies(wl.first + j)[1])).all()
                 for j, e in enumerate(wl.energies()))
    if world > 1:
        import torch.distributed as dist
        gathered = [None] * world
        dist.all_gather_object(gathered, energies)
        energies = np.concatenate(gathered)
    # parity across rank counts: the timed problem grows with the ranks (windows per GPU fixed), so the digest that must not
    # change along a scaling run comes from a problem of FIXED size dealt over the ranks -- 2-D 32 x 32, 8 windows x 2
    # walkers, two stages, joined log g (one all-gather of the window pieces)
    try:
        small = m.WangLandauWindows([32, 32], nwindows=8, walkers=2, overlap=0.5, seed=7, backend=backend)
        small.prepare_().run_(0.25, 30)
        fixed_sha = _sha(np.nan_to_num(small.logdos().values, nan=-1.0))
        small.close()
    except Exception as ex:
        fixed_sha = "error: " + repr(ex)
    # e2e: the same sweeps followed by the step's result on the host -- every walker's energy, or (read_tables: what
    # WangLandauWindows.sweep_ hands back) every walker's whole log-weight table
    t0 = time.perf_counter()
    if read_tables:
        wl.sweep_(sweeps)
        table_bytes = sum(a.nbytes for a in wl._lw)
    else:
        wl.sweep_device_(sweeps)
        table_bytes = sum(np.asarray(e).nbytes for e in wl.energies())      # waits for the sweeps
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    attempts = sweeps * wl.N * walkers * nwin
    peak, _ = measured_peak()
    achieved = 2.0 * attempts / world / dt / 1e9
    out = {"metric": METRIC, "value": attempts / (dt * 1e9), "unit": UNIT, "ms": dt * 1e3, "gpu_launches": int(launches), "n_gpus": world,
           "config": {"workload": "3D Ising L=%d Wang-Landau, %d energy windows x %d walkers over %d GPU(s) (BASELINE.json configs[4])"
                                  % (L, nwin, walkers, world),
                      "bins_per_window": wl.width, "overlap": 0.5, "logf": wl.logf, "sweeps": sweeps, "seeding": "one driven walker per window, replicated",
                      "parallelism": "windows dealt to the ranks in contiguous blocks; no collective while sampling"},
           "prepare_seconds": t_prepare, "walkers_inside_windows": bool(inside),
           "roofline": {"bound": "hbm", "kernel": "k_flat_warp (one warp per walker, private log-weight table, window of it in shared memory)",
                        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                        "note": "2 B/attempt; serial in the walker's energy, latency-bound by construction (BASELINE.md section 3)"},
           "e2e": {"value": attempts / (e2e_s * 1e9), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(table_bytes),
                   "api": "WangLandauWindows.sweep_ (mcx_flat_sweep + mcx_flat_get_logweight of every walker)" if read_tables else
                          "WangLandauWindows.sweep_device_ + energies() (mcx_flat_sweep + mcx_observables); the tables stay on the device"},
           "parity": {"kind": "fixed_problem_sha: joined log g of 2-D 32 x 32, 8 windows x 2 walkers dealt over the ranks, identical at every rank count; "
                              "energies_sha: the timed problem (its window count grows with the ranks: comparable at equal N only)",
                      "fixed_problem_sha": fixed_sha, "energies_sha": _sha(energies.astype(np.int64))}}
    wl.close()
    if cpu and rank == 0:
        try:
            out["cpu_baseline"] = cpu_leg_c5(min(_threads(), 32))
        except Exception as e:
            out["cpu_baseline"] = {"error": repr(e)}
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
