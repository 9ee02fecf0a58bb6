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
    return ap.parse_args()


def workload_config(args, world):
    return {
        "workload": "2D Ising L=%d single-chain checkerboard %s at beta_c (BASELINE.json configs[1])" % (args.L, args.rule),
        "L": args.L, "beta": BETA_C, "rule": args.rule, "storage": "int8 (1 byte/spin, 2 colour planes)",
        "chains_per_gpu": 1, "sweeps_per_step": args.sweeps_per_step,
        "rng": "Philox4x32-10, 32-bit draws (16-bit high half + lazy low half)",
        "l2_policy": "input (%d MiB) larger than L2 (126 MB); no flush needed" % (args.L * args.L >> 20),
        "parallelism": "replicas only: %d independent chain(s), one per GPU" % world,
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
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

    L, S = args.L, args.sweeps_per_step
    N = L * L
    rule = {"metropolis": 0, "glauber": 1, "heatbath": 2}[args.rule]
    sys_ = m.Ising([L, L], ctx=ctx)
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
    bytes_per_launch = 3 * (N // 2)                                   # 3 B/attempt x N/2 attempts per half-sweep
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    traffic = ncu_traffic()
    roofline = {"bound": "hbm", "kernel": "k_ising2d (half-sweep)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src, "bytes_per_attempt": 3,
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
           "serial": {"value": world * args.steps * S * N / (serial_s * 1e9), "ms_per_step": serial_s / args.steps * 1e3,
                      "api": "mcx_lattice_upload + mcx_sweep + mcx_observables, nothing overlapped"}}

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
    return {"metric": "pt_sweeps_per_s", "value": sweeps / (ms * 1e-3), "unit": "PT sweeps/s (all %d replicas swept once)" % n,
            "attempts_per_ns": sweeps * n * L * L / (ms * 1e6), "L": L, "replicas": n, "exchange_every": every,
            "rounds": rounds, "scaling": "strong", "exchange_acceptance": pt.acceptance_rate(),
            "collective": "none (1 rank)" if world == 1 else
                          ("peer stores of %d doubles per exchange over NVLink (CUDA IPC), device-side arrival counters, no collective call" % n
                           if getattr(pt, "_peers", False) else "NCCL all-gather of %d doubles per exchange" % n)}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
