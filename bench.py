#!/usr/bin/env python
"""bench.py -- headline benchmark: kinetic-VQMC walker-steps/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one `val_and_grad(qmc::KineticVQMC, p)` evaluation (wrapper.jl:90-97): every walker
advances STEPS_PER_CALL continuous-time steps, the estimators are reduced and all-reduced.
Workload: HubbardReal1D BoseFS{50,50} u=2.0, GutzwillerAnsatz g=1.0 (the shape the north-star
target is quoted on), 2^24 walkers in total sharded over the N GPUs (strong scaling).

Prints ONE JSON line (rank 0).  Besides the headline it carries a `secondary` block measured in the same run: the
second half of BASELINE.json's metric (basis-sweep states/s, each rate labelled with what was evaluated), BASELINE
configs 3 and 5, the headline shape at g = 0.33, and val_err_and_grad with the device-side blocking kernel.
`--impl reference` times the CPU restatement of the reference (oracle/, all host cores) on a bounded sample of the
same workload.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PART, N_MODES, U, T, G = 50, 50, 2.0, 1.0, 1.0
TOTAL_WALKERS = 1 << 24
STEPS_PER_CALL = 64
SEED = 0x9E3779B97F4A7C15
WORKLOAD = (f"KineticVQMC val_and_grad: HubbardReal1D BoseFS{{{N_PART},{N_MODES}}} u={U} t={T}, GutzwillerAnsatz g={G}, "
            f"{TOTAL_WALKERS} walkers (2^24) x {STEPS_PER_CALL} steps per call")


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_rate(target_seconds: float = 12.0):
    """Reference arm: the CPU restatement of kinetic_vqmc! (oracle/, OpenMP over walkers like
    Threads.@threads, qmc.jl:102) on all host cores, 2*cores walkers (qmc.jl:132), bounded sample."""
    import numpy as np
    from oracle.oracle import Oracle
    cores = os.cpu_count() or 1
    o = Oracle(N_PART, N_MODES, U, T)
    walkers = 2 * cores
    start = np.stack([o.near_uniform()] * walkers)
    # calibrate
    t0 = time.perf_counter()
    o.kinetic_vqmc(G, start.copy(), 50, seed=SEED, n_threads=cores)
    dt = max(time.perf_counter() - t0, 1e-3)
    steps = max(50, int(50 * target_seconds / dt))
    t0 = time.perf_counter()
    out = o.kinetic_vqmc(G, start.copy(), steps, seed=SEED, n_threads=cores)
    dt = time.perf_counter() - t0
    # the tuned CPU code (closed-form ratios, tables; NOT the reference algorithm): upper bound of the CPU side
    fsteps = steps * 40
    t0 = time.perf_counter()
    o.kinetic_vqmc_fast(G, start.copy(), fsteps, seed=SEED, n_threads=cores)
    fdt = time.perf_counter() - t0
    return dict(rate=walkers * steps / dt, cores=cores, walkers=walkers, steps=steps, seconds=dt,
                hops_per_step=out["hops"] / (walkers * steps), val=out["val"],
                tuned_rate=walkers * fsteps / fdt, tuned_seconds=fdt, tuned_steps=fsteps)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per = []
    info = None
    for i in range(args.warmup + args.steps):
        info = cpu_reference_rate(target_seconds=max(2.0, 24.0 / max(1, args.steps + args.warmup)))
        if i >= args.warmup:
            per.append(info)
    rate = sum(p["walkers"] * p["steps"] for p in per) / sum(p["seconds"] for p in per)
    ms = 1e3 * sum(p["seconds"] for p in per) / len(per)
    sample = (f"{info['walkers']} walkers (2 x cores, qmc.jl:132) x {info['steps']} steps of the same "
              f"BoseFS{{{N_PART},{N_MODES}}} workload per step")
    ref_workload = (f"kinetic_vqmc!: HubbardReal1D BoseFS{{{N_PART},{N_MODES}}} u={U} t={T}, GutzwillerAnsatz g={G}, "
                    f"{info['walkers']} walkers (2 x cores, the reference default, qmc.jl:132) x {info['steps']} steps per "
                    f"step -- a bounded sample of the GPU arm's workload ({TOTAL_WALKERS} walkers x {STEPS_PER_CALL} steps)")
    line = {"impl": "reference", "metric": "kinetic-VQMC walker-steps/s", "value": rate, "unit": "walker-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": ref_workload, "gpu_arm_workload": WORKLOAD,
                       "note": "CPU restatement (C, OpenMP) of Gutzwiller.jl+Rimu.jl semantics; "
                               "Julia is not installed in this image"},
            "cpu_baseline": {"value": rate, "unit": "walker-steps/s", "cores": info["cores"], "kind": "port",
                             "sample": sample},
            "e2e": {"value": rate, "unit": "walker-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "hops_per_step": info["hops_per_step"], "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import gutzwiller_b200 as gw

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="gloo")
    gw.init_distributed()
    torch.cuda.set_device(local)
    ctx = gw.default_context()
    info = ctx.info()
    stream = torch.cuda.ExternalStream(ctx.stream(), device=local)

    H = gw.HubbardReal1D(gw.near_uniform(N_PART, N_MODES), u=U, t=T)
    ans = gw.GutzwillerAnsatz(H)
    qmc = gw.KineticVQMC(H, ans, walkers=TOTAL_WALKERS, steps=STEPS_PER_CALL, seed=SEED, record=False)
    wh = qmc.result.walkers_handle
    p = np.array([G])

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # nvidia-smi needs ~0.1 s to start and samples every 0.1 s: it runs from the warm-up through both timed regions
    # (and, for very short runs, some untimed calls of the same workload) so that it always sees the GPU under load
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_load = time.perf_counter()

    # ---- warm-up (also thermalises the walkers) ----
    for _ in range(max(args.warmup, 3)):
        last = wh.run(p, STEPS_PER_CALL)

    # ---- device-resident throughput: K calls, CUDA events on the launching stream ----
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    launches0 = gw.launch_count()
    kernel_ms = []
    hops = 0
    for _ in range(args.steps):
        last = wh.run(p, STEPS_PER_CALL)
        kernel_ms.append(last.kernel_ms)
        hops += last.hops
    ev1.record(stream)
    launches = gw.launch_count() - launches0
    barrier()
    ms = ev0.elapsed_time(ev1)

    # ---- end to end through the public API (host params in, host (val, grad) out), wall clock ----
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        val, grad = gw.val_and_grad(qmc, p)
    ctx.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    tms = torch.tensor([ms, e2e_s * 1e3, max(kernel_ms), sum(kernel_ms) / len(kernel_ms)], dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms, e2e_ms, kmax, kavg = tms.tolist()
    # keep the same load up until the sampler has had 0.6 s (same number of extra calls on every rank)
    window = torch.tensor([time.perf_counter() - t_load], dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(window, op=dist.ReduceOp.MIN)
    extra_calls = 0
    if window.item() < 0.6:
        extra_calls = int((0.6 - window.item()) / (ms * 1e-3 / args.steps)) + 1
        for _ in range(extra_calls):
            wh.run(p, STEPS_PER_CALL)
        barrier()
    clocks = sampler.stop() if rank == 0 else None
    secondary = None if args.no_secondary else secondary_workloads(gw, np, ctx, world, rank, quick=args.quick)
    shard_check = shard_invariance_check(gw, np, world, rank) if world > 1 else None
    if rank != 0:
        return
    clocks["window"] = ("warm-up + device-timed + end-to-end regions" +
                        (f" + {extra_calls} untimed calls of the same workload" if extra_calls else ""))

    total_steps = TOTAL_WALKERS * STEPS_PER_CALL * args.steps
    value = total_steps / (ms * 1e-3)
    e2e = total_steps / (e2e_ms * 1e-3)
    local_walkers = wh.n
    # algorithmic HBM bytes of one walker-kernel launch: the resident walker state (occupation bit-planes:
    # NP planes x ceil(M/32) words) and the 5 running sums, loaded and stored once per launch
    n_planes = 4 if N_PART <= 15 else (6 if N_PART <= 63 else 8)
    state_bytes = 4 * n_planes * ((N_MODES + 31) // 32)
    # (val_and_grad(qmc, p) starts a new estimator, so the sums are written but not read)
    alg_bytes = local_walkers * (2 * state_bytes + 8 * 5)
    hbm_peak, peak_src = _peaks()
    hbm_ach = alg_bytes / (kavg * 1e-3) / 1e9
    hops_per_step = hops / total_steps
    steps_per_s_gpu = local_walkers * STEPS_PER_CALL / (kavg * 1e-3)
    sm_mhz = (clocks or {}).get("sm_mhz") or info["sm_clock_khz"] / 1e3
    issue_peak = info["sm_count"] * 4 * sm_mhz * 1e6  # warp instructions / s at the clock seen under load
    prof = _kernel_profile(f"N{N_PART}M{N_MODES}g{G}")
    inst_per_step = prof.get("inst_per_walker_step")
    roofline = {"bound": "issue", "achieved": None, "peak": issue_peak / 1e9, "unit": "G warp-inst/s", "frac": None,
                "traffic": None, "inst_per_walker_step": inst_per_step, "kernel_source_sha256": prof["sha"],
                "profile": prof.get("profile"), "ncu_issue_active_pct": prof.get("ncu_issue_active_pct"),
                "ncu_shared_pipe_pct": prof.get("ncu_shared_pipe_pct"),
                "note": "the walker kernel is register / shared-memory resident: what binds it is instruction issue (4 "
                        "warp-instructions per SM and clock), not HBM (see hbm_roofline) and not tensor cores.  achieved = "
                        "walker-steps/s of the timed kernel / 32 x SASS instructions per walker-step; the instruction "
                        "count comes from the committed ncu capture of THIS kernel source (matched by hash; null when "
                        "the source has changed since the capture); ncu_* = issue-slot and shared-memory data-pipe "
                        "utilisation of that capture (the second bound of this kernel)"}
    if inst_per_step:
        ach = steps_per_s_gpu / 32.0 * inst_per_step  # warp-instructions/s (32 walkers per warp)
        roofline.update(achieved=ach / 1e9, frac=ach / issue_peak)

    cpu = None
    if args.gpus == 1 and not args.no_cpu_baseline:
        c = cpu_reference_rate(12.0)
        cpu = {"value": c["rate"], "unit": "walker-steps/s", "cores": c["cores"], "kind": "port",
               "sample": f"{c['walkers']} walkers x {c['steps']} steps of the same BoseFS{{{N_PART},{N_MODES}}} workload "
                         f"({c['seconds']:.1f} s, C/OpenMP restatement of the reference)"}

    cpu_tuned = None
    if cpu is not None:
        cpu_tuned = {"value": c["tuned_rate"], "unit": "walker-steps/s", "cores": c["cores"], "kind": "tuned-cpu",
                     "sample": f"{c['walkers']} walkers x {c['tuned_steps']} steps ({c['tuned_seconds']:.1f} s)",
                     "note": "oracle.kinetic_vqmc_fast: closed-form ratios and tables on occupation arrays, the same "
                             "trajectories as the port; an upper bound of the CPU side, not the reference's algorithm"}
    dram = prof.get("dram_bytes_per_walker")
    line = {
        "metric": "kinetic-VQMC walker-steps/s", "value": value, "unit": "walker-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "walkers_per_gpu": local_walkers, "steps_per_call": STEPS_PER_CALL,
                   "kernel": KERNEL_NAME, "parallelism": f"walker-shard x{args.gpus}",
                   "l2": f"walker state {alg_bytes / 2 / 2**20:.0f} MiB per GPU > 126 MB L2 (inputs larger than L2)"
                   if alg_bytes / 2 > 126e6 else "walker state fits in L2; kernel is register-resident (not HBM bound)"},
        "e2e": {"value": e2e, "unit": "walker-steps/s", "h2d_bytes_per_step": 8 + 1040, "d2h_bytes_per_step": 8 * 13 + 4,
                "note": "val_and_grad(qmc, p) from host params to host (val, grad); inputs are one double plus the "
                        "kernel-argument tables, outputs the 13-double accumulator vector"},
        "gpu_launches": launches,
        "roofline": roofline,
        "hbm_roofline": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                         "traffic": dram * local_walkers if dram else None, "peak_source": peak_src,
                         "note": "not the binding resource: HBM traffic is one load + store of the walker state per "
                                 "launch (algorithmic bytes = planes read + planes and sums written)"},
        "fp64_algorithmic": {"flops_per_walker_step": 3.0 * hops_per_step + 12.0,
                             "achieved_tflops": steps_per_s_gpu * (3.0 * hops_per_step + 12.0) / 1e12,
                             "peak_tflops": info["sm_count"] * 64 * 2 * sm_mhz * 1e6 / 1e12,
                             "note": "algorithmic FP64 work of the reference formulation (SURVEY 8d: 3h+12 per step); the kernel "
                                     "replaces most of it by integer class counts and tables, so this is a work-equivalent "
                                     "rate, not executed DFMA"},
        "cpu_baseline": cpu,
        "cpu_tuned": cpu_tuned,
        "clocks": clocks,
        "hops_per_step": hops_per_step, "hop_evals_per_s": value * hops_per_step,
        "kernel_ms_avg": kavg, "kernel_ms_max": kmax,
        "energy": {"val": last.val, "pooled_val": last.pooled_val, "grad": last.grad[0],
                   "cross_walker_se": last.err,
                   "note": f"mean of per-walker ratio estimates (state.jl:137-149) at {STEPS_PER_CALL} steps per walker: "
                           "each walker carries an O(tau_corr/steps) ratio bias that does not average out over walkers, "
                           "so `val` sits away from `pooled_val` by far more than cross_walker_se, which only measures "
                           "the scatter between walkers.  This is the reference's estimator at this shape, not an "
                           "accuracy statement; see secondary.blocking for the reference's own error bar"},
        "device": {"sm_count": info["sm_count"], "cc": list(info["cc"])},
        "secondary": secondary,
    }
    if shard_check is not None:
        line["shard_check"] = shard_check
    if args.extra:
        line["extra"] = extra_workloads(gw, np)
    print(json.dumps(line), flush=True)


KERNEL_NAME = ("walker_fast_kernel: integer hop counts per rate class, hop masks per (rate class, direction) in shared "
               "memory, table-driven updates (GWZ_KERNEL_BITPARALLEL)")
KERNEL_SOURCES = ("gwz_device.cuh", "gwz_planes.cuh", "gwz_incr.cuh", "gwz_fast.cuh", "gwz_walker_fast.cu")


def kernel_source_hash():
    """sha256 over the sources of the production walker kernel: ties a committed ncu capture to the code it profiled"""
    h = hashlib.sha256()
    for name in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "gutzwiller.jl_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _kernel_profile(key):
    """ncu-derived constants of the production kernel (SASS instructions per walker-step, DRAM bytes per walker) from
    profiles/walker_inst_per_step.json -- only if that file was recorded for the kernel source being timed."""
    sha = kernel_source_hash()
    out = {"sha": sha}
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "walker_inst_per_step.json")))
        if d.get("source_sha256") == sha and key in d.get("entries", {}):
            out.update(d["entries"][key])
    except Exception:
        pass
    return out


def secondary_workloads(gw, np, ctx, world, rank, quick=False):
    """The rest of the metric and of BASELINE.json's configs, measured in the same run (every rank executes it: the
    sampling and sweep calls are collective; rank 0 reports).  Wall-clock per API call (each call blocks until its
    result is on the host) next to the device time of its kernel."""
    out = {}

    def timed(fn, reps, warm=1):
        for _ in range(warm):
            fn()
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        return (time.perf_counter() - t0) / reps, r

    # ---- basis-sweep states/s (BASELINE config 2 + sustained size): three different things, labelled ----
    for n in (12, 16):
        H = gw.HubbardReal1D(gw.near_uniform(n, n), u=2.0)
        ans = gw.GutzwillerAnsatz(H)
        dim = int(round(H.model().dimension()))
        entry = {}
        # (i) state by state: every state of the space is evaluated
        os.environ["GWZ_SWEEP_SYM"] = "0"
        le = gw.LocalEnergyEvaluator(H, ans)
        dt, (v, g) = timed(lambda: le.val_and_grad(1.0), 3 if n == 16 else 20)
        entry["state_by_state"] = {"states_evaluated": dim, "ms_per_sweep": dt * 1e3, "kernel_ms": le.last_kernel_ms(),
                                   "states_per_s": dim / dt, "val": v, "grad": float(g[0])}
        del le
        os.environ.pop("GWZ_SWEEP_SYM")
        # (ii) first sweeps of a handle: the rank range is enumerated, one representative per orbit of the chain's
        # translations + reflection is evaluated (weighted with the orbit size)
        os.environ["GWZ_SWEEP_CACHE"] = "0"
        le = gw.LocalEnergyEvaluator(H, ans)
        dt, (v2, g2) = timed(lambda: le.val_and_grad(1.0), 5 if n == 16 else 20)
        entry["orbit_reduced_enumerating"] = {"states_covered": dim, "states_evaluated": "about dim/(2M): one per orbit",
                                              "ms_per_sweep": dt * 1e3, "kernel_ms": le.last_kernel_ms(),
                                              "states_covered_per_s": dim / dt}
        del le
        os.environ.pop("GWZ_SWEEP_CACHE")
        # (iii) repeated sweeps of one LocalEnergyEvaluator (what AMSGrad does 200 times): the representatives are
        # stored once, later sweeps evaluate only them
        le = gw.LocalEnergyEvaluator(H, ans)
        for _ in range(64 if n == 12 else 12):  # past the call at which the handle stores its representative basis
            le.val_and_grad(1.0)
        dt, (v3, g3) = timed(lambda: le.val_and_grad(1.0), 10 if n == 16 else 50)
        entry["cached_representatives"] = {"states_covered": dim, "states_evaluated": "about dim/(2M): stored representatives",
                                           "ms_per_sweep": dt * 1e3, "kernel_ms": le.last_kernel_ms(),
                                           "states_covered_per_s": dim / dt}
        entry["agreement"] = {"orbit_vs_state_by_state": abs(v2 - v) / abs(v), "cached_vs_state_by_state": abs(v3 - v) / abs(v),
                              "grad_cached_vs_state_by_state": abs(float(g3[0]) - float(g[0])) / abs(float(g[0]))}
        del le
        out[f"sweep_bose{n}"] = entry
        if quick:
            break

    # ---- walker-steps/s on the other shapes ----
    def walk(N, M, u, g, walkers, steps, reps=3):
        H = gw.HubbardReal1D(gw.near_uniform(N, M), u=u)
        q = gw.KineticVQMC(H, gw.GutzwillerAnsatz(H), walkers=walkers, steps=steps, record=False)
        dt, _ = timed(lambda: gw.val_and_grad(q, g), reps, warm=2)
        r = q.result.last
        return {"n_gpus": world, "walkers_total": walkers, "steps_per_call": steps, "ms_per_call": dt * 1e3,
                "kernel_ms": r.kernel_ms, "walker_steps_per_s": walkers * steps / dt,
                "walker_steps_per_s_kernel": walkers * steps / (r.kernel_ms * 1e-3) if world == 1 else None,
                "hops_per_step": r.hops / r.steps, "val": r.val, "pooled_val": r.pooled_val}

    out["cfg3_bose20_u4_g0.25_2^20walkers_95steps"] = walk(20, 20, 4.0, 0.25, 1 << 20, 95, reps=5)
    out["headline_shape_g0.33"] = walk(N_PART, N_MODES, U, 0.33, TOTAL_WALKERS if not quick else 1 << 20, STEPS_PER_CALL)
    for g in (1.0, 0.33):
        out[f"cfg5_bose100_u2_g{g}_2^24walkers_sharded_1000steps"] = walk(100, 100, 2.0, g, (1 << 24) if not quick else 1 << 18,
                                                                        1000, reps=2)
        if quick:
            break

    # ---- val_err_and_grad on the headline shape: series kept in HBM + the blocking kernel (a8) ----
    H = gw.HubbardReal1D(gw.near_uniform(N_PART, N_MODES), u=U, t=T)
    nw = (1 << 22) if not quick else 1 << 16
    q = gw.KineticVQMC(H, gw.GutzwillerAnsatz(H), walkers=nw, steps=256, record=False)
    gw.val_and_grad(q, G)
    dt_plain, _ = timed(lambda: gw.val_and_grad(q, G), 2)
    k_plain = q.result.last.kernel_ms
    dt_err, (v, e, g) = timed(lambda: gw.val_err_and_grad(q, G), 2)
    b = q.result.blocking_summary
    local = q.result.walkers_handle.n
    hbm_peak, _ = _peaks()
    alg = local * (16.0 * 256 + 40.0 + 40.0)  # series read once + running sums read + 5 doubles per walker written
    out["blocking"] = {
        "workload": f"val_err_and_grad(qmc, p): BoseFS{{{N_PART},{N_MODES}}} g={G}, {nw} walkers x 256 steps, per-walker "
                    "resample + blocking analysis (state.jl:50-60,150-164) on the device",
        "ms_per_call_val_and_grad": dt_plain * 1e3, "ms_per_call_val_err_and_grad": dt_err * 1e3,
        "walker_kernel_ms": k_plain, "walker_kernel_ms_recording": q.result.last.kernel_ms,
        "blocking_kernel_ms": b.kernel_ms, "val": v, "err": e, "err_err": b.err_err, "failed_walkers": int(b.failed),
        "k_range": [int(b.k_min), int(b.k_max)],
        "roofline": {"bound": "hbm", "achieved": alg / (b.kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": alg / (b.kernel_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                     "note": "series_blocking_kernel: algorithmic bytes = 16 per recorded walker-step + 80 per walker"}}
    return out


def shard_invariance_check(gw, np, world, rank):
    """N > 1: the NCCL-combined result of 2^16 walkers sharded over the N ranks against the same walkers run as N
    logical shards on rank 0's GPU alone (second context, no communicator)."""
    H = gw.HubbardReal1D(gw.near_uniform(N_PART, N_MODES), u=U, t=T)
    total, steps = 1 << 16, 32
    q = gw.KineticVQMC(H, gw.GutzwillerAnsatz(H), walkers=total, steps=steps, seed=SEED + 1, record=False)
    gw.val_and_grad(q, G)
    r = q.result.last
    acc_nccl = np.array(r.acc[:])
    if rank != 0:
        return None
    ctx2 = gw.Context(int(os.environ.get("LOCAL_RANK", "0")))
    model2 = gw.Model(N_PART, N_MODES, U, T, ctx=ctx2)
    sets = []
    for k in range(world):
        b, e = gw.shard_range(total, world, k)
        sets.append(gw.Walkers(model2, e - b, H.address.words, seed=SEED + 1, global_offset=b))
    r2 = gw.run_group(sets, [G], steps)
    acc_one = np.array(r2.acc[:])
    rel = float(np.max(np.abs(acc_nccl - acc_one) / np.maximum(np.abs(acc_one), 1e-300)))
    return {"walkers": total, "steps": steps, "logical_shards_on_rank0": world, "max_rel_diff_of_accumulators": rel,
            "val_nccl": r.val, "val_single_gpu": r2.val,
            "pass": bool(rel < 1e-12 and abs(r.val - r2.val) <= 1e-12 * abs(r2.val))}


def extra_workloads(gw, np):
    """Numbers outside the north-star path: the ansatz-generic engine and the stored-operator path."""
    out = {}
    # ansatz-generic engine: RelativeJastrowAnsatz, 25 parameters, BoseFS{50,50}
    H = gw.HubbardReal1D(gw.near_uniform(50, 50), u=2.0)
    ans = gw.RelativeJastrowAnsatz(H)
    p = np.zeros(ans.N_PARAMS)
    p[0], p[1:] = 0.33, 0.005
    q = gw.KineticVQMC(H, ans, walkers=1 << 16, steps=64, record=False)
    for _ in range(3):
        gw.val_and_grad(q, p)
    t0 = time.perf_counter()
    for _ in range(5):
        gw.val_and_grad(q, p)
    dt = (time.perf_counter() - t0) / 5
    out["generic_relative_jastrow_bose50_2^16walkers_64steps"] = {
        "walker_steps_per_s": q.steps * q.walkers / dt, "ms_per_call": dt * 1e3, "kernel_ms": q.result.last.kernel_ms,
        "parameters": ans.N_PARAMS}
    # stored-operator path: synthetic operator, dim 1e6 x 16 entries per row, GutzwillerAnsatz table
    rng = np.random.default_rng(0)
    dim, L = 1_000_000, 16
    Hs = gw.SparseHamiltonian(np.arange(dim + 1, dtype=np.uint64) * np.uint64(L),
                              rng.integers(0, dim, size=dim * L, dtype=np.uint32), -rng.random(dim * L), rng.normal(size=dim))
    q = gw.KineticVQMC(Hs, gw.SparseGutzwillerAnsatz(Hs), walkers=1 << 18, steps=200, record=False)
    for _ in range(3):
        gw.val_and_grad(q, 0.1)
    t0 = time.perf_counter()
    for _ in range(5):
        gw.val_and_grad(q, 0.1)
    dt = (time.perf_counter() - t0) / 5
    out["stored_operator_dim1e6_16perrow_2^18walkers_200steps"] = {
        "walker_steps_per_s": q.steps * q.walkers / dt, "ms_per_call": dt * 1e3, "kernel_ms": q.result.last.kernel_ms}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extra", action="store_true", help="also time the generic engine and the stored-operator path")
    ap.add_argument("--no-secondary", action="store_true", help="headline only (skip the secondary block)")
    ap.add_argument("--quick", action="store_true", help="smaller secondary workloads (smoke runs of bench.py itself)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
