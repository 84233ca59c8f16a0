#!/usr/bin/env python
"""bench.py -- headline benchmark: kinetic-VQMC walker-steps/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one `val_and_grad(qmc::KineticVQMC, p)` evaluation (wrapper.jl:90-97): every walker
advances STEPS_PER_CALL continuous-time steps, the estimators are reduced and all-reduced.
Workload: HubbardReal1D BoseFS{50,50} u=2.0, GutzwillerAnsatz g=1.0 (the shape the north-star
target is quoted on), 2^24 walkers in total sharded over the N GPUs (strong scaling).

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference
(oracle/, all host cores) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
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
    line = {"impl": "reference", "metric": "kinetic-VQMC walker-steps/s", "value": rate, "unit": "walker-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU restatement (C, OpenMP) of Gutzwiller.jl+Rimu.jl semantics; "
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
    if rank != 0:
        return
    clocks["window"] = ("warm-up + device-timed + end-to-end regions" +
                        (f" + {extra_calls} untimed calls of the same workload" if extra_calls else ""))

    total_steps = TOTAL_WALKERS * STEPS_PER_CALL * args.steps
    value = total_steps / (ms * 1e-3)
    e2e = total_steps / (e2e_ms * 1e-3)
    local_walkers = wh.n
    W = wh.model.words
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
    inst_per_step = _inst_per_walker_step()
    issue = None
    if inst_per_step:
        ach = steps_per_s_gpu / 32.0 * inst_per_step  # warp-instructions/s (32 walkers per warp)
        issue = {"bound": "issue", "achieved": ach / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-inst/s",
                 "frac": ach / issue_peak, "inst_per_walker_step": inst_per_step,
                 "source": "profiles/ (ncu smsp__inst_executed.sum / walker-steps)"}

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
    line = {
        "metric": "kinetic-VQMC walker-steps/s", "value": value, "unit": "walker-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "walkers_per_gpu": local_walkers, "steps_per_call": STEPS_PER_CALL,
                   "kernel": "walker_incr_kernel: occupation bit-planes + incremental bond-class counts (GWZ_KERNEL_BITPARALLEL)", "parallelism": f"walker-shard x{args.gpus}",
                   "l2": f"walker state {alg_bytes / 2 / 2**20:.0f} MiB per GPU > 126 MB L2 (inputs larger than L2)"
                   if alg_bytes / 2 > 126e6 else "walker state fits in L2; kernel is register-resident (not HBM bound)"},
        "e2e": {"value": e2e, "unit": "walker-steps/s", "h2d_bytes_per_step": 8 + 1040, "d2h_bytes_per_step": 8 * 12 + 4,
                "note": "val_and_grad(qmc, p) from host params to host (val, grad); inputs are one double plus the "
                        "kernel-argument tables, outputs the 12-double accumulator vector"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                     "traffic": _dram_traffic(local_walkers), "peak_source": peak_src,
                     "note": "walker kernel keeps its state in registers: HBM traffic is one load+store of the walker "
                             "state per launch; the binding limit is instruction issue, see issue_roofline"},
        "issue_roofline": issue,
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
        "energy": {"val": last.val, "err": last.err, "pooled_val": last.pooled_val, "grad": last.grad[0]},
        "device": {"sm_count": info["sm_count"], "cc": list(info["cc"])},
    }
    if args.extra:
        line["extra"] = extra_workloads(gw, np)
    print(json.dumps(line), flush=True)


def _dram_traffic(walkers):
    """dram__bytes_read.sum + dram__bytes_write.sum of one walker-kernel launch, from the committed ncu --set full
    capture (taken on 2^22 walkers of the same shape), scaled to this launch's walker count."""
    p = os.path.join(ROOT, "profiles", "walker_inst_per_step.json")
    try:
        return float(json.load(open(p))[f"dram_bytes_per_walker_N{N_PART}M{N_MODES}"]) * walkers
    except Exception:
        return None


def _inst_per_walker_step():
    """SASS instructions executed per walker-step for the benchmarked kernel, measured once with ncu
    and committed under profiles/ (bench.py cannot run ncu on itself)."""
    p = os.path.join(ROOT, "profiles", "walker_inst_per_step.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))[f"N{N_PART}M{N_MODES}g{G}"])
        except Exception:
            return None
    return None


def extra_workloads(gw, np):
    """Secondary numbers (not the headline): BASELINE configs 2, 3 and 5 on one GPU."""
    out = {}
    # config 3: KineticVQMC samples=1e8, 2^20 walkers, BoseFS{20,20} u=4
    H = gw.HubbardReal1D(gw.near_uniform(20, 20), u=4.0)
    q = gw.KineticVQMC(H, gw.GutzwillerAnsatz(H), samples=1e8, walkers=1 << 20, record=False)
    for _ in range(3):
        gw.val_and_grad(q, 0.25)
    t0 = time.perf_counter()
    for _ in range(5):
        gw.val_and_grad(q, 0.25)
    dt = (time.perf_counter() - t0) / 5
    out["cfg3_bose20_2^20walkers_95steps"] = {"walker_steps_per_s": q.steps * q.walkers / dt, "ms_per_call": dt * 1e3,
                                               "kernel_ms": q.result.last.kernel_ms}
    # config 5 shape on one GPU: BoseFS{100,100}
    H = gw.HubbardReal1D(gw.near_uniform(100, 100), u=2.0)
    q = gw.KineticVQMC(H, gw.GutzwillerAnsatz(H), walkers=1 << 21, steps=64, record=False)
    for _ in range(3):
        gw.val_and_grad(q, 1.0)
    t0 = time.perf_counter()
    for _ in range(5):
        gw.val_and_grad(q, 1.0)
    dt = (time.perf_counter() - t0) / 5
    out["cfg5_bose100_2^21walkers_64steps"] = {"walker_steps_per_s": q.steps * q.walkers / dt, "ms_per_call": dt * 1e3}
    # config 2: LocalEnergyEvaluator sweep BoseFS{12,12} and sustained BoseFS{16,16}
    for n in (12, 16):
        H = gw.HubbardReal1D(gw.near_uniform(n, n), u=2.0)
        le = gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H))
        for _ in range(64 if n == 12 else 12):  # past the call at which the handle stores its representative basis
            le.val_and_grad(1.0)
        reps = 50 if n == 12 else 5
        t0 = time.perf_counter()
        for _ in range(reps):
            v, g = le.val_and_grad(1.0)
        dt = (time.perf_counter() - t0) / reps
        out[f"cfg2_sweep_bose{n}"] = {"states": le.dim, "states_per_s": le.dim / dt, "ms_per_sweep": dt * 1e3,
                                      "kernel_ms": le.last_kernel_ms(), "val": v, "grad": float(g[0]),
                                      "note": "repeated sweeps of one LocalEnergyEvaluator evaluate its stored orbit "
                                              "representatives (about dim/2M states, weighted with the orbit sizes); "
                                              "states_per_s counts all states of the space"}
        le1 = gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H))  # a fresh handle: its first sweep enumerates the ranks
        t0 = time.perf_counter()
        le1.val_and_grad(1.0)
        out[f"cfg2_sweep_bose{n}"]["first_sweep_ms"] = (time.perf_counter() - t0) * 1e3
        out[f"cfg2_sweep_bose{n}"]["first_sweep_kernel_ms"] = le1.last_kernel_ms()
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
    ap.add_argument("--extra", action="store_true", help="also time BASELINE configs 2, 3, 5 (secondary numbers)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
