#!/usr/bin/env python3
"""bench.py -- closed-loop tuning candidates/s (and QP solves/s) on Shell3x3, BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pop 4096] [--mode gam|vns]

One "step" = one pass of the hot path over one synthetic population (Shell3x3,
nit=500, 4096 seeded candidates per GPU): builder kernels + closed-loop kernels + cost reduction.
  value : candidates/s with the population already resident in HBM (CUDA events, max over ranks)
  e2e   : candidates/s through the reference-facing call (host arrays -> mpcgpu_eval_batch -> host costs)
  roofline / cpu_baseline : see DESIGN.md "Measurement".
Under torchrun (N>1) each rank owns one GPU and one shard of the population (weak scaling: per-GPU work
fixed); the only collective is the per-generation all-gather of fitness over NCCL.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "model-predictive-control-tuning_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

def ncu_traffic(summary_file, kernel_substr, grid):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, read from the committed `ncu --set full` summary
    under profiles/ (profiles/summarize.py) -- only if that capture had the same grid (same population); else None."""
    try:
        for o in json.load(open(os.path.join(ROOT, "profiles", summary_file))):
            if kernel_substr in o.get("Kernel Name", "") and o.get("Grid Size", "").startswith("(%d," % grid):
                tot = 0.0
                for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    v, u = o[k].split()[:2]
                    tot += float(v) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
                return tot
    except Exception:
        pass
    return None


METRIC = "closed-loop tuning candidates/sec (Shell3x3)"
UNIT = "candidates/s"


def linear_problem(args):
    import mpcgpu
    return mpcgpu.shell7x5() if args.config == "shell7x5" else mpcgpu.shell3x3(2)


def metric_name(args):
    return {"shell3x3": METRIC, "shell7x5": "closed-loop tuning candidates/sec (Shell7x5, soft output bands)",
            "dtc": "DTC-GPC sweep candidates/sec (Wood-Berry)", "vdv": "NMPC closed-loop tuning candidates/sec (Van de Vusse)",
            "ssnmpc": "single-shooting NMPC sweep candidates/sec (Explicit NMPC demo, Van de Vusse)"}[args.config]


def workload_string(mode, pop, config="shell3x3"):
    """config.workload, identical in both arms (the driver compares the strings)."""
    if config == "shell7x5":
        return (f"Shell7x5 {mode.upper()} closed-loop evaluation (BASELINE.json configs[2]): nit=200, population {pop} per GPU, "
                "N~U{7..127}, Nu~U{2..15}, lambda log-uniform [1e-4,10], delta = 0 (band control, soft output limits), PCG64 seed 0")
    return (f"Shell3x3 {mode.upper()} closed-loop evaluation (BASELINE.json configs[1]): nit=500, population {pop} per GPU, "
            "N~U{7..127}, Nu~U{2..15}, weights log-uniform [1e-4,10], PCG64 seed 0; one candidate = 500 closed-loop QPs + plant "
            "rollout + cost (GAM_fun.m:81 discards the open-loop optimum, so the GAM objective skips that 501st QP; the VNS "
            "objective keeps it)")


def shard_of(prob, Ng, Nug, dg, lg, world, rank):
    """The product's sharding (mpcgpu.distributed.evaluate_sharded / mpcgpu_create_multi): candidates sorted by the
    a-priori work estimate and dealt round-robin over the ranks."""
    from mpcgpu.distributed import shard_indices, work_estimate
    if world == 1:
        return np.arange(len(Ng))
    work = work_estimate(Ng, Nug, dg, lg, dead_max=int(prob.plant.d.max()), soft=bool(np.isfinite(prob.ymin).any() or np.isfinite(prob.ymax).any()))
    return shard_indices(len(Ng), world, rank, work)


def algorithmic_flops(prob, N, Nu, nit):
    """Executed-algorithm minimum: builder + nit unconstrained controller moves + plant.
    Active-set iterations are extra work and are NOT counted."""
    ny, nu, nw = prob.ny, prob.nu, prob.nu + prob.nd
    hl = np.maximum(prob.plant.d.max(axis=0), np.where(np.arange(nw) < nu, 1, 0))
    nst = ny * nw + int(hl.sum()) - nu + ny
    nz = nu * Nu.astype(np.float64)
    f_build = 2 * ny * nz ** 2 + nz ** 3 / 3 + 2 * nz ** 2 * (nst + nz) + 2 * ny * nz * nst
    f_step = 2 * nz * nst + 2 * nst + 8 * nz + 6 * ny * nw
    return float(np.sum(f_build + nit * f_step)), nst


def survey_flops(prob, N, Nu, nit):
    """SURVEY.md §8(d) F_cand of the dense formulation (for context only)."""
    ny, nu = prob.ny, prob.nu
    nz = nu * Nu.astype(np.float64); p = N.astype(np.float64)
    FH = 2 * ny * p * nz ** 2 + nz ** 3 / 3
    Fqp = 2 * ny * p * nz + 4 * ny * nu * p + 2 * nz ** 2
    return float(np.sum(FH + (nit + 1) * Fqp + 4 * ny * nu * nit))


class ClockSampler:
    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.stop = threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.samples.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=3)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 6 for i in range(4) if s[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def host_threads():
    """All host threads this process may use.  torchrun exports OMP_NUM_THREADS=1, which would silently make the
    reference arm single-threaded, so the count is taken from the affinity mask and passed to the oracle explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def other_configs():
    """One population of each of the other BASELINE.json configurations on this GPU, through the same public API
    (host arrays in and out, wall clock; after one warm-up call).  Context for the headline line, not part of it."""
    import mpcgpu
    out = {}
    try:
        p7 = mpcgpu.shell7x5(); e7 = mpcgpu.Evaluator(p7, device=0)
        P7 = mpcgpu.synthetic_population(p7, 2048, seed=0)
        e7.eval_batch(*P7, mode="gam")   # warm-up at full size (device buffers are allocated on first use)
        t0 = time.perf_counter(); o7 = e7.eval_batch(*P7, mode="gam"); dt = time.perf_counter() - t0
        out["shell7x5_soft_constraints"] = {"candidates": 2048, "candidates_per_s": 2048 / dt, "failed": int((o7["status"] != 0).sum()),
                                            "weights": "lambda log-uniform [1e-4, 10] (the survey's full range), delta = 0 (band control)"}
        e7.close()
        pd = mpcgpu.woodberry_dtc(); ed = mpcgpu.DtcEvaluator(pd, device=0)
        Pd = mpcgpu.synthetic_dtc_population(pd, 16384, seed=0)
        ed.eval_batch(*Pd[:4], alfa=Pd[4], raio=Pd[5])
        t0 = time.perf_counter(); od = ed.eval_batch(*Pd[:4], alfa=Pd[4], raio=Pd[5]); dt = time.perf_counter() - t0   # robustness filters designed on the device
        out["dtc_gpc_sweep"] = {"candidates": 16384, "candidates_per_s": 16384 / dt, "failed": int((od["status"] != 0).sum())}
        ed.close()
        pn = mpcgpu.vandevusse(); en = mpcgpu.NmpcEvaluator(pn, device=0)
        Pn = mpcgpu.synthetic_nmpc_population(pn, 16384, seed=0)
        en.eval_batch(*Pn, mode="gam")
        t0 = time.perf_counter(); on = en.eval_batch(*Pn, mode="gam"); dt = time.perf_counter() - t0
        out["vandevusse_nmpc"] = {"candidates": 16384, "candidates_per_s": 16384 / dt, "failed": int((~np.isin(on["status"], (0, 5))).sum())}
        en.close()
    except Exception as exc:   # context only: never fail the headline line
        out["error"] = repr(exc)
    try:   # single-shooting NMPC sweep (`Explicit NMPC/`): k_ssnmpc, one thread per closed loop of 150 samples
        ps = mpcgpu.explicit_nmpc(); es = mpcgpu.SsnmpcEvaluator(ps, device=0)
        Ps = mpcgpu.synthetic_ssnmpc_population(ps, 4096, seed=0)
        es.eval_batch(*[a[:64] for a in Ps])
        t0 = time.perf_counter(); os_ = es.eval_batch(*Ps); dt = time.perf_counter() - t0
        out["explicit_nmpc_single_shooting"] = {"candidates": 4096, "candidates_per_s": 4096 / dt, "failed": int((os_["status"] != 0).sum()),
                                                "controller_calls": es.counters()["qp_solves"]}
        es.close()
    except Exception as exc:
        out["explicit_nmpc_error"] = repr(exc)
    return out


def cpu_reference_rate(prob, N, Nu, delta, lam, mode, budget_s=12.0, nthreads=0):
    """Times the CPU oracle (the port of the reference's path) on a bounded sample of the same population."""
    from oracle import oracle as orc
    op = orc.OracleProblem(prob)
    if nthreads <= 0:
        nthreads = host_threads()
    cores = nthreads
    probe = min(len(N), max(2 * cores, 16))
    t0 = time.perf_counter()
    orc.eval_batch(op, N[:probe], Nu[:probe], delta[:probe], lam[:probe], mode, nthreads)
    dt = time.perf_counter() - t0
    n = int(min(len(N), max(probe, probe * budget_s / max(dt, 1e-3))))
    t0 = time.perf_counter()
    orc.eval_batch(op, N[:n], Nu[:n], delta[:n], lam[:n], mode, nthreads)
    dt = time.perf_counter() - t0
    return n / dt, cores, n


def linear_config(args, world, n, nit):
    """`config` of the bench line: ONE function for both arms, so that the driver's same-config check compares equal dicts.  The L2
    and multi-GPU entries describe how the GPU arm is measured; the reference arm (host CPU) carries them unchanged."""
    return {"workload": workload_string(args.mode, args.pop, args.config),
            "population_per_gpu": int(n), "nit": int(nit), "cost_mode": args.mode,
            "l2": "GPU arm: flushed between timed steps (256 MiB write, untimed); reference arm: host CPU, not applicable",
            "multi_gpu": ("population dealt over the ranks by estimated work (sorted round-robin, mpcgpu.distributed), "
                          "one NCCL all-gather of fitness per step") if world > 1 else "single GPU"}


def other_config(args, n):
    return {"workload": other_workload(args), "population_per_gpu": int(n),
            "value_is": "GPU arm: device time of the kernel (CUDA events inside the C ABI call, inputs resident); reference arm: wall clock",
            "l2": "GPU arm: flushed between timed steps (256 MiB write, untimed); reference arm: host CPU, not applicable"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  MATLAB + the closed-source MPC
    Toolbox cannot run here, so this is the oracle port on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import mpcgpu
    prob = linear_problem(args)
    world = max(1, args.gpus)
    Ng, Nug, dg, lg = mpcgpu.synthetic_population(prob, args.pop * world, seed=0)
    idx = shard_of(prob, Ng, Nug, dg, lg, world, 0)          # rank 0's shard of the GPU arm's population
    N, Nu, delta, lam = Ng[idx], Nug[idx], dg[idx], lg[idx]
    from oracle import oracle as orc
    op = orc.OracleProblem(prob)
    cores = host_threads()
    # bounded sample per step: ~ (2.5 s * cores) of CPU work
    t0 = time.perf_counter()
    probe = min(args.pop, max(2 * cores, 16))
    orc.eval_batch(op, N[:probe], Nu[:probe], delta[:probe], lam[:probe], args.mode, cores)
    dt = time.perf_counter() - t0
    n = int(min(args.pop, max(probe, probe * 2.5 / max(dt, 1e-3))))
    for _ in range(args.warmup):
        orc.eval_batch(op, N[:n], Nu[:n], delta[:n], lam[:n], args.mode, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.eval_batch(op, N[:n], Nu[:n], delta[:n], lam[:n], args.mode, cores)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    sample = f"first {n} of the {args.pop} candidates of rank 0's shard per step (oracle/mpc_oracle.c, OpenMP over candidates)"
    line = {"impl": "reference", "metric": metric_name(args), "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": linear_config(args, world, len(N), prob.nit),
            "reference_note": "CPU oracle port (the reference's MATLAB + MPC Toolbox cannot run here); " + sample,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "qp_solves_per_s": val * (500 if args.mode == "gam" else 3 * 501)}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------------
# The two configurations whose evaluator takes host arrays only (DTC-GPC sweep, Van de Vusse NMPC): same contract, one GPU
# per rank, candidates dealt round-robin (their cost does not depend on the weights the way the linear path's does).
# ------------------------------------------------------------------------------------------------------------------------
def other_workload(args):
    if args.config == "dtc":
        return (f"DTC-GPC sweep on Wood-Berry (BASELINE.json configs[3], DTC_GPC_WW.m:56-164): nit=200, population {args.pop} per GPU, "
                "p_i~U{1..30}, m_j~U{1..min(p,10)}, delta, lambda log-uniform [1e-3,1e2], robustness filter (alfa, raio) per candidate designed on the device, PCG64 seed 0")
    if args.config == "ssnmpc":
        return (f"single-shooting NMPC sweep (Explicit NMPC/ClosedLoopNMPC.m:1-110, SURVEY 8f rank 4): nit=150 (147 controller calls per "
                f"candidate), population {args.pop} per GPU, N~U{{2..12}}, Nu_j~U{{1..min(4,N)}} per input, Q log-uniform [0.1,10], "
                "W log-uniform [1e-5,1e-2], tracking cost, PCG64 seed 0")
    return (f"Van de Vusse NMPC closed-loop evaluation (BASELINE.json configs[4], closedloop_toolbox_nmpc.m:36-97): nit=60, "
            f"population {args.pop} per GPU, N~U{{3..31}}, Nu~U{{2..15}}, delta, lambda log-uniform [1e-3,10], GAM cost, PCG64 seed 0")


def other_flops(args, pop):
    """SURVEY 8(d)-style algorithmic flops of one population (fp64)."""
    if args.config == "dtc":
        p, m = pop[0].astype(float), pop[1].astype(float)
        P, M = p.sum(axis=1), m.sum(axis=1)
        nit = 200
        return float(np.sum(2 * P * M * M + M ** 3 / 3 + 2 * P * M + nit * (2 * 2 * (P + 40) + 60)))
    if args.config == "ssnmpc":
        N, nz = pop[0].astype(float), pop[1].astype(float).sum(axis=1)
        per_gn = N * 16 * 420 + N * 2 * nz * nz + nz ** 3 / 3 + 2 * N * 16 * 120    # the same convention as the other NMPC line: 3 iterations per call
        return float(np.sum(147 * 3 * per_gn))
    N, Nu = pop[0].astype(float), pop[1].astype(float)
    nz = 2 * Nu
    per_sqp = N * 16 * 420 + N * 2 * nz * nz + nz ** 3 / 3 + 2 * N * 16 * 120     # rollout+sensitivities, H, factor, 2 cost rollouts
    return float(np.sum(59 * 3 * per_sqp))


def run_other(args):
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    import mpcgpu
    ref = args.impl == "reference"
    if ref and rank != 0:
        return
    n_all = args.pop * (1 if ref else world)
    if args.config == "dtc":
        prob = mpcgpu.woodberry_dtc()
        pop = mpcgpu.synthetic_dtc_population(prob, n_all, seed=0)   # (p, m, delta, lambda, alfa, raio): the robustness filter of every candidate is designed on the device
    elif args.config == "ssnmpc":
        prob = mpcgpu.explicit_nmpc()
        pop = mpcgpu.synthetic_ssnmpc_population(prob, n_all, seed=0)
    else:
        prob = mpcgpu.vandevusse()
        pop = mpcgpu.synthetic_nmpc_population(prob, n_all, seed=0)
    sl = slice(0, args.pop) if ref else slice(rank, None, world)
    mine = [a[sl] for a in pop]
    n = len(mine[0])
    cores = host_threads()

    def cpu_rate(nth, budget_s):
        """oracle leg: DTC-GPC = oracle/dtc_gpc_oracle.py (numpy restatement of the reference's own MATLAB, one process);
        Van de Vusse = oracle/nmpc_port.cpp (the kernel's algorithm compiled for the host, OpenMP over candidates)."""
        if args.config == "dtc":
            from oracle import dtc_gpc_oracle as dorc
            t0 = time.perf_counter(); k = 0
            while time.perf_counter() - t0 < budget_s and k < n:
                fr = dorc.mimofilter_Fr(prob.pnz, float(mine[4][k]), float(mine[5][k]))
                dorc.dtc_gpc_closed_loop(prob, mine[0][k], mine[1][k], mine[2][k], mine[3][k], fr)
                k += 1
            return k / (time.perf_counter() - t0), 1, k
        from oracle import nmpc_port
        if args.config == "ssnmpc":
            k = min(n, max(2 * nth, 16))
            t0 = time.perf_counter(); nmpc_port.ssnmpc_eval_batch(prob, *[a[:k] for a in mine], nthreads=nth); dt = time.perf_counter() - t0
            k = int(min(n, max(k, k * budget_s / max(dt, 1e-3))))
            t0 = time.perf_counter(); nmpc_port.ssnmpc_eval_batch(prob, *[a[:k] for a in mine], nthreads=nth); dt = time.perf_counter() - t0
            return k / dt, nth, k
        k = min(n, max(2 * nth, 16))
        t0 = time.perf_counter(); nmpc_port.eval_batch(prob, *[a[:k] for a in mine], "gam", nth); dt = time.perf_counter() - t0
        k = int(min(n, max(k, k * budget_s / max(dt, 1e-3))))
        t0 = time.perf_counter(); nmpc_port.eval_batch(prob, *[a[:k] for a in mine], "gam", nth); dt = time.perf_counter() - t0
        return k / dt, nth, k

    if ref:
        for _ in range(min(args.warmup, 1)):
            cpu_rate(cores, 1.0)
        t0 = time.perf_counter(); tot = 0
        for _ in range(args.steps):
            v_, c_, k_ = cpu_rate(cores, 2.0)
            tot += k_
        dt = time.perf_counter() - t0
        val = tot / dt
        sample = f"~2 s of CPU work per step on rank 0's population ({tot // max(args.steps, 1)} candidates per step)"
        kind_note = ("oracle/dtc_gpc_oracle.py: numpy restatement of DTC_GPC_WW.m, single process" if args.config == "dtc"
                     else "oracle/nmpc_port.cpp: the restated NLP + Gauss-Newton SQP on the host, OpenMP over candidates")
        if args.config == "ssnmpc":
            kind_note = "oracle/nmpc_port.cpp: csrc/mpc_ssnmpc_core.h (the kernel's per-run source) on the host, OpenMP over candidates"
        print(json.dumps({"impl": "reference", "metric": metric_name(args), "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": other_config(args, args.pop), "reference_note": kind_note + "; " + sample,
                          "cpu_baseline": {"value": val, "unit": UNIT, "cores": c_, "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: mpcgpu has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.config == "dtc":
        ev = mpcgpu.DtcEvaluator(prob, device=local)
        call = lambda: ev.eval_batch(*mine[:4], alfa=mine[4], raio=mine[5])
        width, key = 2, "ise"
        h2d = n * (2 * 4 + 2 * 4 + 2 * 8 + 2 * 8 + 2 * 8)
    elif args.config == "ssnmpc":
        ev = mpcgpu.SsnmpcEvaluator(prob, device=local)
        call = lambda: ev.eval_batch(*mine)
        width, key = 2, "cost"
        h2d = n * (4 + 8 + 16 + 16)
    else:
        ev = mpcgpu.NmpcEvaluator(prob, device=local)
        call = lambda: ev.eval_batch(*mine, mode="gam")
        width, key = 2, "cost"
        h2d = n * (4 + 4 + 16 + 16)
    gathered = torch.empty(n * width * world, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    for _ in range(max(args.warmup, 3)):
        out = call()
    fp64_peak = mpcgpu.measure_fp64_peak(local)
    c0 = ev.counters()
    kern, wall = [], []
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local) as clk:
        for k in range(args.steps):
            flush.fill_(float(k)); torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = call()                      # host arrays -> C ABI -> host costs (H2D, kernel, D2H inside)
            if world > 1:
                dist.all_gather_into_tensor(gathered, torch.from_numpy(out[key].reshape(-1)).cuda())
                torch.cuda.synchronize()
            wall.append(time.perf_counter() - t0)
            kern.append(ev.counters()["last_sim_ms"])
    c1 = ev.counters()
    t = torch.tensor([sum(kern), sum(wall) * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    kern_ms, wall_ms = [float(x) for x in t.tolist()]
    if rank == 0:
        fl = other_flops(args, mine)
        achieved = fl / (kern_ms / args.steps) / 1e9
        ok = np.isin(out["status"], (0, 5)) if args.config == "vdv" else out["status"] == 0
        line = {"metric": metric_name(args), "value": n * world * args.steps / (kern_ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": kern_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": other_config(args, n),
                "e2e": {"value": n * world * args.steps / (wall_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(n * width * 8 + n * 4), "timing": "wall clock around the synchronous C-ABI call"},
                "gpu_launches": int(c1["kernel_launches"] - c0["kernel_launches"]),
                "clocks": clk.summary(),
                "roofline": {"bound": "fp64_fma (serial per-run chains; neither hbm nor tensor)", "achieved": achieved, "peak": fp64_peak,
                             "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak else None,
                             "kernel": {"dtc": "k_dtc (warp per candidate)", "ssnmpc": "k_ssnmpc (one thread per closed loop, sorted by horizons)"}.get(
                                 args.config, "k_nmpc_g<16> (sixteen lanes per run, two runs per warp, sorted by horizons)"),
                             "algorithmic_flops_per_launch": fl,
                             "traffic": None if args.config in ("dtc", "ssnmpc") else ncu_traffic("r2_other_kernels.json", "k_nmpc_g", (n + 1) // 2),
                             "traffic_source": "profiles/r2_other_kernels.json (ncu --set full, same population; null otherwise)",
                             "peak_source": "mpcgpu_measure_fp64_peak, measured live"},
                "failed_candidates": int((~ok).sum())}
        if not args.no_cpu_baseline and world == 1:
            v_, c_, k_ = cpu_rate(cores, 10.0)
            line["cpu_baseline"] = {"value": v_, "unit": UNIT, "cores": c_, "kind": "port", "sample": f"first {k_} candidates of the same population"}
            if args.config in ("vdv", "ssnmpc"):
                v1, _, k1 = cpu_rate(1, 4.0)
                line["cpu_baseline"]["single_thread"] = {"value": v1, "cores": 1, "sample": f"first {k1} candidates"}
        print(json.dumps(line), flush=True)
    ev.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pop", type=int, default=0, help="candidates per GPU (default: 4096; shell7x5 2048; dtc 16384; vdv 16384)")
    ap.add_argument("--mode", default="gam", choices=["gam", "vns"])
    ap.add_argument("--no-other-configs", action="store_true", help="skip the Shell7x5 / DTC-GPC / NMPC context numbers")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of every timed candidate")
    ap.add_argument("--parity-n", type=int, default=512, help="candidates checked against the (slow) oracle for the configurations other than the headline")
    ap.add_argument("--config", default="shell3x3", choices=["shell3x3", "shell7x5", "dtc", "vdv", "ssnmpc"],
                    help="headline (BASELINE.json configs[1], default) or one of the other configurations")
    ap.add_argument("--fixed", default="", help="p,m : pin every candidate's horizons (diagnostic populations)")
    ap.add_argument("--weights", default="", help="lo,hi : log-uniform weight range (default 1e-4,10)")
    ap.add_argument("--lam", default="", help="lo,hi : override the lambda range only (diagnostics)")
    args = ap.parse_args()
    if args.pop <= 0:
        args.pop = {"shell3x3": 4096, "shell7x5": 2048, "dtc": 16384, "vdv": 16384, "ssnmpc": 4096}[args.config]
    if args.config in ("dtc", "vdv", "ssnmpc"):
        return run_other(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import mpcgpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: mpcgpu has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    prob = linear_problem(args)
    nit = prob.nit
    # weak scaling: every rank evaluates its own `pop` candidates of one global seeded population
    fixed = tuple(int(x) for x in args.fixed.split(",")) if args.fixed else None
    wlo, whi = (float(x) for x in args.weights.split(",")) if args.weights else (1e-4, 10.0)
    Ng, Nug, dg, lg = mpcgpu.synthetic_population(prob, args.pop * world, seed=0, fixed=fixed, wlo=wlo, whi=whi)
    if args.lam:
        llo, lhi = (float(x) for x in args.lam.split(","))
        lg = np.exp(np.random.default_rng(5).uniform(np.log(llo), np.log(lhi), size=lg.shape))
    sl = shard_of(prob, Ng, Nug, dg, lg, world, rank)
    N, Nu, delta, lam = Ng[sl], Nug[sl], dg[sl], lg[sl]
    n = len(N)
    ev = mpcgpu.Evaluator(prob, device=local)
    ncost = n * prob.ny if args.mode == "gam" else n
    gathered = torch.empty(ncost * world, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # 256 MiB > 126 MB L2
    tstream = torch.cuda.Stream()          # a real (non-NULL) stream: the events below must see the kernels
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    class _DevView:
        def __init__(self, ptr, count):
            self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3}

    def step_resident(ev_mid=None):
        ev.run(args.mode, stream=stream)
        if ev_mid is not None:
            ev_mid.record()          # end of this rank's kernels, start of the fitness all-gather
        if world > 1:
            ptr, cnt = ev.cost_device_ptr(args.mode)
            local_cost = torch.as_tensor(_DevView(ptr, cnt), device="cuda")
            dist.all_gather_into_tensor(gathered, local_cost)

    ev.upload(N, Nu, delta, lam)
    for _ in range(max(args.warmup, 3)):
        step_resident()
    torch.cuda.synchronize()
    fp64_peak = mpcgpu.measure_fp64_peak(local)

    c0 = ev.counters()
    ev_a = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev_b = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev_m = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sim_ms, build_ms = [], []
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local) as clk:
        for k in range(args.steps):
            flush.fill_(float(k))            # evict L2 between timed iterations (untimed)
            ev_a[k].record()
            step_resident(ev_m[k])
            ev_b[k].record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        step_ms = [a.elapsed_time(b) for a, b in zip(ev_a, ev_b)]
        kern_ms = float(np.mean([a.elapsed_time(m_) for a, m_ in zip(ev_a, ev_m)]))      # this rank's kernels
        gath_ms = float(np.mean([m_.elapsed_time(b) for m_, b in zip(ev_m, ev_b)]))      # wait for the slowest rank + all-gather
        c_mid = ev.counters()
        # e2e: host arrays -> C ABI -> host costs, every step (synchronous call)
        for _ in range(2):
            ev.eval_batch(N, Nu, delta, lam, mode=args.mode)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = ev.eval_batch(N, Nu, delta, lam, mode=args.mode)
            if world > 1:
                dist.all_gather_into_tensor(gathered, torch.from_numpy(out["cost"].reshape(-1)).cuda())
                torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
    c1 = ev.counters()
    res = ev.download(args.mode)
    nfail = int((res["status"] != 0).sum())
    # ---- parity of what was timed: EVERY candidate of this rank's shard against the CPU oracle (checker, not product) ----
    par = None
    cpu_rate_all = None
    if not args.no_parity:
        from oracle import oracle as orc
        from oracle import parity as opar
        op = orc.OracleProblem(prob)
        nth = max(1, host_threads() // max(1, world if os.environ.get("LOCAL_RANK") is not None else 1))
        npar = len(N) if args.config == "shell3x3" else min(len(N), args.parity_n)   # the soft-constraint oracle does ~10 candidates/s/thread
        t0 = time.perf_counter()
        g0, st0, _ = orc.eval_batch(op, N[:npar], Nu[:npar], delta[:npar], lam[:npar], args.mode, nth)
        cpu_rate_all = (npar / (time.perf_counter() - t0), nth)
        sens = opar.sensitivity(op, N[:npar], Nu[:npar], delta[:npar], lam[:npar], args.mode, g0, nth)
        par = opar.summary(res["cost"][:npar], res["status"][:npar], g0, st0, sens)
        par["checked"] = f"first {npar} of the {len(N)} timed candidates of this rank" if npar < len(N) else "every timed candidate"
        e2e_cost = out["cost"]
        par["e2e_call_bit_identical_to_resident_run"] = bool(np.array_equal(e2e_cost, res["cost"], equal_nan=True))
        if world > 1:   # every rank checked its own shard: add up
            keys = ["n", "n_compared", "n_gt_1e-6", "n_sensitivity_relaxed", "n_out_of_tolerance", "n_status_nonzero", "n_status_nonzero_oracle"]
            tsum = torch.tensor([par[k] for k in keys], dtype=torch.float64, device="cuda")
            tmax = torch.tensor([par["max_rel"] or 0.0, par["max_rel_well_posed"] or 0.0], dtype=torch.float64, device="cuda")
            dist.all_reduce(tsum); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            for k, vv in zip(keys, tsum.tolist()):
                par[k] = int(vv)
            par["max_rel"], par["max_rel_well_posed"] = [float(x) for x in tmax.tolist()]
            par["frac_le_1e-6"] = 1.0 - par["n_gt_1e-6"] / max(par["n_compared"], 1)
            par["median_rel"] = None
            # the gathered fitness equals what each rank computed (rank 0 checks its own slab of the all-gather)
            loc = torch.from_numpy(res["cost"].reshape(-1)).cuda()
            par["allgather_slab_bit_identical"] = bool(torch.equal(gathered[rank * ncost:(rank + 1) * ncost], loc))
    # per-phase device times of the last resident run
    ev.run(args.mode, stream=stream)
    torch.cuda.synchronize()
    ev.download(args.mode)
    cn = ev.counters()
    build_ms, sim_ms = cn["last_build_ms"], cn["last_sim_ms"]

    total_ms = float(sum(step_ms))
    t = torch.tensor([total_ms, e2e_s * 1e3, sim_ms, build_ms], dtype=torch.float64, device="cuda")
    per_rank = torch.zeros(2 * world, dtype=torch.float64, device="cuda")
    per_rank[2 * rank] = kern_ms; per_rank[2 * rank + 1] = gath_ms
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(per_rank)
    total_ms, e2e_ms, sim_ms, build_ms = [float(x) for x in t.tolist()]
    per_rank = per_rank.tolist()
    value = n * world * args.steps / (total_ms * 1e-3)
    e2e_val = n * world * args.steps / (e2e_ms * 1e-3)

    if rank == 0:
        flops, nst = algorithmic_flops(prob, N, Nu, nit)
        f_survey = survey_flops(prob, N, Nu, nit)
        runs = prob.ny if (args.mode == "vns" and prob.square) else 1
        # roofline.achieved uses SURVEY.md 8(d)'s ALGORITHMIC figure F_cand (dense formulation of the reference's math);
        # the flops this implementation actually executes (prefix-Gram tables remove most of them) are reported next to it
        achieved = f_survey * runs / (sim_ms + build_ms) / 1e9   # TFLOP/s: flops / (ms * 1e-3) / 1e12
        executed = flops * runs / (sim_ms + build_ms) / 1e9
        hbm_alg = float(np.sum((nst * prob.nu * Nu + (prob.nu * Nu) ** 2) * 8.0 * 2)) + n * (8 + 8 * (prob.ny + prob.nu) + 8 * prob.ny)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        launches = c_mid["kernel_launches"] - c0["kernel_launches"]   # kernels of this library inside the timed region
        line = {
            "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": linear_config(args, world, n, nit),
            "qp_solves_per_s": value * runs * (nit + (1 if args.mode == "vns" else 0)),
            "per_rank_ms": {"kernel_ms": [round(per_rank[2 * r_], 4) for r_ in range(world)],
                            "allgather_ms_incl_wait_for_slowest_rank": [round(per_rank[2 * r_ + 1], 4) for r_ in range(world)]},
            "parity": par,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(n * (8 + 8 * (prob.ny + prob.nu) + 16 + 8)),
                    "d2h_bytes_per_step": int(ncost * 8 + n * 4 + 16), "timing": "wall clock around the synchronous C-ABI call"},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": {"bound": "fp64_fma (latency-bound serial QP chain; neither hbm nor tensor, SURVEY.md 8d)",
                         "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak else None,
                         "peak_source": "mpcgpu_measure_fp64_peak, measured live (MEASURED_PEAKS.json has no fp64 entry)",
                         "kernel": "k_build + " + ("k_soft<3,16> (block per run, soft output limits)" if (np.isfinite(prob.ymin).any() or np.isfinite(prob.ymax).any()) else "k_sim<3,16,lean,spec[,M in shared memory for small populations]> (warp per run, speculative)") + ", one launch per population",
                         "kernel_ms": {"k_sim": sim_ms, "k_build": build_ms},
                         "algorithmic_flops_per_launch": f_survey * runs,
                         "executed_flops_per_launch": flops * runs, "executed_tflops": executed,
                         "executed_frac": executed / fp64_peak if fp64_peak else None,
                         "hbm": {"algorithmic_bytes": hbm_alg, "achieved_gbs": hbm_alg / (sim_ms + build_ms) / 1e6,
                                 "peak_gbs": peaks.get("hbm_gbs"), "frac": (hbm_alg / (sim_ms + build_ms) / 1e6) / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None},
                         # dram__bytes_read.sum + dram__bytes_write.sum of the closed-loop kernel of one population: not
                         # measurable inside bench.py (needs ncu); cited from the committed capture of the same command
                         "traffic": (ncu_traffic("r2_other_kernels.json", "k_soft", n * runs) if (np.isfinite(prob.ymin).any() or np.isfinite(prob.ymax).any())
                                     else ncu_traffic("r2_k_sim_summary.json", "k_sim", n * runs)) if args.mode == "gam" else None,
                         "traffic_source": "profiles/r2_k_sim_summary.json / r2_other_kernels.json (ncu --set full of the same kernel on the same population; null if the population differs)"},
            "counters": {k: cn[k] for k in ("qp_constrained", "as_iterations", "qp_solves", "closed_loops")},
            "failed_candidates": nfail,
        }
        if not args.no_cpu_baseline and world == 1:
            v, cores, ns = cpu_reference_rate(prob, N, Nu, delta, lam, args.mode)
            v1, _, ns1 = cpu_reference_rate(prob, N, Nu, delta, lam, args.mode, budget_s=4.0, nthreads=1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {ns} candidates of the same seeded population, oracle/mpc_oracle.c, OpenMP over candidates",
                                    "single_thread": {"value": v1, "cores": 1, "sample": f"first {ns1} candidates"},
                                    "per_core": v / cores}
        if world == 1 and not args.no_other_configs and args.config == "shell3x3":
            line["other_configs"] = other_configs()
        print(json.dumps(line), flush=True)
    ev.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
