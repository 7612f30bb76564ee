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

METRIC = "closed-loop tuning candidates/sec (Shell3x3)"
UNIT = "candidates/s"


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
        P7 = mpcgpu.synthetic_population(p7, 2048, seed=0, wlo=1e-2)
        e7.eval_batch(*[a[:64] for a in P7], mode="gam")
        t0 = time.perf_counter(); o7 = e7.eval_batch(*P7, mode="gam"); dt = time.perf_counter() - t0
        out["shell7x5_soft_constraints"] = {"candidates": 2048, "candidates_per_s": 2048 / dt, "failed": int((o7["status"] != 0).sum()),
                                            "weights": "lambda log-uniform [1e-2, 10], delta = 0 (band control)"}
        e7.close()
        pd = mpcgpu.woodberry_dtc(); ed = mpcgpu.DtcEvaluator(pd, device=0)
        Pd = mpcgpu.synthetic_dtc_population(pd, 16384, seed=0)
        fil = [mpcgpu.mimo_filter(pd.pnz, float(a), float(r)) for a, r in zip(Pd[4][:256], Pd[5][:256])] * 64
        ed.eval_batch(*[a[:64] for a in Pd[:4]], filters=fil[:64])
        t0 = time.perf_counter(); od = ed.eval_batch(*Pd[:4], filters=fil); dt = time.perf_counter() - t0
        out["dtc_gpc_sweep"] = {"candidates": 16384, "candidates_per_s": 16384 / dt, "failed": int((od["status"] != 0).sum())}
        ed.close()
        pn = mpcgpu.vandevusse(); en = mpcgpu.NmpcEvaluator(pn, device=0)
        Pn = mpcgpu.synthetic_nmpc_population(pn, 4096, seed=0)
        en.eval_batch(*[a[:64] for a in Pn], mode="gam")
        t0 = time.perf_counter(); on = en.eval_batch(*Pn, mode="gam"); dt = time.perf_counter() - t0
        out["vandevusse_nmpc"] = {"candidates": 4096, "candidates_per_s": 4096 / dt, "failed": int((~np.isin(on["status"], (0, 5))).sum())}
        en.close()
    except Exception as exc:   # context only: never fail the headline line
        out["error"] = repr(exc)
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


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  MATLAB + the closed-source MPC
    Toolbox cannot run here, so this is the oracle port on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import mpcgpu
    prob = mpcgpu.shell3x3(2)
    N, Nu, delta, lam = mpcgpu.synthetic_population(prob, args.pop, seed=0)
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
    sample = f"first {n} of the {args.pop}-candidate seeded Shell3x3 population per step"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"Shell3x3 {args.mode.upper()} closed-loop evaluation, nit=500, CPU oracle port", "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "qp_solves_per_s": val * 501 if args.mode == "gam" else None}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pop", type=int, default=4096, help="candidates per GPU")
    ap.add_argument("--mode", default="gam", choices=["gam", "vns"])
    ap.add_argument("--no-other-configs", action="store_true", help="skip the Shell7x5 / DTC-GPC / NMPC context numbers")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fixed", default="", help="p,m : pin every candidate's horizons (diagnostic populations)")
    ap.add_argument("--weights", default="", help="lo,hi : log-uniform weight range (default 1e-4,10)")
    ap.add_argument("--lam", default="", help="lo,hi : override the lambda range only (diagnostics)")
    args = ap.parse_args()
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

    prob = mpcgpu.shell3x3(2)
    nit = prob.nit
    # weak scaling: every rank evaluates its own `pop` candidates of one global seeded population
    fixed = tuple(int(x) for x in args.fixed.split(",")) if args.fixed else None
    wlo, whi = (float(x) for x in args.weights.split(",")) if args.weights else (1e-4, 10.0)
    Ng, Nug, dg, lg = mpcgpu.synthetic_population(prob, args.pop * world, seed=0, fixed=fixed, wlo=wlo, whi=whi)
    if args.lam:
        llo, lhi = (float(x) for x in args.lam.split(","))
        lg = np.exp(np.random.default_rng(5).uniform(np.log(llo), np.log(lhi), size=lg.shape))
    sl = slice(rank, None, world)   # round-robin shard (sizes are i.i.d., so this is work-balanced)
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

    def step_resident():
        ev.run(args.mode, stream=stream)
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
    sim_ms, build_ms = [], []
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local) as clk:
        for k in range(args.steps):
            flush.fill_(float(k))            # evict L2 between timed iterations (untimed)
            ev_a[k].record()
            step_resident()
            ev_b[k].record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        step_ms = [a.elapsed_time(b) for a, b in zip(ev_a, ev_b)]
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
    # per-phase device times of the last resident run
    ev.run(args.mode, stream=stream)
    torch.cuda.synchronize()
    ev.download(args.mode)
    cn = ev.counters()
    build_ms, sim_ms = cn["last_build_ms"], cn["last_sim_ms"]

    total_ms = float(sum(step_ms))
    t = torch.tensor([total_ms, e2e_s * 1e3, sim_ms, build_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, sim_ms, build_ms = [float(x) for x in t.tolist()]
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
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"Shell3x3 {args.mode.upper()} closed-loop evaluation (BASELINE.json configs[1]): nit=500, "
                                   f"population {n} per GPU, N~U{{7..127}}, Nu~U{{2..15}}, weights log-uniform [1e-4,10], PCG64 seed 0",
                       "population_per_gpu": n, "nit": nit, "cost_mode": args.mode,
                       "l2": "flushed between timed steps (256 MiB write, untimed)",
                       "multi_gpu": "population sharded round-robin, one NCCL all-gather of fitness per step" if world > 1 else "single GPU"},
            "qp_solves_per_s": value * runs * (nit + (1 if args.mode == "vns" else 0)),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(n * (8 + 8 * (prob.ny + prob.nu) + 16 + 8)),
                    "d2h_bytes_per_step": int(ncost * 8 + n * 4 + 16), "timing": "wall clock around the synchronous C-ABI call"},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": {"bound": "fp64_fma (latency-bound serial QP chain; neither hbm nor tensor, SURVEY.md 8d)",
                         "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak else None,
                         "peak_source": "mpcgpu_measure_fp64_peak, measured live (MEASURED_PEAKS.json has no fp64 entry)",
                         "kernel": "k_build + k_sim<3,16> (one launch per population)",
                         "kernel_ms": {"k_sim": sim_ms, "k_build": build_ms},
                         "algorithmic_flops_per_launch": f_survey * runs,
                         "executed_flops_per_launch": flops * runs, "executed_tflops": executed,
                         "executed_frac": executed / fp64_peak if fp64_peak else None,
                         "hbm": {"algorithmic_bytes": hbm_alg, "achieved_gbs": hbm_alg / (sim_ms + build_ms) / 1e6,
                                 "peak_gbs": peaks.get("hbm_gbs"), "frac": (hbm_alg / (sim_ms + build_ms) / 1e6) / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None},
                         # dram__bytes_read.sum + dram__bytes_write.sum of the k_sim launch of one population,
                         # ncu --set full capture profiles/r1k_k_sim_summary.json (same population, GAM mode)
                         "traffic": 65.2e6 if (args.mode == "gam" and n == 4096) else None},
            "counters": {k: cn[k] for k in ("qp_constrained", "as_iterations", "qp_solves", "closed_loops")},
            "failed_candidates": nfail,
        }
        if not args.no_cpu_baseline and world == 1:
            v, cores, ns = cpu_reference_rate(prob, N, Nu, delta, lam, args.mode)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {ns} candidates of the same seeded population, oracle/mpc_oracle.c, OpenMP over candidates"}
        if world == 1 and not args.no_other_configs:
            line["other_configs"] = other_configs()
        print(json.dumps(line), flush=True)
    ev.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
