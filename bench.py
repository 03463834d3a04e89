#!/usr/bin/env python
"""Benchmark of the GPcounts per-gene GP fitting hot path on B200 (contract: see the task prompt / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--genes G] [--impl reference]

Workload (BASELINE.json configs[4], the configuration the metric is quoted on): synthetic C5 generator
 - N = 1000 spots, sparse NB with M = 64 inducing points, One_sample_test (dynamic RBF fit +
constant fit + LLR per gene).  A *step* fits one fresh batch of G genes per GPU (default 8192) drawn from that
generator; value = genes fitted per second over all GPUs with the batch resident in HBM, e2e = the same through
``Fit_GPcounts(...).One_sample_test`` from host DataFrames (H2D / D2H inside the timed region).
Under torchrun every rank fits its own batches (weak scaling, no data-path collective); time = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "genes_fitted_per_sec"
UNIT = "genes/s"


def synth_batch(G, N, seed):
    """Vectorised C5 generator: even genes dynamic, odd genes constant."""
    X = np.sort(np.random.default_rng(0).uniform(0, 1, N))[:, None]
    rng = np.random.default_rng(seed)
    b = rng.uniform(np.log(0.5), np.log(50), G)
    a = np.exp(rng.uniform(np.log(0.02), np.log(1.5), G))
    a[1::2] = 0.0
    om = rng.uniform(0.5, 2, G)
    ph = rng.uniform(0, 2 * np.pi, G)
    alpha = np.exp(rng.uniform(np.log(0.05), np.log(2), G))
    f = b[:, None] + a[:, None] * np.sin(2 * np.pi * om[:, None] * X[None, :, 0] + ph[:, None])
    r = (1.0 / alpha)[:, None]
    Y = rng.negative_binomial(r, r / (r + np.exp(f))).astype(np.float64)
    return X, Y


def svgp_flops(M, N):
    """Algorithmic FLOPs of one SVGP ELBO+gradient evaluation."""
    return 6.0 * M * M * N + (4.0 / 3.0) * M ** 3 + 12.0 * M * N


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_one_sample(X, Y, M, n_jobs):
    """Time the CPU oracle (restatement of the reference's GPflow + SciPy path) on the genes of Y."""
    from joblib import Parallel, delayed

    def one(g):
        import torch
        torch.set_num_threads(1)
        from oracle.fit import OracleFit
        o = OracleFit(X, Y, sparse=True, M=M, inducing_variance=1.0)
        r = o.one_sample_test("Negative_binomial", genes=[g])[0]
        return sum(f.nfev for f in r["fits"] if f is not None)

    t0 = time.time()
    nfev = Parallel(n_jobs=n_jobs)(delayed(one)(g) for g in range(Y.shape[0]))
    dt = time.time() - t0
    return Y.shape[0] / dt, float(np.sum(nfev)) / dt, dt


def dist_info():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  GPflow / TensorFlow / RobustGP are not
    installable here (no wheels, no network - DESIGN.md), so this is the oracle port: torch-FP64 restatement of the
    GPflow ELBO + the installed SciPy L-BFGS-B, one process per host core."""
    rank, world, _ = dist_info()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = int(args.cpu_genes or max(16, min(4 * cores, 96)))
    times, evals = [], []
    for s in range(args.warmup + args.steps):
        X, Y = synth_batch(sample, args.N, seed=1234 + s)
        gps, eps, dt = cpu_one_sample(X, Y, args.M, cores)
        if s >= args.warmup:
            times.append(dt); evals.append(eps)
    T = float(np.sum(times))
    value = sample * len(times) / T
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, sample),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d genes per step (One_sample_test, N=%d, M=%d), oracle port: torch-FP64 ELBO + SciPy "
                                   "L-BFGS-B, one process per core" % (sample, args.N, args.M),
                         "elbo_grad_evals_per_sec": float(np.mean(evals))},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, genes_per_step):
    return {"workload": "C5 synthetic sparse NB One_sample_test (SURVEY 8d generator): N=%d spots, M=%d inducing, "
                        "20k-gene generator, step = fresh batch of genes" % (args.N, args.M),
            "genes_per_step_per_gpu": genes_per_step, "N": args.N, "M": args.M, "models_per_gene": 2,
            "optimizer": "L-BFGS-B m=10 ftol=2.22e-9 gtol=1e-5 maxiter=5000 (SciPy defaults of the reference)",
            "l2_policy": "fresh gene batch every step (64 MB of counts) and 148 per-slot workspaces of ~3.2 MB (~470 MB) "
                         "that every evaluation rewrites: the working set exceeds the 126 MB L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--genes", type=int, default=8192, help="genes per step per GPU")
    ap.add_argument("--N", type=int, default=1000)
    ap.add_argument("--M", type=int, default=64)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-genes", type=int, default=0, help="genes in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--maxiter", type=int, default=5000, help="profiling only: cap L-BFGS iterations")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank, world, local = dist_info()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import pandas as pd
    from gpcounts_b200 import Fit_GPcounts, engine, host
    from gpcounts_b200._lib import S_NFEV_TOTAL, S_STATUS

    G, N, M = args.genes, args.N, args.M
    nsteps = args.warmup + args.steps
    # ---- device-resident leg: models / inducing points built once, batches already in HBM -------------------------
    X, _ = synth_batch(1, N, seed=0)
    ls0 = host.default_lengthscale(X)
    Zsets = host.inducing_sets(X, M, ls0)
    rtab, _ = host.restart_table(X)
    models = [engine.Model("SVGP", "RBF", "NB", X, Z=Zsets, restart=rtab, device=dev),
              engine.Model("SVGP", "CONST", "NB", X, Z=Zsets, restart=rtab, handoff_alpha=True, device=dev)]
    opt = engine.default_opt(maxiter=args.maxiter)
    batches = []
    for s in range(nsteps):
        _, Y = synth_batch(G, N, seed=1234 + 1000 * rank + s)
        h = np.zeros((G, 2, 4))
        h[:, :, 0] = host.default_variance(Y, "Negative_binomial", True)[:, None]
        h[:, :, 1], h[:, :, 2], h[:, :, 3] = ls0, 1.0, 35.0
        batches.append((Y, torch.as_tensor(Y, device=dev), torch.as_tensor(h, device=dev)))
    torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    launches0 = engine.launch_count()
    sampler = ClockSampler(local)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
    stats_all = []
    for s in range(nsteps):
        if s == args.warmup:
            barrier()
            sampler.start()
            launches0 = engine.launch_count()
            t_wall0 = time.time()
        ev[s][0].record()
        stats, _ = engine.fit(models, batches[s][1], None, batches[s][2], chain=True, opt=opt, return_theta=False)
        ev[s][1].record()
        stats_all.append(stats)
    barrier()
    t_wall = time.time() - t_wall0
    clocks = sampler.stop()
    launches = engine.launch_count() - launches0
    kern_ms = [ev[s][0].elapsed_time(ev[s][1]) for s in range(args.warmup, nsteps)]
    T = float(np.sum(kern_ms)) * 1e-3
    if world > 1:
        t = torch.tensor([T], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        T = float(t.item())
    value = G * args.steps * world / T

    st = torch.stack(stats_all[args.warmup:]).cpu().numpy()          # (K, G, 2, 16)
    nfev = st[..., S_NFEV_TOTAL].sum()
    flops = nfev * svgp_flops(M, N)
    ok_frac = float((st[..., S_STATUS] == 0).mean())
    eval_cyc, fit_cyc = st[..., 12].sum(), st[..., 13].sum()

    # ---- end-to-end leg: the public API from host DataFrames --------------------------------------------------------
    cells = ["c%d" % i for i in range(N)]
    Xdf = pd.DataFrame(X, index=cells, columns=["pseudotime"])
    e2e_T, h2d, d2h = 0.0, 0, 0
    e2e_steps = min(args.steps, 2)
    barrier()
    for s in range(e2e_steps):
        Y = batches[args.warmup + s][0]
        Ydf = pd.DataFrame(Y, index=["g%d" % i for i in range(G)], columns=cells)
        barrier()
        t0 = time.time()
        gp = Fit_GPcounts(Xdf, Ydf, sparse=True, M=M, device=dev)
        gp.distributed = False          # weak scaling like `value`: every rank pushes its own batch through the API
        gp.optimizer_options = {"maxiter": args.maxiter}
        df = gp.One_sample_test("Negative_binomial")
        llr_sum = float(np.nansum(df["log_likelihood_ratio"].values))     # result read on the host
        torch.cuda.synchronize(dev)
        e2e_T += time.time() - t0
        h2d += Y.nbytes + G * 2 * 4 * 8 + X.nbytes + Zsets.nbytes + 2 * rtab.nbytes
        d2h += G * 2 * 16 * 8 + (4 + M + M * (M + 1) // 2) * 8      # statistics rows + fitted parameters of the last gene
    if world > 1:
        t = torch.tensor([e2e_T], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_T = float(t.item())
    e2e_value = G * e2e_steps * world / e2e_T

    if rank == 0:
        # FP64 roofline denominator: cuBLAS DGEMM measured on this box right now (MEASURED_PEAKS.json has no FP64 entry)
        a = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
        for _ in range(2):
            a @ a
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); a @ a; e1.record(); torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        dgemm_tf = 2 * 4096 ** 3 / best * 1e-9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):      # measured per gene on a 148-gene ncu capture, scaled to the genes of one launch
            traffic = json.load(open(tp)).get("dram_bytes_per_gene", 0.0) * G or None
        achieved = flops / (float(np.sum(kern_ms)) * 1e-3) * 1e-12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, G),
            "gene_fits_per_hour_per_gpu": value * 3600 / world,
            "elbo_grad_evals_per_sec": float(nfev) * world / T,
            "evals_per_gene": float(nfev) / (G * args.steps), "fit_ok_fraction": ok_frac,
            "time_split": {"eval_share_of_fit_cycles": float(eval_cyc / fit_cyc),
                           "us_per_eval_in_fit": float(eval_cyc / nfev / (clocks.get("sm_mhz") or 1965.0)),
                           "slot_busy_fraction": float(fit_cyc / ((clocks.get("sm_mhz") or 1965.0) * 1e3 * float(np.sum(kern_ms)) * min(engine.default_slots(models), G)))},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": dgemm_tf, "unit": "TFLOP/s",
                         "frac": achieved / dgemm_tf, "traffic": traffic,
                         "pipe": "FP64 tensor pipe (DMMA.8x8x4; tcgen05 has no FP64 kind)",
                         "peak_source": "FP64: cuBLAS DGEMM 4096^3 measured in this run - MEASURED_PEAKS.json holds HBM and "
                                        "bf16 only, so the FP64 denominator is measured here (DFMA 35.7 / DMMA 37.1 TFLOP/s "
                                        "microbenchmark: profiles/r01_fp64_peak.txt)",
                         "algorithmic_flops_per_eval": svgp_flops(M, N), "kernel": "k_fit (one launch per step)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d / e2e_steps),
                    "d2h_bytes_per_step": int(d2h / e2e_steps), "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "wall_s_timed_region": t_wall,
        }
        if not args.no_cpu and world == 1:
            cores = os.cpu_count() or 1
            sample = int(args.cpu_genes or max(16, min(4 * cores, 96)))
            Ys = batches[args.warmup][0][:sample]
            gps, eps, dt = cpu_one_sample(X, Ys, M, cores)
            line["cpu_baseline"] = {"value": gps, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "first %d genes of the first timed batch, %.1f s wall; oracle port (torch-FP64 "
                                              "ELBO + SciPy L-BFGS-B), one process per core" % (sample, dt),
                                    "elbo_grad_evals_per_sec": eps}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
