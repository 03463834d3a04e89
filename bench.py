#!/usr/bin/env python
"""Benchmark of the GPcounts per-gene GP fitting hot path on B200 (contract: task prompt / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5] [--genes G] [--total-genes D] [--impl reference]

Workloads = BASELINE.json's configurations on synthetic counts of each configuration's shape (default c5, the one the
metric is quoted on):

    c5            sparse NB One_sample_test, N = 1000 spots, M = 64 inducing points (SURVEY 8d generator)   [headline]
    c5-n10k       same, N = 10000, M = 64          c5-m256   same, N = 1000, M = 256
    c1            fission-yeast shape: full VGP NB One_sample_test, N = 18 (6 time points x 3 replicates)
    c2 / c2-zinb  MouseHSC shape: sparse NB / ZINB One_sample_test, N = 2660, M = 133 (the reference's default 5 % of N)
    c3            MouseOB shape: full VGP scaled NB One_sample_test, N = 260 spots, 2-D RBF
    c4            branching shape: Infer_branching_location, N = 532, 25 candidate branching times (26 VGP fits per gene)

A *step* fits one fresh batch of G genes per GPU.  `value` = genes fitted per second over all GPUs with the batch resident in
HBM (CUDA events around the C-ABI fit calls), `e2e` = the same through the public API (`Fit_GPcounts(...)`) from host
DataFrames with the H2D / D2H copies inside the timed region, over the same K steps.
Default scaling mode is weak (every rank fits its own batches, no data-path collective; time = max over ranks).
`--total-genes D` switches to strong scaling: ONE batch of D genes is split over the ranks by the product's own sharding
(`Fit_GPcounts.distributed`, genes r::G per rank) and the all-gather of the statistics is inside both timed regions.
"""
import argparse
import hashlib
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
LIK_NAME = {"NB": "Negative_binomial", "ZINB": "Zero_inflated_negative_binomial"}


# ------------------------------------------------------------------------------------------------ synthetic inputs
def _nb_counts(rng, f, alpha):
    r = (1.0 / alpha)[:, None]
    return rng.negative_binomial(r, r / (r + np.exp(f))).astype(np.float64)


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
    return X, _nb_counts(rng, f, alpha)


def synth_fission(G, seed):
    """C1 shape: 6 time points (minutes) x 3 replicates, bulk-RNA-seq sized counts (data/fission_col_data.csv)."""
    X = np.repeat(np.array([0.0, 15.0, 30.0, 60.0, 120.0, 180.0]), 3)[:, None]
    rng = np.random.default_rng(seed)
    b = rng.uniform(np.log(20.0), np.log(5000.0), G)
    a = np.exp(rng.uniform(np.log(0.02), np.log(1.0), G))
    a[1::2] = 0.0
    om, ph = rng.uniform(0.3, 1.0, G), rng.uniform(0, 2 * np.pi, G)
    alpha = np.exp(rng.uniform(np.log(0.005), np.log(0.2), G))
    f = b[:, None] + a[:, None] * np.sin(2 * np.pi * om[:, None] * X[None, :, 0] / 180.0 + ph[:, None])
    return X, _nb_counts(rng, f, alpha)


def synth_spatial(G, N, seed):
    """C3 shape: N spots on a 2-D tissue section (coordinate range of data/MouseOB/locations.csv), per-spot scale factors."""
    X = np.random.default_rng(0).uniform(7.9, 28.0, (N, 2))
    rng = np.random.default_rng(seed)
    b = rng.uniform(np.log(0.5), np.log(20), G)
    a = np.exp(rng.uniform(np.log(0.02), np.log(1.5), G))
    a[1::2] = 0.0
    w = rng.uniform(0.1, 0.4, (G, 2))
    ph = rng.uniform(0, 2 * np.pi, G)
    alpha = np.exp(rng.uniform(np.log(0.05), np.log(1.0), G))
    f = b[:, None] + a[:, None] * np.sin(X[None, :, 0] * w[:, :1] + X[None, :, 1] * w[:, 1:] + ph[:, None])
    scale = rng.uniform(0.5, 2.0, (N, G))
    return X, _nb_counts(rng, f + np.log(scale.T), alpha), scale


def synth_branching(G, N, seed):
    """C4 shape: pseudotime + two branch labels; the branches separate after t = 0.5 for even genes."""
    rs = np.random.default_rng(0)
    t = np.sort(rs.uniform(0, 1, N))
    lab = np.where(rs.uniform(0, 1, N) < 0.5, 1.0, 2.0)
    rng = np.random.default_rng(seed)
    b = rng.uniform(np.log(2.0), np.log(50), G)
    a = rng.uniform(0.5, 2.5, G)
    a[1::2] = 0.0
    alpha = np.exp(rng.uniform(np.log(0.05), np.log(1.0), G))
    f = b[:, None] + a[:, None] * np.where((lab == 2) & (t > 0.5), t - 0.5, 0.0)[None, :]
    return t[:, None], lab, _nb_counts(rng, f, alpha)


def svgp_flops(M, N):
    """Algorithmic FLOPs of one SVGP ELBO+gradient evaluation."""
    return 6.0 * M * M * N + (4.0 / 3.0) * M ** 3 + 12.0 * M * N


def vgp_flops(N, frozen=False):
    """Full VGP: (7/3) N^3 + 6 N^2; with a frozen kernel the factorisation and its reverse pass are not part of
    an evaluation: fwd tri x tri N^3/3 + bwd 2 N^3/3 = N^3."""
    return (1.0 if frozen else 7.0 / 3.0) * N ** 3 + 6.0 * N ** 2


WORKLOADS = {
    #            kind    lik    N      M    d  genes/step  description
    "c5":       ("SVGP", "NB", 1000, 64, 1, 8192, "C5 synthetic sparse NB One_sample_test (SURVEY 8d generator): N=1000 spots, M=64 inducing"),
    "c5-n10k":  ("SVGP", "NB", 10000, 64, 1, 592, "C5 sweep: sparse NB One_sample_test, N=10000 spots, M=64 inducing"),
    "c5-m256":  ("SVGP", "NB", 1000, 256, 1, 296, "C5 sweep: sparse NB One_sample_test, N=1000 spots, M=256 inducing"),
    "c1":       ("VGP", "NB", 18, 0, 1, 6459, "C1 fission-yeast shape: full VGP NB One_sample_test, N=18 samples, 6459 genes"),
    "c2":       ("SVGP", "NB", 2660, 133, 1, 152, "C2 MouseHSC shape: sparse NB One_sample_test, N=2660 cells, M=133 inducing"),
    "c2-zinb":  ("SVGP", "ZINB", 2660, 133, 1, 152, "C2 MouseHSC shape: sparse ZINB One_sample_test, N=2660 cells, M=133 inducing"),
    "c3":       ("VGP", "NB", 260, 0, 2, 489, "C3 MouseOB shape: full VGP scaled NB One_sample_test, N=260 spots, 2-D RBF, per-spot scale"),
    "c4":       ("VGP", "NB", 532, 0, 2, 8, "C4 branching shape: Infer_branching_location, N=532 cells, 25 candidate times (26 VGP fits per gene)"),
}
BINS = 25


class Workload:
    """Inputs, device-resident fit and public-API call of one configuration."""

    def __init__(self, name, genes, seed_base):
        self.name = name
        self.kind, self.lik, self.N, self.M, self.d, g_default, self.desc = WORKLOADS[name]
        self.G = genes or g_default
        self.seed_base = seed_base
        self.sparse = self.kind == "SVGP"

    # ---- inputs
    def batch(self, step, G=None):
        G = G or self.G
        seed = self.seed_base + step
        if self.name == "c1":
            X, Y = synth_fission(G, seed)
            return dict(X=X, Y=Y)
        if self.name == "c3":
            X, Y, scale = synth_spatial(G, self.N, seed)
            return dict(X=X, Y=Y, scale=scale)
        if self.name == "c4":
            t, lab, Y = synth_branching(G, self.N, seed)
            return dict(X=t, Y=Y, labels=lab)
        X, Y = synth_batch(G, self.N, seed)
        return dict(X=X, Y=Y)

    def flops_per_eval(self, frozen=False):
        return svgp_flops(self.M, self.N) if self.sparse else vgp_flops(self.N, frozen)

    def bytes_per_gene_fit(self):
        """Algorithmic HBM bytes of one gene-fit (the N=18 regime): y + scale + theta0 in, theta-hat + stats out."""
        mv = self.M if self.sparse else self.N
        P = 4 + mv + mv * (mv + 1) // 2
        models = 26 if self.name == "c4" else 2
        return 8 * self.N * (2 if self.name == "c3" else 1) + models * (2 * 8 * P + 8 * 16)

    # ---- device-resident leg (C ABI through gpcounts_b200.engine)
    def prepare_device(self, dev, steps, opt):
        import torch
        from gpcounts_b200 import engine, host
        self.engine, self.host, self.dev, self.opt = engine, host, dev, opt
        b0 = self.batch(0, 1)
        X = b0["X"]
        self.X = X
        lik = self.lik
        if self.name == "c4":
            self.Xb = np.c_[X[:, 0], b0["labels"]]
            rtab, _ = host.restart_table(self.Xb)
            rtab = rtab.copy()
            rtab[:, 0], rtab[:, 1] = 1.0, 1.0                        # BranchKernel(RBF()) on every attempt (:606-617)
            self.free = [engine.Model("VGP", "BRANCH", lik, self.Xb, xp=-1000.0, restart=rtab, device=dev)]
            tt = np.linspace(X[:, 0].min(), X[:, 0].max(), BINS)
            self.grid = []
            for i in range(BINS):
                Xi = self.Xb.copy()
                Xi[Xi[:, 0] <= tt[i], 1] = 1
                self.grid.append(engine.Model("VGP", "BRANCH", lik, Xi, xp=float(tt[i]), train=(False, False, True, True),
                                              restart=None, device=dev))
            self.rtab = rtab
        else:
            ls0 = host.default_lengthscale(X)
            self.ls0 = ls0
            rtab, _ = host.restart_table(X)
            Z = host.inducing_sets(X, self.M, ls0) if self.sparse else None
            self.Z, self.rtab = Z, rtab
            self.models = [engine.Model(self.kind, "RBF", lik, X, Z=Z, restart=rtab, device=dev),
                           engine.Model(self.kind, "CONST", lik, X, Z=Z, restart=rtab, handoff_alpha=(lik == "NB"), device=dev)]
        self.dev_batches = []
        for s in range(steps):
            b = self.batch(s)
            Y = b["Y"]
            G = Y.shape[0]
            h = np.zeros((G, 2, 4))
            h[:, :, 0] = host.default_variance(Y, LIK_NAME[lik], True)[:, None]
            h[:, :, 1], h[:, :, 2], h[:, :, 3] = (1.0 if self.name == "c4" else self.ls0), 1.0, 35.0
            if self.name == "c4":
                h[:, :, 0] = 1.0
            d = dict(Y=torch.as_tensor(Y, device=dev), h=torch.as_tensor(h, device=dev), host=b)
            if "scale" in b:
                d["scale"] = torch.as_tensor(np.ascontiguousarray(b["scale"].T), device=dev)
            self.dev_batches.append(d)

    def fit_device(self, s):
        """One step on device-resident inputs; returns the list of stats tensors (one per launch) with a 'frozen' flag."""
        import torch
        eng = self.engine
        d = self.dev_batches[s]
        if self.name != "c4":
            st, _ = eng.fit(self.models, d["Y"], d.get("scale"), d["h"], chain=True, opt=self.opt, return_theta=False)
            return [(st, False)]
        st0, _ = eng.fit(self.free, d["Y"], None, d["h"][:, :1].contiguous(), chain=True, opt=self.opt, return_theta=False)
        G = d["Y"].shape[0]
        hg = torch.empty(G, BINS, 4, dtype=torch.float64, device=self.dev)
        hg[:, :, 0] = st0[:, 0, 5:6]                                  # fitted kernel variance / lengthscale, frozen on the grid
        hg[:, :, 1] = st0[:, 0, 6:7]
        hg[:, :, 2], hg[:, :, 3] = 1.0, 35.0
        st1, _ = eng.fit(self.grid, d["Y"], None, hg, chain=False, opt=self.opt, return_theta=False)
        return [(st0, False), (st1, True)]

    # ---- public API leg
    def frames(self, b):
        import pandas as pd
        cells = ["c%d" % i for i in range(b["X"].shape[0])]
        genes = ["g%d" % i for i in range(b["Y"].shape[0])]
        Xdf = pd.DataFrame(b["X"], index=cells)
        Ydf = pd.DataFrame(b["Y"], index=genes, columns=cells)
        sdf = pd.DataFrame(b["scale"], index=cells, columns=genes) if "scale" in b else None
        return Xdf, Ydf, sdf

    def fit_api(self, b, dev, distributed, maxiter):
        """The call a user makes.  Returns (a host-side scalar read from the result, h2d bytes, d2h bytes)."""
        from gpcounts_b200 import Fit_GPcounts
        Xdf, Ydf, sdf = self.frames(b)
        Y = b["Y"]
        G = Y.shape[0]
        mv = self.M if self.sparse else self.N
        P = 4 + mv + mv * (mv + 1) // 2
        if self.name == "c4":
            tot, h2d, d2h = 0.0, 0, 0
            for g in range(G):                                       # the reference's branching API is per gene (:264-265)
                gp = Fit_GPcounts(Xdf, Ydf.iloc[[g]], device=dev)
                gp.distributed = False
                gp.optimizer_options = {"maxiter": maxiter}
                res = gp.Infer_branching_location(b["labels"], bins_num=BINS)
                tot += float(res["logBayesFactor"])
                h2d += Y[g].nbytes + (1 + BINS) * (b["X"].shape[0] * 2 * 8 + 4 * 8)
                d2h += (1 + BINS) * 16 * 8 + 2 * P * 8
            return tot, h2d, d2h
        gp = Fit_GPcounts(Xdf, Ydf, scale=sdf, sparse=self.sparse, M=self.M, device=dev)
        gp.distributed = distributed
        gp.optimizer_options = {"maxiter": maxiter}
        df = gp.One_sample_test(LIK_NAME[self.lik])
        val = float(np.nansum(df["log_likelihood_ratio"].values))    # result read on the host
        from gpcounts_b200 import sharding
        rk, wd = sharding.rank_world() if distributed else (0, 1)
        shard = len(range(rk, G, wd))
        h2d = shard * (Y.shape[1] * 8 * (2 if sdf is not None else 1) + 2 * 4 * 8) + b["X"].nbytes * 2
        if self.sparse:
            h2d += 2 * (11 * self.M * self.d * 8 + 10 * 4 * 8)
        d2h = G * 2 * 16 * 8 + P * 8
        return val, h2d, d2h

    # ---- CPU oracle port (bench baseline / reference arm)
    def cpu_fit(self, b, n_jobs):
        from joblib import Parallel, delayed
        X, Y = b["X"], b["Y"]
        scale = b.get("scale")
        lik_name = LIK_NAME[self.lik]
        sparse, M, name = self.sparse, self.M, self.name

        def one(g):
            import torch
            torch.set_num_threads(1)
            from oracle.fit import OracleFit
            o = OracleFit(X, Y, scale=scale, sparse=sparse, M=M, inducing_variance=1.0 if sparse else None)
            r = o.one_sample_test(lik_name, genes=[g])[0]
            return sum(f.nfev for f in r["fits"] if f is not None)

        CPU_BINS = 4      # bounded sample: the free fit + 4 of the 25 candidate times of ONE gene, scaled to 26 fits per gene

        def one_branching(g):
            import torch
            torch.set_num_threads(max(1, n_jobs))
            from oracle.fit import OracleFit
            o = OracleFit(X, Y)
            r = o.infer_branching_location(b["labels"], bins_num=CPU_BINS, gene=g)
            return r["free_fit"].nfev + sum(f.nfev for f in r["fits"])

        t0 = time.time()
        if name == "c4":
            nfev = [one_branching(g) for g in range(Y.shape[0])]
        else:
            nfev = Parallel(n_jobs=n_jobs)(delayed(one)(g) for g in range(Y.shape[0]))
        dt = time.time() - t0
        if name == "c4":
            return Y.shape[0] * (1 + CPU_BINS) / (1.0 + BINS) / dt, float(np.sum(nfev)) / dt, dt
        return Y.shape[0] / dt, float(np.sum(nfev)) / dt, dt

    def cpu_sample_genes(self, cores):
        per_gene_core_s = {"c5": 1.5, "c5-n10k": 20.0, "c5-m256": 8.0, "c1": 2.0, "c2": 12.0, "c2-zinb": 12.0, "c3": 8.0}
        if self.name == "c4":
            return 1
        return int(max(cores, min(96, round(20.0 * cores / per_gene_core_s[self.name]))))

    def config(self, args, strong):
        l2 = ("a fresh gene batch every step; 256 MB written between timed steps to flush the 126 MB L2; the per-slot "
              "workspaces (factors, optimiser vectors, L-BFGS memory) are rewritten by every evaluation")
        c = {"workload": self.desc, "name": self.name, "N": self.N, "M": self.M, "d": self.d,
             "test": "Infer_branching_location (25 bins)" if self.name == "c4" else "One_sample_test",
             "likelihood": LIK_NAME[self.lik], "models_per_gene": 26 if self.name == "c4" else 2,
             "optimizer": "L-BFGS-B m=10 ftol=2.22e-9 gtol=1e-5 maxiter=5000 (SciPy defaults of the reference)",
             "l2_policy": l2}
        if strong:
            c["total_genes"] = args.total_genes
        return c


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


def dist_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  GPflow / TensorFlow / RobustGP are not
    installable here (no wheels, no network - DESIGN.md), so this is the oracle port: torch-FP64 restatement of the
    GPflow ELBO + the installed SciPy L-BFGS-B, one process per host core, on a bounded sample of the workload per step."""
    rank, world, _ = dist_info()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    wl = Workload(args.workload, 0, 1234)
    sample = int(args.cpu_genes or wl.cpu_sample_genes(cores))
    times, evals = [], []
    for s in range(args.warmup + args.steps):
        gps, eps, dt = wl.cpu_fit(wl.batch(s, sample), cores)
        if s >= args.warmup:
            times.append(dt); evals.append(eps)
    T = float(np.sum(times))
    value = sample * len(times) / T
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T / len(times), "higher_is_better": True,
        "scaling": "strong" if args.total_genes else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": wl.config(args, bool(args.total_genes)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d genes per step of the workload, oracle port: torch-FP64 ELBO + SciPy L-BFGS-B, one "
                                   "process per core" % sample, "elbo_grad_evals_per_sec": float(np.mean(evals))},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "genes_per_step": sample}))


def _lib_hash():
    p = os.path.join(ROOT, "gpcounts_b200", "lib", "libgpcounts_b200.so")
    return hashlib.sha1(open(p, "rb").read()).hexdigest()[:12] if os.path.exists(p) else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--genes", type=int, default=0, help="genes per step per GPU (0 = the workload's default)")
    ap.add_argument("--total-genes", type=int, default=0, help="strong scaling: ONE batch of this many genes over all GPUs")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-genes", type=int, default=0, help="genes in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--maxiter", type=int, default=5000, help="profiling only: cap L-BFGS iterations")
    ap.add_argument("--variant", type=int, default=-1, help="A/B measurements only: force a kernel instantiation (gpc_opt_t.variant)")
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
    from gpcounts_b200 import engine, sharding
    from gpcounts_b200._lib import S_NFEV_TOTAL, S_STATUS

    strong = args.total_genes > 0
    nsteps = args.warmup + args.steps
    if strong:
        wl = Workload(args.workload, args.total_genes, 1234)          # every rank generates the same batch ...
        full = [wl.batch(s) for s in range(nsteps)]
        mine = sharding.shard_indices(args.total_genes, rank, world)   # ... and keeps the product's shard r::G of it
        wl.batch = lambda s, G=None, _f=full, _m=mine: {k: (v[_m] if k == "Y" else (v[:, _m] if k == "scale" else v))
                                                         for k, v in _f[s].items()}
        G = args.total_genes
    else:
        wl = Workload(args.workload, args.genes, 1234 + 1000 * rank)
        G = wl.G
    opt = engine.default_opt(maxiter=args.maxiter, variant=args.variant)
    wl.prepare_device(dev, nsteps, opt)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident leg ------------------------------------------------------------------------------------------
    sampler = ClockSampler(local)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
    stats_all, launches0, t_wall0 = [], 0, time.time()
    for s in range(nsteps):
        if s == args.warmup:
            barrier()
            sampler.start()
            launches0 = engine.launch_count()
            t_wall0 = time.time()
        flush.fill_(s & 0xFF)                                          # L2 flush between timed steps (not timed)
        ev[s][0].record()
        out = wl.fit_device(s)
        if strong:                                                     # the product's one collective, inside the timed region
            gathered = [sharding.gather_rows(st, G, rank, world) for st, _ in out]
        ev[s][1].record()
        stats_all.append(out)
    barrier()
    t_wall = time.time() - t_wall0
    clocks = sampler.stop()
    launches = engine.launch_count() - launches0
    kern_ms = [ev[s][0].elapsed_time(ev[s][1]) for s in range(args.warmup, nsteps)]
    T_local = float(np.sum(kern_ms)) * 1e-3
    T = T_local
    if world > 1:
        t = torch.tensor([T], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        T = float(t.item())
    genes_timed = (G if strong else G * world) * args.steps
    value = genes_timed / T

    nfev = flops = eval_cyc = fit_cyc = 0.0
    n_ok = n_fits = 0
    for out in stats_all[args.warmup:]:
        for st, frozen in out:
            a = st.cpu().numpy()
            ne = a[..., S_NFEV_TOTAL].sum()
            nfev += ne
            flops += ne * wl.flops_per_eval(frozen)
            n_ok += int((a[..., S_STATUS] == 0).sum()); n_fits += a[..., 0].size
            eval_cyc += a[..., 12].sum(); fit_cyc += a[..., 13].sum()

    # ---- end-to-end leg: the public API from host DataFrames, same K steps ---------------------------------------------
    e2e_T, h2d, d2h = 0.0, 0, 0
    barrier()
    for s in range(args.warmup, nsteps):
        b = full[s] if strong else wl.dev_batches[s]["host"]
        flush.fill_(s & 0xFF)
        barrier()
        t0 = time.time()
        _val, bi, bo = wl.fit_api(b, dev, strong and world > 1, args.maxiter)
        torch.cuda.synchronize(dev)
        e2e_T += time.time() - t0
        h2d += bi; d2h += bo
    if world > 1:
        t = torch.tensor([e2e_T], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_T = float(t.item())
    e2e_value = genes_timed / e2e_T

    if rank == 0:
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        n_slots = min(engine.default_slots(wl.free if wl.name == "c4" else wl.models), max(1, G // world if strong else G))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": wl.config(args, strong),
            "genes_per_step": G if strong else G * world,
            "gene_fits_per_hour_per_gpu": value * 3600 / world,
            "elbo_grad_evals_per_sec": float(nfev) * (1 if strong else world) / T,
            "evals_per_gene": float(nfev) / (G / world if strong else G) / args.steps, "fit_ok_fraction": n_ok / max(n_fits, 1),
            "time_split": {"eval_share_of_fit_cycles": float(eval_cyc / max(fit_cyc, 1.0)),
                           "us_per_eval_in_fit": float(eval_cyc / max(nfev, 1.0) / sm_mhz),
                           "slot_busy_fraction": float(fit_cyc / (sm_mhz * 1e3 * float(np.sum(kern_ms)) * n_slots))},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d / args.steps),
                    "d2h_bytes_per_step": int(d2h / args.steps), "steps": args.steps,
                    "collective_in_timed_region": bool(strong and world > 1)},
            "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": t_wall,
        }
        # roofline of the dominant kernel (k_fit): rank 0's own launches, CUDA events on the launching stream
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        if wl.name == "c1":
            # latency / HBM regime: algorithmic bytes per gene-fit x genes of this rank / kernel time
            genes_rank = (len(mine) if strong else G) * args.steps
            ach = wl.bytes_per_gene_fit() * genes_rank / T_local * 1e-9
            peak = float(peaks.get("hbm_gbs", 6458.4))
            line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                "traffic": None, "algorithmic_bytes_per_gene_fit": wl.bytes_per_gene_fit(),
                                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6458.4 GB/s",
                                "fp64_tflops_algorithmic": flops / T_local * 1e-12 / (1 if strong else 1),
                                "kernel": "k_fit (one launch per step)"}
        else:
            a = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
            for _ in range(2):
                a @ a
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); a @ a; e1.record(); torch.cuda.synchronize(dev)
                best = min(best, e0.elapsed_time(e1))
            dgemm_tf = 2 * 4096 ** 3 / best * 1e-9
            achieved = flops / T_local * 1e-12
            traffic, tsrc = None, None
            tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
            if os.path.exists(tp):       # ncu dram bytes per gene of THIS build of the library (hash-checked), else null
                tj = json.load(open(tp)).get(wl.name)
                if tj and tj.get("lib_sha1") == _lib_hash():
                    traffic = tj["dram_bytes_per_gene"] * (len(mine) if strong else G)
                    tsrc = tj.get("source")
            line["roofline"] = {"bound": "tensor", "achieved": achieved, "peak": dgemm_tf, "unit": "TFLOP/s",
                                "frac": achieved / dgemm_tf, "traffic": traffic, "traffic_source": tsrc,
                                "pipe": "FP64 tensor pipe (DMMA.8x8x4; tcgen05 has no FP64 kind)",
                                "peak_source": "FP64: cuBLAS DGEMM 4096^3 measured in this run - MEASURED_PEAKS.json holds HBM "
                                               "and bf16 only (DFMA 35.7 / DMMA 37.1 TFLOP/s microbenchmark: "
                                               "profiles/r01_fp64_peak.txt)",
                                "algorithmic_flops_per_eval": wl.flops_per_eval(),
                                "kernel": "k_fit (%d launch%s per step)" % (2 if wl.name == "c4" else 1, "es" if wl.name == "c4" else "")}
        if not args.no_cpu and world == 1:
            cores = os.cpu_count() or 1
            sample = int(args.cpu_genes or wl.cpu_sample_genes(cores))
            b = wl.batch(args.warmup, sample) if not strong else {k: (v[:sample] if k == "Y" else (v[:, :sample] if k == "scale" else v))
                                                                   for k, v in full[args.warmup].items()}
            gps, eps, dt = wl.cpu_fit(b, cores)
            line["cpu_baseline"] = {"value": gps, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d genes of the workload, %.1f s wall; oracle port (torch-FP64 ELBO + SciPy "
                                              "L-BFGS-B), one process per core" % (sample, dt),
                                    "elbo_grad_evals_per_sec": eps}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
