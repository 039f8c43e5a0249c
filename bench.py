#!/usr/bin/env python
"""Benchmark of the SIMCA hot path on B200 (contract: see the task's bench.py section).

Headline workload = BASELINE.json configs[2]: bulk scoring of 10 M x 1,024 fp32 synthetic
spectra against a fixed 10-PC SIMCA model, fused T2+Q+decision, rows sharded over the GPUs
(weak scaling: every rank scores its own 10 M-row shard, no data-path collective).
A "step" is one pass of the fused kernel over the rank's shard (41 GB >> 126 MB L2, so
every step streams from HBM; no explicit L2 flush is needed).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
  python bench.py --impl reference ...                           # CPU oracle arm (reference algorithm)

Extra keys beyond the base contract: `roofline`, `cpu_baseline`, `e2e`, `clocks`, `gpu_launches`, and the legs
  `strong` -- the same 10 M rows in TOTAL split over the N GPUs (configs[2] read as strong scaling),
  `fit`    -- BASELINE configs[3]: row-sharded PCA fit (column statistics, integer tcgen05 Gram, NCCL all-reduce of the
              partial moments, eigensolve) at 2 M x 2048 rows per GPU, tensor roofline,
  `vae`    -- BASELINE configs[4]: ConvVAE1D (latent 16) bf16 DDP training step + row-sharded fused latent-T2 /
              reconstruction-Q scoring of 1 M x 1024 spectra,
  `cv`     -- BASELINE.json's second metric: CVSIMCA fits/s on configs[1].
The CPU arms time the UNMODIFIED reference staged in baseline/_ref (kind "reference"); when that directory is absent
(nothing staged) they fall back to the oracle's restatement (kind "port").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 to every worker unless the user set it.  Rank 0 also times the CPU arms (the
# reference's numpy / OpenBLAS arithmetic "with all the host threads it can use"), and OpenBLAS sizes its pool when numpy
# is imported -- so rank 0 gets the cores back BEFORE that import.  The other ranks keep torchrun's setting.
if os.environ.get("TORCHELASTIC_RUN_ID") and os.environ.get("RANK", "0") == "0":
    _n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(_n)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_SPECTRUM = lambda p: 4 * p + 4 + 4 + 1     # read x, write T2, Q, decision
FALLBACK_HBM_GBS = 6650.0                           # B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000, help="rows per GPU (weak scaling)")
    ap.add_argument("--p", type=int, default=1024)
    ap.add_argument("--A", type=int, default=10)
    ap.add_argument("--e2e-rows", type=int, default=1_000_000)
    ap.add_argument("--cpu-rows", type=int, default=200_000)
    ap.add_argument("--fit-rows", type=int, default=2_000_000, help="rows per GPU of the PCA-fit leg")
    ap.add_argument("--fit-p", type=int, default=2048)
    ap.add_argument("--no-fit", action="store_true")
    ap.add_argument("--vae-rows", type=int, default=1_000_000, help="rows in TOTAL of the VAE-SIMCA leg (configs[4])")
    ap.add_argument("--no-vae", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-cv", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--force-simt", action="store_true", help="time the FP32-pipe kernel instead of the dispatcher's choice")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_tensor_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["bf16_tflops_sustained"]), float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained / bf16_tflops)"
    except Exception:
        return 1400.0, 1590.0, "fallback (B200_PROFILING.md: 1.59 PFLOP/s burst, ~1.4 sustained)"


def reference_modules():
    """The unmodified reference staged in baseline/_ref, or None (then the oracle port is timed)."""
    try:
        from baseline import ref_loader
        if ref_loader.available():
            return ref_loader.load()
    except Exception:
        pass
    return None


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, BEFORE the pinned staging buffers are
    allocated (first touch puts their pages on that node).  Returns a description for the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(torch.cuda.get_device_properties(index), "pci_bus_id") else None
        dom = getattr(torch.cuda.get_device_properties(index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(index), "pci_device_id", 0)
        if bus is None:
            return {"numa": "unknown (no pci id)"}
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return {"numa": "single node / not reported"}
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            ids.update(range(int(lo), int(hi or lo) + 1))
        allowed = ids & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus_bound": len(allowed)}
    except Exception as e:
        return {"numa": f"not bound ({type(e).__name__})"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# model used by both arms: 10-PC SIMCA fitted on 5,000 synthetic rows (host fp64 arrays)
# ------------------------------------------------------------------------------------------------
def fit_bench_model_gpu(p, A, device):
    import ocm_vae_simca_b200 as ocm
    from ocm_vae_simca_b200 import engine, limits, synth
    Xfit = synth.make_spectra_device(5000, p, device, seed=99)
    pca = engine.fit_pca(Xfit, None, A)
    lim = limits.all_limits("alt", "Fdist", 0.95, "jm", 0.95, 0.95, A, pca.n_samples, pca.eigs_all)
    return pca, lim


def _all_cores():
    """BLAS / OpenMP pools opened to every host core for the CPU arms (torchrun exports OMP_NUM_THREADS=1, which would
    silently time a one-thread baseline).  Returns (context manager, number of threads the BLAS pools will use)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        ctx = threadpool_limits(limits=n)
        used = max([int(i.get("num_threads", 1)) for i in threadpool_info() if i.get("user_api") == "blas"] or [1])
        return ctx, used                      # what the BLAS pool really has after the request
    except Exception:
        import contextlib
        return contextlib.nullcontext(), int(os.environ.get("OMP_NUM_THREADS", n))


def _cpu_predict_runner(args):
    """Returns (predict(X) -> decisions, kind, description): the staged reference's own SIMCA.fit / predict
    (utils/SIMCA.py), else the oracle port.  fp32 rows, as in the GPU workload."""
    from ocm_vae_simca_b200 import synth
    p, A = args.p, args.A
    Xfit = synth.make_spectra(5000, p, seed=99).astype(np.float32)
    ref = reference_modules()
    if ref is not None:
        import contextlib, io
        np.random.seed(0)                      # the reference's second PCA is an unseeded randomized SVD
        m = ref[0].SIMCA(n_components=A, model_class=0, verbose=False)
        with contextlib.redirect_stdout(io.StringIO()):
            m.fit(Xfit, np.zeros(Xfit.shape[0], dtype=np.int64))

        def run(X):
            with contextlib.redirect_stdout(io.StringIO()):
                return m.predict(X)
        return run, "reference", "baseline/_ref utils/SIMCA.py SIMCA.predict (unmodified reference, numpy/sklearn/OpenBLAS)"
    from oracle import simca_oracle as so
    model = so.fit_one_class(Xfit.astype(np.float64), A, faithful_solver=True)
    return (lambda X: so.predict(model, X)), "port", "oracle.simca_oracle.predict (restatement; baseline/_ref not staged)"


def reference_arm(args, rank):
    """CPU arm: the reference's own SIMCA.predict on the box's host cores (all threads), on a bounded sample of the
    same workload per step.  Rank 0 only."""
    if rank != 0:
        return
    from ocm_vae_simca_b200 import synth
    p, rows = args.p, args.cpu_rows
    run, kind, what = _cpu_predict_runner(args)
    X = synth.make_spectra(rows, p, seed=123).astype(np.float32)
    pool, cores = _all_cores()
    with pool:
        for _ in range(max(args.warmup, 1)):
            run(X[: min(rows, 20000)])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run(X)
        dt = time.perf_counter() - t0
    val = rows * args.steps / dt
    line = {
        "impl": "reference", "metric": "spectra scored/s (fused T2+Q)", "value": val, "unit": "spectra/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": val, "unit": "spectra/s", "cores": cores, "kind": kind,
                         "sample": f"{rows} x {p} fp32 rows per step, {what}, {cores} threads"},
        "e2e": {"value": val, "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": f"bulk SIMCA scoring: {args.rows} x {args.p} fp32 synthetic spectra per GPU vs fixed "
                        f"{args.A}-PC model (BASELINE configs[2]), fused T2+Q+decision",
            "rows_per_gpu": args.rows, "p": args.p, "n_components": args.A, "rule": "alt/Fdist/jm",
            "sharding": "rows, no data-path collective",
            "l2": "inputs (4 KB/row x rows) exceed the 126 MB L2; no flush needed"}


def cpu_baseline_leg(args):
    from ocm_vae_simca_b200 import synth
    p, rows = args.p, args.cpu_rows
    run, kind, what = _cpu_predict_runner(args)
    X = synth.make_spectra(rows, p, seed=123).astype(np.float32)
    pool, cores = _all_cores()
    with pool:
        run(X[:20000])
        best = float("inf")
        for _ in range(3):
            t0 = time.perf_counter()
            run(X)
            best = min(best, time.perf_counter() - t0)
    return {"value": rows / best, "unit": "spectra/s", "cores": cores, "kind": kind,
            "sample": f"{rows} x {p} fp32 rows, best of 3, {what}, {cores} threads"}


def fit_leg(args, device, group, rank, world):
    """BASELINE configs[3] per GPU: rows sharded, partial {Gram, sums, count} all-reduced (NCCL), eigensolve on every
    rank.  Device-timed phases (CUDA events), max over ranks of the total."""
    import torch
    from ocm_vae_simca_b200 import dist as odist
    from ocm_vae_simca_b200 import engine, native, synth
    rows, p, k = args.fit_rows, args.fit_p, 20
    X = torch.empty((rows, p), dtype=torch.float32, device=device)
    synth.make_spectra_device(rows, p, device, seed=2000 + rank, out=X)
    engine.fit_pca(X, None, k, group=group)                     # warm-up (workspace, NCCL channels, cuTensorMap encoder)
    engine.fit_pca(X, None, k, group=group)
    torch.cuda.synchronize()
    if group is not None:
        import torch.distributed as dist
        dist.barrier()
    reps, totals, phases = 3, [], None
    l0 = all_launches()
    for _ in range(reps):
        tm = engine.PhaseTimer()
        pca = engine.fit_pca(X, None, k, group=group, timer=tm)
        ph = tm.result()
        totals.append(sum(ph.values()))
        if phases is None or totals[-1] == min(totals):
            phases = ph
    launches = (all_launches() - l0) // reps
    total_ms = odist.max_over_ranks(min(totals), device)
    gram_ms = odist.max_over_ranks(phases.get("gram", float("nan")), device)
    del X
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    sus, burst, src = measured_tensor_peak()
    flop = rows * p * (p + 1.0)                                  # symmetric rank-k update, per GPU
    achieved = flop / (gram_ms * 1e-3) / 1e12
    out = {"metric": "PCA fit rows/s (Gram + all-reduce + eigensolve)", "value": world * rows / (total_ms * 1e-3), "unit": "rows/s",
           "n_gpus": world, "rows_per_gpu": rows, "p": p, "n_components": k, "scaling": "weak",
           "ms": {"total_max_over_ranks": total_ms, **{k_: round(v, 3) for k_, v in phases.items()}},
           "gram": {"kernel": "gtc::gram_tc_kernel (tcgen05.mma kind::i8, int32 TMEM accumulators) + gram_quantize_kernel",
                    "digit_products": engine.GRAM_TC_PRODUCTS, "ms": gram_ms, "algorithmic_tflops_per_gpu": achieved,
                    "physical_int8_tops_per_gpu": achieved * engine.GRAM_TC_PRODUCTS * (p / 128 + 1) / (p / 128) },
           "roofline": {"bound": "tensor", "achieved": achieved, "peak": sus, "unit": "TFLOP/s", "frac": achieved / sus,
                        "frac_of_burst_peak": achieved / burst, "peak_source": src,
                        "algorithmic_flop_per_launch_series": flop, "traffic": None,
                        "note": "algorithmic flops n*p*(p+1) against the measured dense bf16 peak; the kernel issues "
                                f"{engine.GRAM_TC_PRODUCTS} int8 digit products per output element (not credited)"},
           "allreduce_payload_bytes": (p * p + 2 * p + 3) * 8 if world > 1 else 0,
           "gpu_launches": int(launches),
           "check": {"lambda_1": float(pca.eigs_all[0]), "lambda_20": float(pca.eigs_all[19])}}
    # CPU: the reference's SIMCA.fit (two sklearn PCA fits + distances + limits) on a bounded sample of the same rows
    try:
        ref = reference_modules()
        crow = 20000
        Xc = synth.make_spectra(crow, p, seed=2000).astype(np.float32)
        pool, cores = _all_cores()
        with pool:
            t0 = time.perf_counter()
            if ref is not None:
                import contextlib, io
                np.random.seed(0)
                with contextlib.redirect_stdout(io.StringIO()):
                    ref[0].SIMCA(n_components=k, model_class=0, verbose=False).fit(Xc, np.zeros(crow, dtype=np.int64))
                kind, what = "reference", "baseline/_ref utils/SIMCA.py SIMCA.fit"
            else:
                from oracle import simca_oracle as so
                so.fit_one_class(Xc.astype(np.float64), k, faithful_solver=True)
                kind, what = "port", "oracle.simca_oracle.fit_one_class"
            cdt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": crow / cdt, "unit": "rows/s", "cores": cores, "kind": kind,
                               "sample": f"one {what} on {crow} x {p} fp32 rows ({cdt:.1f} s)"}
    except Exception as e:
        out["cpu_baseline"] = {"error": repr(e)}
    return out


def all_launches():
    """kernel launches issued through the library so far: explicit ones plus those replayed from its CUDA graphs"""
    from ocm_vae_simca_b200 import engine, native
    return native.launch_count() + engine.graph_replayed_launches


def cv_leg(device, group=None, rank=0, world=1):
    """BASELINE configs[1]: CVSIMCA venetian blinds 10-fold x 1..20 PCs on 2000 x 1050 (+ refit).  With several ranks
    the folds are dealt round-robin over the GPUs (every rank holds the 8.4 MB X; two small all-reduces merge the
    confusion counts and the pooled decisions); the time is the max over ranks."""
    import torch
    from ocm_vae_simca_b200 import dist as odist
    from sklearn.model_selection import PredefinedSplit
    from ocm_vae_simca_b200 import native, synth
    from ocm_vae_simca_b200.utils import SIMCA, cross_validate_simca_grid
    X, y = synth.make_class_dataset(1500, 500, 1050, seed=1234)
    cv = PredefinedSplit(np.arange(X.shape[0]) % 10)
    est = SIMCA(model_class=0, verbose=False)
    run = lambda: cross_validate_simca_grid(est, X, y, cv, LV_min=1, LV_max=20, class_index=0, print_summary=False,
                                            group=group)
    run()
    torch.cuda.synchronize()
    if group is not None:
        import torch.distributed as dist
        dist.barrier()
    l0 = all_launches()
    t0 = time.perf_counter()
    res = run()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if group is not None:
        dt = odist.max_over_ranks(dt, device)
    if rank != 0:
        return None
    fits = 10 * 20 + 1
    out = {"metric": "CVSIMCA fits/s", "value": fits / dt, "unit": "fits/s", "seconds": dt, "fits": fits,
           "config": "2000 x 1050 fp64 host arrays, 1500 target / 500 other, PredefinedSplit(i % 10), LV 1..20, refit",
           "n_gpus": world, "sharding": "folds dealt round-robin over ranks" if world > 1 else "one GPU",
           "best_LV": int(res["best_LV"]), "gpu_launches": all_launches() - l0,
           "timing": "host wall clock around the public call (includes H2D of X and host limit arithmetic)"}
    # CPU: the reference's own per-(fold, A) cell (SIMCA.fit on the fold's training rows + predict on its test rows,
    # what CVSIMCA.py:183-190 does 200 times), bounded to 3 cells, reported per fit
    try:
        tr, te = next(iter(cv.split()))
        ref = reference_modules()
        pool, cores = _all_cores()
        with pool:
            t0 = time.perf_counter()
            cells = 0
            for A in (5, 10, 15):
                if ref is not None:
                    import contextlib, io
                    np.random.seed(0)
                    with contextlib.redirect_stdout(io.StringIO()):
                        m = ref[0].SIMCA(n_components=A, model_class=0, verbose=False).fit(X[tr], y[tr])
                        m.predict(X[te])
                else:
                    from oracle import simca_oracle as so
                    m = so.fit_one_class(X[tr][y[tr] == 0], A, faithful_solver=True)
                    so.predict(m, X[te])
                cells += 1
            cdt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": cells / cdt, "unit": "fits/s", "cores": cores,
                               "kind": "reference" if ref is not None else "port",
                               "sample": "3 (fold, A) cells of the same grid: SIMCA.fit on the fold + predict on its test rows"}
    except Exception as e:  # the CPU arm is a reported baseline; its absence must not kill the GPU line
        out["cpu_baseline"] = {"error": repr(e)}
    return out


def vae_leg(args, device, group, rank, world, local_rank):
    """BASELINE configs[4]: VAE-SIMCA on 1 M x 1024 spectra (latent 16), rows sharded over the GPUs.
    (a) training: ConvVAE1D (the reference architecture, /root/reference/vae_model.py:6-129, 17 M parameters) under
        DistributedDataParallel, bf16 autocast, fused Adam, cosine beta-VAE loss (vae_bce_nut.py:178-192 trains in fp32
        at batch 512, one GPU); the step time includes the NCCL gradient all-reduce;
    (b) scoring: one encoder pass + deterministic decode per batch in bf16, then ONE fused kernel per batch
        (ocm_latent_score_f32: latent T2 against the calibration mean / covariance + reconstruction Q from the bf16
        reconstruction), the whole shard per step (final_vaesimca.py:471-533 does this with a double forward pass and
        per-batch numpy on the host)."""
    import torch
    from ocm_vae_simca_b200 import dist as odist
    from ocm_vae_simca_b200 import engine, native, synth
    from ocm_vae_simca_b200.vae import ConvVAE1D
    from ocm_vae_simca_b200.vae.convvae import beta_vae_cosine_loss
    from ocm_vae_simca_b200.vae.vae_simca import latent_recon_scores
    L, d, rows_total = 1024, 16, args.vae_rows
    lo, hi = odist.shard_rows(rows_total, rank, world)
    n = hi - lo
    X = synth.make_spectra_device(n, L, device, seed=3000 + rank)
    mean = X[:65536].mean(0); std = X[:65536].std(0) + 1e-8
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(mean, group=group); mean /= world
        dist.all_reduce(std, group=group); std /= world
    torch.manual_seed(0)
    vae = ConvVAE1D(L, d, mean.cpu().numpy(), std.cpu().numpy()).to(device)
    vae.channels_last = True          # same parameters and arithmetic, 2-D channels-last execution (convvae.py)
    model = vae
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(vae, device_ids=[local_rank], gradient_as_bucket_view=True)
    opt = torch.optim.Adam(vae.parameters(), lr=1e-4, fused=True)
    Bt = 4096

    def train_step(i):
        xb = X[(i * Bt) % (n - Bt):][:Bt]
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            xr, mu_, lv_ = model(xb)
            loss = beta_vae_cosine_loss(xb, xr.float(), mu_.float(), lv_.float(), as_floats=False)[0]
        loss.backward()
        opt.step()
        return loss
    vae.train()
    for i in range(3):
        train_step(i)
    torch.cuda.synchronize()
    if group is not None:
        dist.barrier()
    tsteps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(tsteps):
        loss = train_step(3 + i)
    e1.record()
    torch.cuda.synchronize()
    train_ms = odist.max_over_ranks(e0.elapsed_time(e1), device) / tsteps
    final_loss = float(loss.detach())

    # calibration statistics of the latent means (VAE_SIMCA.py:245-248), all-reduced over the shards
    vae.eval()
    Bs = 8192
    with torch.no_grad():                                     # calibration statistics in fp32
        mu_cal = torch.cat([vae.encode((X[r:r + Bs] - vae.spec_mean) / vae.spec_std)[0].float() for r in range(0, min(n, 65536), Bs)])
    mu_cal = torch.nan_to_num(mu_cal)
    cnt = torch.tensor([float(mu_cal.shape[0])], dtype=torch.float64, device=device)
    s1 = mu_cal.double().sum(0); s2 = mu_cal.double().T @ mu_cal.double()
    if group is not None:
        for t in (cnt, s1, s2):
            dist.all_reduce(t, group=group)
    m = s1 / cnt
    cov = (s2 - cnt * torch.outer(m, m)) / (cnt - 1) + 1e-6 * torch.eye(d, dtype=torch.float64, device=device)
    vae.latent_mean.copy_(m.float()); vae.latent_cov_inv.copy_(torch.linalg.pinv(cov, hermitian=True).float())

    d2_all = torch.empty(n, device=device); q_all = torch.empty(n, device=device)

    from ocm_vae_simca_b200.vae import FrozenScorer
    scorer = FrozenScorer(vae)            # bf16 inference copy of the trained module: BatchNorm folded, channels-last convolutions

    def score_pass():
        for r in range(0, n, Bs):
            xb = X[r:r + Bs]
            d2, q, _ = scorer(xb)
            d2_all[r:r + xb.shape[0]] = d2; q_all[r:r + xb.shape[0]] = q
    score_pass()
    torch.cuda.synchronize()
    if group is not None:
        dist.barrier()
    l0 = native.launch_count()
    e0.record()
    score_pass()
    e1.record()
    torch.cuda.synchronize()
    launches = native.launch_count() - l0
    score_ms = odist.max_over_ranks(e0.elapsed_time(e1), device)

    # the fused epilogue alone on resident tensors of one pass worth of rows (its HBM roofline)
    ne = min(n, 262144)
    z = torch.randn(ne, d, device=device); xr = X[:ne].to(torch.bfloat16)
    for _ in range(2):
        engine.latent_score(z=z, m=vae.latent_mean, Sinv=vae.latent_cov_inv, x=X[:ne], xrec=xr)
    e0.record()
    for _ in range(5):
        engine.latent_score(z=z, m=vae.latent_mean, Sinv=vae.latent_cov_inv, x=X[:ne], xrec=xr)
    e1.record()
    torch.cuda.synchronize()
    epi_ms = e0.elapsed_time(e1) / 5
    epi_bytes = ne * (4 * L + 2 * L + 4 * d + 8)
    acc_frac = float(((d2_all <= torch.quantile(d2_all[:100000], 0.95)) & (q_all <= torch.quantile(q_all[:100000], 0.95))).float().mean())
    del X
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    hbm, _ = measured_peak()
    return {"metric": "VAE-SIMCA spectra/s (ConvVAE1D latent 16, bf16)", "n_gpus": world, "rows_total": rows_total, "rows_per_gpu": n,
            "scaling": "strong (1 M rows in total, sharded)",
            "scoring": {"value": rows_total / (score_ms * 1e-3), "unit": "spectra/s", "ms_per_pass": score_ms, "batch": Bs,
                        "what": "bf16 encoder + decoder through vae.FrozenScorer (GEMM-shaped layers: PyTorch cuDNN/cuBLAS on channels-last "
                                "activations, BatchNorm folded, weights cast once; first block and every bias + activation: kernels of "
                                "this library) + fused latent T2 / reconstruction Q kernel, whole shard",
                        "gpu_launches": int(launches), "accepted_frac_p95_rule": acc_frac},
            "training": {"value": world * Bt / (train_ms * 1e-3), "unit": "spectra/s", "ms_per_step": train_ms, "batch_per_gpu": Bt,
                         "what": "DistributedDataParallel bf16-autocast step (fwd + bwd + NCCL gradient all-reduce + fused Adam), channels-last execution" if world > 1
                                 else "bf16-autocast step (fwd + bwd + fused Adam), channels-last execution",
                         "parameters": int(sum(p_.numel() for p_ in vae.parameters())), "loss_after": final_loss},
            "epilogue_roofline": {"kernel": "latent_score_kernel<bf16 xrec>", "bound": "hbm", "achieved": epi_bytes / (epi_ms * 1e-3) / 1e9,
                                  "peak": hbm, "unit": "GB/s", "frac": epi_bytes / (epi_ms * 1e-3) / 1e9 / hbm,
                                  "algorithmic_bytes_per_launch": epi_bytes, "traffic": None}}


def main():
    args = parse()
    from ocm_vae_simca_b200 import dist as odist
    rank, local_rank, world = odist.env_world()
    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    from ocm_vae_simca_b200 import engine, native, synth
    rank, local_rank, world = odist.init("nccl")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    import torch.distributed as dist
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)

    p, A, rows = args.p, args.A, args.rows
    peak, peak_src = measured_peak()
    pca, lim = fit_bench_model_gpu(p, A, device)
    mean32, P32 = pca.arrays32()
    Sinv = torch.diag(torch.as_tensor(1.0 / pca.eigs_all[:A], device=device)).float()
    ct, cq = lim.reduced_scales("alt")

    X = torch.empty((rows, p), dtype=torch.float32, device=device)
    synth.make_spectra_device(rows, p, device, seed=1000 + rank, out=X)
    torch.cuda.synchronize()

    lib = native.lib()
    if args.force_simt:
        kernel_name = "score_simt_kernel"
    elif lib.ocm_simca_score_tc_supported(rows, p, p, A, 0, C_void(X)):
        bf16a = os.environ.get("OCM_SCORE_BF16A", "0") == "1"
        kernel_name = ("score_tc_kernel<false> (tcgen05.mma kind::f16, 3-term bf16 operands, TMEM A operand, TMA ring)" if bf16a else
                       "score_tc_kernel<true> (tcgen05.mma kind::f16, 2-term fp16 A operand from TMEM prescaled from the model's x_rms, "
                       "3-term fp16 loadings, TMA ring) + score_fixup_kernel (fp32 recomputation of rows outside the fp16 range)")
    elif lib.ocm_simca_score_rowteam_supported(rows, p, p, A, 0, C_void(X)):
        kernel_name = f"score_rowteam_kernel<{A},{p // 128 if p <= 1024 else 8},{128 if p <= 1024 else 256}> (FFMA2)"
    else:
        kernel_name = "score_simt_kernel"

    def step():
        return engine.score(X, mean32, P32, Sinv, rule="alt", ct=ct, cq=cq, dlim=lim.D_limit,
                            force_simt=args.force_simt, x_rms=pca.x_rms)

    args.warmup = max(args.warmup, 3)          # timing rule: at least 3 warm-up steps
    for _ in range(args.warmup):
        out = step()
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    l0 = native.launch_count()
    evs[0].record()
    for i in range(args.steps):
        out = step()
        evs[i + 1].record()
    torch.cuda.synchronize()
    barrier()
    launches = native.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = evs[0].elapsed_time(evs[-1])
    per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    total_ms_max = odist.max_over_ranks(total_ms, device)
    kernel_ms = float(np.mean(per_step))           # a step = the scoring kernel (+ its 30 us fp32 fix-up scan in the fp16 mode): charged to the kernel
    accepted = int(out[2].sum().item())

    value = world * rows * args.steps / (total_ms_max * 1e-3)
    achieved = BYTES_PER_SPECTRUM(p) * rows / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("score_kernel_dram_bytes_per_launch") if (tj.get("rows") == rows and p == 1024) else None
        except Exception:
            traffic = None

    # ---- strong scaling: the same `rows` in TOTAL, split over the ranks (BASELINE configs[2] "rows sharded over 1/2/4/8") ---
    strong = None
    if not args.no_strong:
        lo, hi = odist.shard_rows(rows, rank, world)
        Xs = X[: hi - lo]                              # this rank's shard size; the rows themselves are rank-private already
        sstep = lambda: engine.score(Xs, mean32, P32, Sinv, rule="alt", ct=ct, cq=cq, dlim=lim.D_limit, force_simt=args.force_simt, x_rms=pca.x_rms)
        for _ in range(3):
            sstep()
        torch.cuda.synchronize(); barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            sstep()
        s1.record()
        torch.cuda.synchronize(); barrier()
        sms_ = odist.max_over_ranks(s0.elapsed_time(s1), device)
        strong = {"metric": "spectra scored/s (fused T2+Q)", "scaling": "strong", "rows_total": rows, "rows_per_gpu": hi - lo,
                  "value": rows * args.steps / (sms_ * 1e-3), "unit": "spectra/s", "ms_per_step": sms_ / args.steps,
                  "roofline_frac": BYTES_PER_SPECTRUM(p) * (hi - lo) / (sms_ / args.steps * 1e-3) / 1e9 / peak}

    # ---- e2e: host buffers through the C ABI (H2D + kernel + D2H inside the timed region) -----
    e2e = None
    if not args.no_e2e:
        import ctypes as C
        er = min(args.e2e_rows, rows)
        numa = bind_to_gpu_numa_node(local_rank)       # before the pinned allocations: first touch on the GPU's own node
        hX = torch.empty((er, p), dtype=torch.float32).pin_memory()
        hX.copy_(X[:er])
        hT2 = torch.empty(er, dtype=torch.float32).pin_memory()
        hQ = torch.empty(er, dtype=torch.float32).pin_memory()
        hA = torch.empty(er, dtype=torch.uint8).pin_memory()
        h_mu = mean32.cpu().contiguous(); h_P = P32.cpu().contiguous(); h_S = Sinv.cpu().contiguous()
        prm = native.ScoreParams(A=A, rule=1, snv=0, snv_eps=1e-8, ct=float(ct), cq=float(cq), dlim=float(lim.D_limit), x_rms=float(pca.x_rms))
        vp = lambda t: C.c_void_p(t.data_ptr())
        call = lambda: native.check(native.lib().ocm_simca_score_host_f32(
            vp(hX), er, p, p, vp(h_mu), vp(h_P), vp(h_S), C.byref(prm), 0, vp(hT2), vp(hQ), vp(hA)), "host score")
        call()
        torch.cuda.synchronize(); barrier()
        esteps = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(esteps):
            call()
        torch.cuda.synchronize()
        edt = time.perf_counter() - t0
        edt = odist.max_over_ranks(edt, device)
        # what the link itself gives: one bare pinned H2D copy of the same buffer, all ranks at once
        dX = X[:er]
        torch.cuda.synchronize(); barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            dX.copy_(hX, non_blocking=True)
        torch.cuda.synchronize()
        link = 2 * er * p * 4 / odist.max_over_ranks(time.perf_counter() - t0, device) / 1e9
        e2e = {"value": world * er * esteps / edt, "unit": "spectra/s",
               "h2d_bytes_per_step": int(er * p * 4), "d2h_bytes_per_step": int(er * 9),
               "rows_per_step": er, "steps": esteps,
               "limit": {"what": "PCIe host->device link of each rank (the data plane of this call); the kernel needs 4 105 B per "
                                 "spectrum from HBM at ~4.5 TB/s, the link delivers 4 096 B per spectrum at the rate below",
                         "bare_pinned_h2d_gbs_per_rank": link, "achieved_gbs_per_rank": er * p * 4 * esteps / edt / 1e9,
                         "host_binding": numa},
               "api": "ocm_simca_score_host_f32 (pinned host X -> T2,Q,accept on host; chunked double-buffered H2D)",
               "check": {"accepted_first_chunk": int(hA[: 1 << 10].sum().item())}}
        if world == 1:
            # the same call on a plain numpy matrix (pageable memory: what SIMCA.predict receives from a reference user);
            # reported next to the pinned figure, not instead of it
            try:
                Xnp = hX.numpy().copy()
                hstep = lambda: engine.score_host(Xnp, pca.mean, pca.P, Sinv, rule="alt", ct=float(ct), cq=float(cq),
                                                  dlim=float(lim.D_limit), x_rms=float(pca.x_rms))
                hstep()
                t0 = time.perf_counter()
                for _ in range(2):
                    acc_np = hstep()[2]
                ndt = time.perf_counter() - t0
                e2e["numpy_pageable"] = {"value": 2 * er / ndt, "unit": "spectra/s", "link_gbs": 2 * er * p * 4 / ndt / 1e9,
                                         "api": "engine.score_host (SIMCA.predict on a large numpy matrix) -> ocm_simca_score_host_f32: "
                                                "pageable rows through pinned bounce buffers filled by host threads",
                                         "check": {"accepted_first_chunk": int(acc_np[: 1 << 10].sum())}}
                del Xnp
            except Exception as e:
                e2e["numpy_pageable"] = {"error": repr(e)}
        del hX

    del X
    torch.cuda.empty_cache()
    fit_out = None
    if not args.no_fit:
        try:
            fit_out = fit_leg(args, device, dist.group.WORLD if world > 1 else None, rank, world)
        except Exception as e:
            fit_out = {"error": repr(e)}
    vae_out = None
    if not args.no_vae:
        try:
            vae_out = vae_leg(args, device, dist.group.WORLD if world > 1 else None, rank, world, local_rank)
        except Exception as e:
            vae_out = {"error": repr(e)}
    cv_out = None
    if not args.no_cv:
        try:
            cv_out = cv_leg(device, dist.group.WORLD if world > 1 else None, rank, world)
        except Exception as e:
            cv_out = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {
        "metric": "spectra scored/s (fused T2+Q)", "value": value, "unit": "spectra/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": "one ncu --set full capture of this kernel (profiles/, dram__bytes_read + write per launch); not re-measured in this run",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": BYTES_PER_SPECTRUM(p) * rows,
                     "kernel": kernel_name,
                     "kernel_ms": kernel_ms},
        "clocks": clocks, "gpu_launches": int(launches),
        "check": {"accepted_frac": accepted / rows},
    }
    if e2e is not None:
        line["e2e"] = e2e
    if not args.no_cpu:
        try:
            line["cpu_baseline"] = cpu_baseline_leg(args)
        except Exception as e:
            line["cpu_baseline"] = {"error": repr(e)}
    if strong is not None:
        line["strong"] = strong
    if fit_out is not None:
        line["fit"] = fit_out
    if vae_out is not None:
        line["vae"] = vae_out
    if cv_out is not None:
        line["cv"] = cv_out
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def C_void(t):
    import ctypes as C
    return C.c_void_p(t.data_ptr())


if __name__ == "__main__":
    main()
