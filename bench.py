#!/usr/bin/env python
"""Benchmark of the SIMCA hot path on B200 (contract: see the task's bench.py section).

Headline workload = BASELINE.json configs[2]: bulk scoring of 10 M x 1,024 fp32 synthetic
spectra against a fixed 10-PC SIMCA model, fused T2+Q+decision, rows sharded over the GPUs
(weak scaling: every rank scores its own 10 M-row shard, no data-path collective).
A "step" is one pass of the fused kernel over the rank's shard (41 GB >> 126 MB L2, so
every step streams from HBM; no explicit L2 flush is needed).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
  python bench.py --impl reference ...                           # CPU oracle arm (reference algorithm)

Extra keys beyond the base contract: `roofline`, `cpu_baseline`, `e2e`, `clocks`,
`gpu_launches`, and `cv` (BASELINE.json's second metric: CVSIMCA fits/s on configs[1]).
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


def reference_arm(args, rank):
    """CPU arm: the oracle's restatement of SIMCA.predict (numpy/sklearn, all host threads) on a
    bounded sample of the same workload.  Rank 0 only."""
    if rank != 0:
        return
    from ocm_vae_simca_b200 import synth
    from oracle import simca_oracle as so
    p, A = args.p, args.A
    rows = args.cpu_rows
    Xfit = synth.make_spectra(5000, p, seed=99).astype(np.float32)
    model = so.fit_one_class(Xfit.astype(np.float64), A, faithful_solver=True)
    X = synth.make_spectra(rows, p, seed=123).astype(np.float32)
    pool, cores = _all_cores()
    with pool:
        for _ in range(max(args.warmup, 1)):
            so.predict(model, X[: min(rows, 20000)])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            so.predict(model, X)
        dt = time.perf_counter() - t0
    val = rows * args.steps / dt
    line = {
        "impl": "reference", "metric": "spectra scored/s (fused T2+Q)", "value": val, "unit": "spectra/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": val, "unit": "spectra/s", "cores": cores, "kind": "port",
                         "sample": f"{rows} x {p} fp32 rows per step, oracle.simca_oracle.predict (numpy/OpenBLAS, {cores} threads)"},
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
    from oracle import simca_oracle as so
    p, A, rows = args.p, args.A, args.cpu_rows
    Xfit = synth.make_spectra(5000, p, seed=99)
    model = so.fit_one_class(Xfit, A, faithful_solver=True)
    X = synth.make_spectra(rows, p, seed=123).astype(np.float32)
    pool, cores = _all_cores()
    with pool:
        so.predict(model, X[:20000])
        best = float("inf")
        for _ in range(3):
            t0 = time.perf_counter()
            so.predict(model, X)
            best = min(best, time.perf_counter() - t0)
    return {"value": rows / best, "unit": "spectra/s", "cores": cores, "kind": "port",
            "sample": f"{rows} x {p} fp32 rows, best of 3, oracle.simca_oracle.predict (numpy/OpenBLAS, {cores} threads)"}


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
    l0 = native.launch_count()
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
           "best_LV": int(res["best_LV"]), "gpu_launches": native.launch_count() - l0,
           "timing": "host wall clock around the public call (includes H2D of X and host limit arithmetic)"}
    # CPU: the oracle's per-(fold, A) fit + predict, bounded to 3 cells, extrapolated per fit
    try:
        from oracle import simca_oracle as so
        tr, te = next(iter(cv.split()))
        Xtr = X[tr][y[tr] == 0]
        pool, cores = _all_cores()
        with pool:
            t0 = time.perf_counter()
            cells = 0
            for A in (5, 10, 15):
                m = so.fit_one_class(Xtr, A, faithful_solver=True)
                so.predict(m, X[te])
                cells += 1
            cdt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": cells / cdt, "unit": "fits/s", "cores": cores, "kind": "port",
                               "sample": "3 (fold, A) cells of the same grid: oracle fit_one_class + predict"}
    except Exception as e:  # the oracle is a checker; its absence must not kill the GPU line
        out["cpu_baseline"] = {"error": repr(e)}
    return out


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
        kernel_name = "score_tc_kernel<false> (tcgen05.mma kind::f16 bf16x3, TMEM operands, TMA ring)"
    elif lib.ocm_simca_score_rowteam_supported(rows, p, p, A, 0, C_void(X)):
        kernel_name = f"score_rowteam_kernel<{A},{p // 128 if p <= 1024 else 8},{128 if p <= 1024 else 256}> (FFMA2)"
    else:
        kernel_name = "score_simt_kernel"

    def step():
        return engine.score(X, mean32, P32, Sinv, rule="alt", ct=ct, cq=cq, dlim=lim.D_limit,
                            force_simt=args.force_simt)

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
    kernel_ms = float(np.mean(per_step))           # one launch per step: step time == kernel time
    accepted = int(out[2].sum().item())

    value = world * rows * args.steps / (total_ms_max * 1e-3)
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_SPECTRUM(p) * rows / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("score_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None

    # ---- e2e: host buffers through the C ABI (H2D + kernel + D2H inside the timed region) -----
    e2e = None
    if not args.no_e2e:
        import ctypes as C
        er = min(args.e2e_rows, rows)
        hX = torch.empty((er, p), dtype=torch.float32).pin_memory()
        hX.copy_(X[:er])
        hT2 = torch.empty(er, dtype=torch.float32).pin_memory()
        hQ = torch.empty(er, dtype=torch.float32).pin_memory()
        hA = torch.empty(er, dtype=torch.uint8).pin_memory()
        h_mu = mean32.cpu().contiguous(); h_P = P32.cpu().contiguous(); h_S = Sinv.cpu().contiguous()
        prm = native.ScoreParams(A=A, rule=1, snv=0, snv_eps=1e-8, ct=float(ct), cq=float(cq), dlim=float(lim.D_limit))
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
        e2e = {"value": world * er * esteps / edt, "unit": "spectra/s",
               "h2d_bytes_per_step": int(er * p * 4), "d2h_bytes_per_step": int(er * 9),
               "rows_per_step": er, "steps": esteps,
               "api": "ocm_simca_score_host_f32 (pinned host X -> T2,Q,accept on host; chunked double-buffered H2D)",
               "check": {"accepted_first_chunk": int(hA[: 1 << 10].sum().item())}}
        del hX

    cv_out = None
    if not args.no_cv:
        del X
        torch.cuda.empty_cache()
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
                     "traffic": traffic, "peak_source": peak_src,
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
