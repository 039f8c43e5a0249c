"""Config-4 style measurement: row-sharded PCA fit (Gram on the FP64 tensor pipe + NCCL all-reduce of the
moments + eigensolve), one process per GPU (torchrun).  Prints one JSON line on rank 0.

    python -m torch.distributed.run --nproc-per-node N tools/bench_fit_dist.py --rows-per-gpu 1000000 --p 2048
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from ocm_vae_simca_b200 import dist as odist, engine, native, synth

ap = argparse.ArgumentParser()
ap.add_argument("--rows-per-gpu", type=int, default=1_000_000)
ap.add_argument("--p", type=int, default=2048)
ap.add_argument("--A", type=int, default=20)
args = ap.parse_args()
rank, local, world = odist.init("nccl")
dev = torch.device("cuda", local)
n, p = args.rows_per_gpu, args.p
X = synth.make_spectra_device(n, p, dev, seed=100 + rank)
group = dist.group.WORLD if world > 1 else None
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
# warm-up (NCCL communicator, kernels)
Cw, _, _ = engine.covariance(X[:4096].contiguous(), None, group)
engine.eigh_topk(Cw, args.A)            # cuSOLVER handle / workspace
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
ev[0].record()
s0, _, cnt = engine.colstats(X)
ntot = torch.tensor([float(cnt)], dtype=torch.float64, device=dev)
if group is not None:
    dist.all_reduce(s0, group=group); dist.all_reduce(ntot, group=group)
shift = s0 / ntot
ev[1].record()
G = engine.gram(X, None, shift)
s1, _, _ = engine.colstats(X, None, shift)
ev[2].record()
if group is not None:
    dist.all_reduce(G, group=group); dist.all_reduce(s1, group=group)
ev[3].record()
N = int(round(float(ntot.item())))
Cm, mean = engine.cov_from_parts(G.unsqueeze(0), s1.unsqueeze(0), [N], [1.0], shift)
evals, evecs = engine.eigh_topk(Cm, args.A)
ev[4].record()
torch.cuda.synchronize()
t = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
tmax = [odist.max_over_ranks(v, dev) for v in t]
if rank == 0:
    flops = float(world) * n * p * (p + 1)
    total = sum(tmax)
    print(json.dumps({
        "metric": "PCA fit (partitioned Gram + all-reduce + eigensolve)", "n_gpus": world, "rows_total": world * n, "p": p,
        "ms": {"colsum+allreduce": tmax[0], "gram(fp64 DMMA)+colsum": tmax[1], "allreduce(G,s)": tmax[2], "cov+eigh": tmax[3], "total": total},
        "gram_tflops_algorithmic_all_gpus": flops / (tmax[1] * 1e-3) / 1e12,
        "rows_per_s": world * n / (total * 1e-3), "allreduce_payload_MB": (p * p + p) * 8 / 1e6,
        "top_eigs": [float(v) for v in evals[:3]]}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
