"""Multi-GPU check (torchrun, NCCL): row-sharded PCA fit == single-GPU fit; sharded scoring == whole."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from ocm_vae_simca_b200 import dist as odist, engine, synth
rank, local, world = odist.init("nccl")
dev = torch.device("cuda", local)
n, p, A = 60_001, 1024, 10
Xall = synth.make_spectra(n, p, seed=5).astype(np.float32)      # same on every rank (seeded)
lo, hi = odist.shard_rows(n, rank, world)
Xs = torch.as_tensor(Xall[lo:hi], device=dev)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
Cm, mean, N = engine.covariance(Xs, None, group=dist.group.WORLD)
e1.record(); torch.cuda.synchronize()
evals, evecs = engine.eigh_topk(Cm, A)
if rank == 0:
    Cf, meanf, Nf = engine.covariance(torch.as_tensor(Xall, device=dev))
    ef, vf = engine.eigh_topk(Cf, A)
    print("world", world, "N", N, Nf, "cov max abs diff %.3e (scale %.3e)" % ((Cm - Cf).abs().max().item(), Cf.abs().max().item()),
          "eig rel diff %.3e" % ((evals[:A] - ef[:A]).abs() / ef[:A]).max().item(),
          "loadings diff %.3e" % (evecs - vf).abs().max().item(), "sharded covariance+allreduce ms %.3f" % e0.elapsed_time(e1))
dist.barrier()
dist.destroy_process_group()
