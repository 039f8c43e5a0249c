"""Row-sharded PCA fit == single-process fit: partial {Gram, sums, count} all-reduced, both Gram paths.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dist_fit_check.py
OCM_DIST_BACKEND=gloo lets several ranks share one GPU (the GPU test-suite does that); default nccl, one GPU per rank."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from ocm_vae_simca_b200 import dist as odist, engine, synth

backend = os.environ.get("OCM_DIST_BACKEND", "nccl")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local % torch.cuda.device_count())
dist.init_process_group(backend, rank=rank, world_size=world)
dev = torch.device("cuda", torch.cuda.current_device())
n, p, A = 60_001, 384, 10
Xall = synth.make_spectra(n, p, seed=5).astype(np.float32)      # same on every rank (seeded)
lo, hi = odist.shard_rows(n, rank, world)
Xs = torch.as_tensor(Xall[lo:hi], device=dev)
out = {"world": world, "backend": backend}
ok = True
for path in ("dmma", "tc"):
    pca = engine.fit_pca(Xs, None, A, group=dist.group.WORLD, gram_path=path)
    full = engine.fit_pca(torch.as_tensor(Xall, device=dev), None, A, gram_path=path)
    e_l = float(np.max(np.abs(pca.eigs_all[:A] - full.eigs_all[:A]) / full.eigs_all[:A]))
    e_p = float((pca.P - full.P).abs().max())
    e_m = float((pca.mean - full.mean).abs().max())
    out[path] = {"n": pca.n_samples, "eig_rel": e_l, "loadings_abs": e_p, "mean_abs": e_m}
    # the integer Gram quantises each shard on its own column ranges: 2^-30 of the range per element
    tol_l, tol_m = (1e-9, 1e-12) if path == "dmma" else (1e-7, 1e-9)
    ok = ok and pca.n_samples == n and e_l < tol_l and e_p < 1e-6 and e_m < tol_m
flags = [None] * world
dist.all_gather_object(flags, ok)
out["identical_on_all_ranks"] = all(flags)
if rank == 0:
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if all(flags) else 1)
