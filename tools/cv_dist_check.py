"""CV grid with the folds dealt over the ranks: every rank must return the single-process table.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/cv_dist_check.py
OCM_DIST_BACKEND=gloo lets several ranks share one GPU (tests); the default is nccl, one GPU per rank."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from sklearn.model_selection import PredefinedSplit
from ocm_vae_simca_b200 import synth
from ocm_vae_simca_b200.utils import SIMCA, ClasswiseKFoldWithExternalVal, cross_validate_simca_grid

backend = os.environ.get("OCM_DIST_BACKEND", "nccl")
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local % torch.cuda.device_count())
dist.init_process_group(backend, rank=rank, world_size=world)
n_t, n_o, p = (1500, 500, 1050) if os.environ.get("OCM_CV_FULL", "1") == "1" else (270, 90, 160)
X, y = synth.make_class_dataset(n_t, n_o, p, seed=1234)
cv = PredefinedSplit(np.arange(X.shape[0]) % 10)
est = SIMCA(model_class=0, verbose=False)
kw = dict(LV_min=1, LV_max=20, class_index=0, print_summary=False, param_grid={"type": ["alt", "dd"]})
single = cross_validate_simca_grid(est, X, y, cv, **kw)
run = lambda: cross_validate_simca_grid(est, X, y, cv, group=dist.group.WORLD, **kw)
multi = run()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter(); run(); torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
key = lambda r: (sorted(r["params"].items()), r["LV"])
a = sorted(single["results"], key=key); b = sorted(multi["results"], key=key)
same = len(a) == len(b) and all(x["LV"] == z["LV"] and x["params"] == z["params"] and
                                 np.allclose([x["spec"], x["sens"], x["eff"]], [z["spec"], z["sens"], z["eff"]], rtol=0, atol=1e-12, equal_nan=True)
                                 for x, z in zip(a, b))
# a splitter whose test sets OVERLAP (every non-target row is tested in every fold): the pooled per-row decisions must
# be those of the last fold that tested the row, as in the single-process run and in the reference
cvk = ClasswiseKFoldWithExternalVal(n_splits=5, cls_label=0)
kw2 = dict(LV_min=2, LV_max=6, class_index=0, print_summary=False, store_predictions=True)
s2 = cross_validate_simca_grid(est, X, y, cvk, **kw2)
m2 = cross_validate_simca_grid(est, X, y, cvk, group=dist.group.WORLD, **kw2)
same_pooled = all(np.array_equal(u["prediction"], v["prediction"]) for u, v in zip(s2["by_combo"], m2["by_combo"]))
same_pooled = same_pooled and all(np.allclose([u["spec"], u["sens"]], [v["spec"], v["sens"]], atol=1e-12, equal_nan=True)
                                  for u, v in zip(s2["results"], m2["results"]))
flags = [None] * world
dist.all_gather_object(flags, bool(same and same_pooled and single["best_LV"] == multi["best_LV"] and single["best_params"] == multi["best_params"]))
if rank == 0:
    print(json.dumps({"world": world, "backend": backend, "identical_on_all_ranks": all(flags), "fits": 2 * 10 * 20 + 1,
                      "seconds": dt, "fits_per_s": (2 * 10 * 20 + 1) / dt, "best_LV": int(multi["best_LV"])}))
dist.destroy_process_group()
sys.exit(0 if all(flags) else 1)
