"""Where does the batched CV grid spend its time? (GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sklearn.model_selection import PredefinedSplit
from ocm_vae_simca_b200 import engine, limits, native, synth
from ocm_vae_simca_b200.utils import SIMCA, cross_validate_simca_grid
X, y = synth.make_class_dataset(1500, 500, 1050, seed=1234)
cv = PredefinedSplit(np.arange(X.shape[0]) % 10)
est = SIMCA(model_class=0, verbose=False)
run = lambda: cross_validate_simca_grid(est, X, y, cv, LV_min=1, LV_max=20, class_index=0, print_summary=False)
run(); torch.cuda.synchronize()
t0 = time.perf_counter(); run(); torch.cuda.synchronize(); print("total", time.perf_counter() - t0)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); run(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
Xd = engine.to_device(X)
idx = engine.index_tensor(np.flatnonzero(y == 0)[:1350], Xd.device)
def T(f, n=5):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("to_device ms", T(lambda: engine.to_device(X)))
print("covariance ms", T(lambda: engine.covariance(Xd, idx)))
Cm, mean, n = engine.covariance(Xd, idx)
print("eigh ms", T(lambda: engine.eigh_topk(Cm, 20)))
Cb = torch.stack([Cm] * 10)
print("eigh batched(10) ms", T(lambda: torch.linalg.eigh(Cb)))
print("eigvalsh ms", T(lambda: torch.linalg.eigvalsh(Cm)))
