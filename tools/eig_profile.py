import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ocm_vae_simca_b200 import engine, synth
dev = torch.device("cuda")
X, y = synth.make_class_dataset(1500, 500, 1050, seed=1234)
Xd = engine.to_device(X)
rows = np.flatnonzero(y == 0)
Cs = torch.stack([engine.covariance(Xd, engine.index_tensor(rows[np.arange(len(rows)) % 10 != f], dev))[0] for f in range(10)])
def T(f, n=5):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
Q = torch.linalg.qr(torch.randn(10, 1050, 32, dtype=torch.float64, device=dev))[0]
print("bmm C@Q ms", T(lambda: Cs @ Q))
Z = Cs @ Q
print("qr ms", T(lambda: torch.linalg.qr(Z)))
H = Q.transpose(1, 2) @ Z
print("small bmm ms", T(lambda: Q.transpose(1, 2) @ Z))
print("jacobi ms", T(lambda: engine.eigh_small(H)))
print("torch eigh small ms", T(lambda: torch.linalg.eigh(H)))
print("R@R ms", T(lambda: Cs @ Cs))
print("subspace total ms", T(lambda: engine.eig_topk_subspace(Cs, 20), 3))
w, V, t, conv = engine.eig_topk_subspace(Cs, 20)
print("converged", conv.tolist())
import cProfile, pstats
