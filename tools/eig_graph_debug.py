import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from ocm_vae_simca_b200 import engine, synth
X, y = synth.make_class_dataset(1500, 500, 1050, seed=1234)
Xd = engine.to_device(X)
rows = np.flatnonzero(y == 0)
sets = [engine.index_tensor(rows[np.arange(rows.size) % 10 != f], Xd.device) for f in range(10)]
covs, means, ns, pool = engine._fold_covariances(Xd, sets)
Cs = torch.stack(covs + [pool["cov"]])
for use in (True, False, True, True, False):
    engine.EIG_GRAPH = use
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r0 = engine.graph_replayed_launches
    w, V, tails, conv = engine.eig_topk_subspace(Cs, 20)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("graph" if use else "eager", f"{dt*1e3:.2f} ms", "conv", conv.cpu().numpy().astype(int), "replayed", engine.graph_replayed_launches - r0,
          "w20", float(w[0, 19]), "graph object", None if not use else type(engine._EIG_GRAPHS.get((0, 11, 1050, 32))).__name__)
