"""One short integer-Gram run for ncu (tools/gram_tc_prof.py [rows] [p] [products])."""
import sys
import torch
sys.path.insert(0, ".")
from ocm_vae_simca_b200 import engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 196224
p = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
pr = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda")
X = torch.empty((n, p), dtype=torch.float32, device=dev)
synth.make_spectra_device(n, p, dev, seed=5, out=X)
s, mn, mx, cnt = engine.colstats_minmax(X)
shift = s / cnt
for _ in range(2):
    G, cs, flags = engine.gram_tc(X, None, shift, mn, mx, products=pr)
torch.cuda.synchronize()
print("ok", flags.tolist(), float(G[0, 0]))
