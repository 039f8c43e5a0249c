"""Gram accumulation alone (fit path): TFLOP/s of n*p*(p+1) algorithmic flops.  python tools/gram_bench.py [--rows N --p P]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ocm_vae_simca_b200 import engine, synth
ap = argparse.ArgumentParser(); ap.add_argument("--rows", type=int, default=500_000); ap.add_argument("--p", type=int, default=2048)
ap.add_argument("--reps", type=int, default=3); a = ap.parse_args()
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
X = synth.make_spectra_device(a.rows, a.p, dev, seed=2)
shift = X[:4096].double().mean(0)
G = torch.zeros((a.p, a.p), dtype=torch.float64, device=dev)
f = lambda: engine.gram(X, None, shift, out=G)
f(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps): f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
G.zero_(); f()
ref = (X[:20000].double() - shift).T @ (X[:20000].double() - shift)
G2 = torch.zeros_like(G); engine.gram(X[:20000].contiguous(), None, shift, out=G2)
err = float(((torch.triu(G2) - torch.triu(ref)).abs().max() / ref.abs().max()).item())
print(json.dumps({"rows": a.rows, "p": a.p, "ms": ms, "tflops_algorithmic": a.rows * a.p * (a.p + 1) / ms / 1e9, "max_rel_err_20k_rows": err}))
