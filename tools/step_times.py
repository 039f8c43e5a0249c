"""Per-launch device times of the headline scoring step, back to back (does a power cap bite over a long run?).
    python tools/step_times.py [--steps 60] [--rows 10000000]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ocm_vae_simca_b200 import engine, limits, synth

ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=60); ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--p", type=int, default=1024); ap.add_argument("--gap-ms", type=float, default=0.0)
ap.add_argument("--snv", type=int, default=0); ap.add_argument("--f16", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
pca = engine.fit_pca(synth.make_spectra_device(5000, a.p, dev, seed=99), None, 10)
lim = limits.all_limits("alt", "Fdist", 0.95, "jm", 0.95, 0.95, 10, pca.n_samples, pca.eigs_all)
Sinv = torch.diag(torch.as_tensor(1.0 / pca.eigs_all[:10], device=dev))
X = synth.make_spectra_device(a.rows, a.p, dev, seed=123)
kw = dict(rule="alt", ct=1.0 / lim.T2_limit, cq=1.0 / lim.Q_limit, dlim=lim.D_limit)
run = lambda: engine.score(X, pca.mean, pca.P, Sinv, snv=bool(a.snv), x_rms=pca.x_rms if a.f16 else 0.0, **kw)
for _ in range(3): run()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
ev[0].record()
for i in range(a.steps):
    run(); ev[i + 1].record()
    if a.gap_ms: torch.cuda._sleep(int(a.gap_ms * 1.9e6))
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.steps)]
print(json.dumps({"first5": [round(x, 3) for x in ms[:5]], "last5": [round(x, 3) for x in ms[-5:]],
                  "min": round(min(ms), 3), "median": round(sorted(ms)[len(ms) // 2], 3), "max": round(max(ms), 3)}))
