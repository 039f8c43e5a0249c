"""FrozenScorer (fused path): batch size and cudnn.benchmark."""
import json, sys
import torch
sys.path.insert(0, ".")
from ocm_vae_simca_b200 import synth
from ocm_vae_simca_b200.vae import ConvVAE1D, FrozenScorer
dev = torch.device("cuda")
L, d, n = 1024, 16, 262144
X = synth.make_spectra_device(n, L, dev, seed=3000)
torch.manual_seed(0)
vae = ConvVAE1D(L, d, X[:65536].mean(0).cpu().numpy(), (X[:65536].std(0) + 1e-8).cpu().numpy()).to(dev).eval()
sc = FrozenScorer(vae)
out = {}
def timed(Bs, reps=2):
    def f():
        for r in range(0, n, Bs):
            sc(X[r:r + Bs])
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record(); torch.cuda.synchronize()
    return n * reps / (e0.elapsed_time(e1) * 1e-3)
for Bs in (4096, 8192, 16384, 32768):
    out[f"fused B={Bs}"] = timed(Bs)
torch.backends.cudnn.benchmark = True
for Bs in (8192, 16384):
    out[f"fused B={Bs} cudnn.benchmark"] = timed(Bs)
print(json.dumps(out))
