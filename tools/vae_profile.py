"""Kernel-level time split of one FrozenScorer batch and one channels-last training step (torch.profiler)."""
import sys
import torch
sys.path.insert(0, ".")
from ocm_vae_simca_b200 import synth
from ocm_vae_simca_b200.vae import ConvVAE1D, FrozenScorer
from ocm_vae_simca_b200.vae.convvae import beta_vae_cosine_loss
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda")
L, d = 1024, 16
X = synth.make_spectra_device(32768, L, dev, seed=3000)
vae = ConvVAE1D(L, d, X.mean(0).cpu().numpy(), (X.std(0) + 1e-8).cpu().numpy()).to(dev)
vae.channels_last = True
sc = FrozenScorer(vae.eval())
for _ in range(2):
    sc(X[:8192])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    for r in range(0, 32768, 8192):
        sc(X[r:r + 8192])
    torch.cuda.synchronize()
print("=== scoring, 4 batches of 8192 ===")
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=90))
vae.train()
opt = torch.optim.Adam(vae.parameters(), lr=1e-4, fused=True)
def step(i):
    xb = X[i * 4096:(i + 1) * 4096]
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        xr, mu_, lv_ = vae(xb)
        loss = beta_vae_cosine_loss(xb, xr.float(), mu_.float(), lv_.float(), as_floats=False)[0]
    loss.backward(); opt.step()
for i in range(2):
    step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(2, 4):
        step(i)
    torch.cuda.synchronize()
print("=== training, 2 steps of 4096 ===")
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=90))
