"""Launch every hot kernel of the library a few times on representative shapes (for ncu: `-k regex:...` picks one).
usage: python tools/kernels_once.py [score|gram|f64|vae|all]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from ocm_vae_simca_b200 import engine, synth

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda")
if what in ("score", "all"):
    n, p, A = 2_000_000, 1024, 10
    X = synth.make_spectra_device(n, p, dev, seed=1)
    pca = engine.fit_pca(X[:5000].contiguous(), None, A)
    Sinv = torch.diag(torch.as_tensor(1.0 / pca.eigs_all[:A], device=dev))
    for _ in range(3):
        engine.score(X, pca.mean, pca.P, Sinv, rule="alt", ct=0.05, cq=100.0, x_rms=pca.x_rms)     # fp16 operand mode + fix-up
    engine.score(X, pca.mean, pca.P, Sinv, rule="alt", ct=0.05, cq=100.0)                           # bf16 x 3 mode
    torch.cuda.synchronize()
    if what == "score":
        # 2048 wavelengths: the cluster kernel, four CTAs per tile
        X2 = synth.make_spectra_device(1_000_000, 2048, dev, seed=2)
        pc2 = engine.fit_pca(X2[:5000].contiguous(), None, A)
        S2 = torch.diag(torch.as_tensor(1.0 / pc2.eigs_all[:A], device=dev))
        for _ in range(3):
            engine.score(X2, pc2.mean, pc2.P, S2, rule="alt", ct=0.05, cq=100.0)
        torch.cuda.synchronize()
        del X2
    if what == "all":
        # generic kernels: multi-A (CV grid shape), fused SNV, fp64 rows
        Xs = X[:200_000, :1000].contiguous()
        pc2 = engine.fit_pca(Xs[:3000].contiguous(), None, 20)
        inv = torch.as_tensor(1.0 / pc2.eigs_all[:20], device=dev)
        for _ in range(2):
            engine.score_multiA(Xs, pc2.mean, pc2.P, inv)
            engine.score(Xs, pc2.mean, pc2.P[:10].contiguous(), torch.diag(inv[:10]), snv=True)
        # preprocessing
        for _ in range(2):
            engine.preprocess(X[:1_000_000], "snv")
            engine.savgol(X[:1_000_000], 5, 2, 1, snv=True)
        # VAE epilogue
        z = torch.randn(1_000_000, 16, device=dev)
        xr = X[:1_000_000].to(torch.bfloat16)
        for _ in range(2):
            engine.latent_score(z=z, m=torch.zeros(16, device=dev), Sinv=torch.eye(16, device=dev), x=X[:1_000_000], xrec=xr)
        # small Jacobi
        H = torch.randn(10, 32, 32, dtype=torch.float64, device=dev)
        engine.eigh_small(H @ H.transpose(1, 2))
        torch.cuda.synchronize()
    del X
if what in ("gram", "all"):
    n, p = 196_224, 2048
    X = synth.make_spectra_device(n, p, dev, seed=5)
    s, mn, mx, cnt = engine.colstats_minmax(X)
    for _ in range(2):
        engine.gram_tc(X, None, s / cnt, mn, mx, products=13)
    engine.gram(X[:50_000], None, s / cnt)
    torch.cuda.synchronize()
if what == "f64":
    # fp64 rows on the staged bulk path: fp64 centring -> fp32 scratch -> tcgen05 kernel
    n, p, A = 500_000, 1024, 10
    X = synth.make_spectra_device(n, p, dev, seed=7).double()
    pca = engine.fit_pca(X[:5000].contiguous(), None, A)
    Sinv = torch.diag(torch.as_tensor(1.0 / pca.eigs_all[:A], device=dev))
    for snv in (False, False, True):
        engine.score(X, pca.mean, pca.P, Sinv, rule="alt", ct=0.05, cq=100.0, x_rms=pca.x_rms, snv=snv)
    torch.cuda.synchronize()
if what == "vae":
    # config-5 scoring pass through vae.FrozenScorer: first block, bias + activation, decoder tail (csrc/vae_layers.cu)
    from ocm_vae_simca_b200.vae import ConvVAE1D, FrozenScorer
    X = synth.make_spectra_device(16384, 1024, dev, seed=9)
    torch.manual_seed(0)
    vae = ConvVAE1D(1024, 16, X.mean(0).cpu().numpy(), (X.std(0) + 1e-8).cpu().numpy()).to(dev).eval()
    sc = FrozenScorer(vae)
    for r in (0, 8192, 0):
        sc(X[r:r + 8192])
    torch.cuda.synchronize()
print("ok")
