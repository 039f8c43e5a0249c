python - <<'PY'
import ctypes as C, numpy as np, torch, sys
sys.path.insert(0, ".")
from ocm_vae_simca_b200 import engine, native, synth
from oracle import simca_oracle as so
dev = torch.device("cuda"); lib = native.lib()
vp = lambda t: C.c_void_p(t.data_ptr())
for seed in (1, 2, 3):
    n, A, p = 40_000, 10, 1024
    g = torch.Generator(device=dev); g.manual_seed(seed)
    Xraw = synth.make_spectra_device(n, p, dev, seed=seed)
    Xraw = Xraw * (1.0 + 0.3 * torch.rand(n, 1, device=dev, generator=g)) + 0.5 * torch.rand(n, 1, device=dev, generator=g)
    X64 = Xraw.double(); Xs = (X64 - X64.mean(1, keepdim=True)) / (X64.std(1, unbiased=False, keepdim=True) + 1e-8)
    pca = engine.fit_pca(Xs[:3000].contiguous(), None, A)
    Sinv = torch.diag(torch.as_tensor(1.0 / pca.eigs_all[:A], device=dev))
    model = dict(pred_mean=pca.mean.cpu().numpy(), pred_P=pca.P.cpu().numpy(), Sinv=Sinv.cpu().numpy())
    T2r, Qr = so.score(model, Xs.cpu().numpy())
    for x_rms in (0.0, pca.x_rms):
        T2, Q, acc = engine.score(Xraw, pca.mean, pca.P, Sinv, snv=True, x_rms=x_rms)
        print(seed, x_rms > 0, "max rel T2 %.2e Q %.2e" % (np.max(np.abs(T2.cpu().numpy() - T2r) / T2r), np.max(np.abs(Q.cpu().numpy() - Qr) / Qr)))
PY
timeout 600 python -m pytest tests -x -q -m gpu -k "snv or tensor_core" 2>&1 | tail -3
