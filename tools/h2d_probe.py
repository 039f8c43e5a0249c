"""How fast can this box move pinned host spectra to the GPU, and how close does the host-buffer scoring call get?
    python tools/h2d_probe.py [--rows 1000000]"""
import argparse, ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ocm_vae_simca_b200 import engine, native, synth

ap = argparse.ArgumentParser(); ap.add_argument("--rows", type=int, default=1_000_000); a = ap.parse_args()
dev = torch.device("cuda", 0); n, p, A = a.rows, 1024, 10
hX = torch.empty((n, p), dtype=torch.float32).pin_memory()
hX.copy_(synth.make_spectra_device(n, p, dev, seed=5).cpu())
dX = torch.empty((n, p), dtype=torch.float32, device=dev)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
s = t(lambda: dX.copy_(hX, non_blocking=True))
out = {"h2d_1d_GBs": n * p * 4 / s / 1e9}
pca = engine.fit_pca(dX[:5000].contiguous(), None, A)
mu, P = [x.cpu().numpy().astype(np.float32) for x in pca.arrays32()]
S = np.diag(1.0 / pca.eigs_all[:A]).astype(np.float32)
prm = native.ScoreParams(A=A, rule=1, snv=0, snv_eps=1e-8, ct=0.05, cq=50.0, dlim=float(np.sqrt(2)))
T2 = torch.empty(n).pin_memory(); Q = torch.empty(n).pin_memory(); acc = torch.empty(n, dtype=torch.uint8).pin_memory()
vp = lambda x: C.c_void_p(x.ctypes.data if isinstance(x, np.ndarray) else x.data_ptr())
lib = native.lib()
for chunk in (1 << 15, 1 << 16, 1 << 17, 1 << 18):
    s = t(lambda: native.check(lib.ocm_simca_score_host_f32(vp(hX), n, p, p, vp(mu), vp(P), vp(S), C.byref(prm), chunk, vp(T2), vp(Q), vp(acc))), reps=3)
    out[f"host_api_chunk{chunk}_Mspectra_s"] = n / s / 1e6
    out[f"host_api_chunk{chunk}_GBs"] = n * p * 4 / s / 1e9
print(json.dumps(out))
