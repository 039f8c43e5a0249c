"""fp64 rows: bulk path (fp64 centring -> fp32 scratch -> tcgen05 kernel) against the generic kernel's fp64 path.
usage: python tools/f64_bench.py [rows] [p]"""
import ctypes as C
import json
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from ocm_vae_simca_b200 import engine, native, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
A = 10
dev = torch.device("cuda")
X = synth.make_spectra_device(n, p, dev, seed=1).double()
pca = engine.fit_pca(X[:5000].contiguous(), None, A)
Sinv = torch.diag(torch.as_tensor(1.0 / pca.eigs_all[:A], device=dev))
lib = native.lib()
vp = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
mu64 = pca.mean.contiguous(); _, P32 = pca.arrays32(); S32 = Sinv.float().contiguous()
T2 = torch.empty(n, device=dev); Q = torch.empty(n, device=dev); acc = torch.empty(n, dtype=torch.uint8, device=dev)
out = {"rows": n, "p": p, "A": A}
for snv in (0, 1):
    prm = native.ScoreParams(A=A, rule=1, snv=snv, snv_eps=1e-8, ct=0.05, cq=100.0, dlim=float(np.sqrt(2)), x_rms=pca.x_rms)
    for name, fn in (("bulk", lib.ocm_simca_score_f64in), ("generic", lib.ocm_simca_score_simt_f64in)):
        ts = []
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            native.check(fn(vp(X), n, p, p, vp(mu64), vp(P32), vp(S32), C.byref(prm), vp(T2), vp(Q), vp(acc), None, st))
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts[1:]))
        out[f"{name}_snv{snv}_ms"] = round(ms, 3)
        out[f"{name}_snv{snv}_GBs_of_8p"] = round(8.0 * p * n / ms / 1e6, 1)
del X
# fp32 rows, fused SNV, 2048 wavelengths: staged (SNV + centring kernel -> scratch -> cluster kernel) against the generic kernel
n2, p2 = n // 2, 2048
X = synth.make_spectra_device(n2, p2, dev, seed=2)
pca = engine.fit_pca(engine.preprocess(X[:5000].contiguous(), "snv"), None, A)
Sinv = torch.diag(torch.as_tensor(1.0 / pca.eigs_all[:A], device=dev))
mu32, P32 = pca.arrays32(); S32 = Sinv.float().contiguous()
prm = native.ScoreParams(A=A, rule=1, snv=1, snv_eps=1e-8, ct=0.05, cq=100.0, dlim=float(np.sqrt(2)), x_rms=pca.x_rms)
for name, fn in (("staged", lib.ocm_simca_score_f32), ("generic", lib.ocm_simca_score_simt_f32)):
    ts = []
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        native.check(fn(vp(X), n2, p2, p2, vp(mu32), vp(P32), vp(S32), C.byref(prm), vp(T2), vp(Q), vp(acc), None, st))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts[1:]))
    out[f"f32_snv_p2048_{name}_ms"] = round(ms, 3)
    out[f"f32_snv_p2048_{name}_GBs_of_4p"] = round(4.0 * p2 * n2 / ms / 1e6, 1)
print(json.dumps(out))
