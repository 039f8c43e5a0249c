"""Diagnostic: tcgen05 scoring kernel vs the generic kernel and the fp64 arithmetic (GPU box)."""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ocm_vae_simca_b200 import engine, native, synth
dev = torch.device("cuda"); lib = native.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128 * 148 * 2 + 5
A = int(sys.argv[2]) if len(sys.argv) > 2 else 10
p = 1024
X = synth.make_spectra_device(n, p, dev, seed=3)
pca = engine.fit_pca(X[:3000].contiguous(), None, A)
Sinv = torch.diag(torch.as_tensor(1.0 / pca.eigs_all[:A], device=dev)).float().contiguous()
mu32, P32 = pca.arrays32()
prm = native.ScoreParams(A=A, rule=1, snv=0, snv_eps=1e-8, ct=0.05, cq=50.0, dlim=float(np.sqrt(2)))
vp = lambda t: C.c_void_p(t.data_ptr()); st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(fn, extra=()):
    T2 = torch.full((n,), -1.0, device=dev); Q = torch.full((n,), -1.0, device=dev); acc = torch.full((n,), 7, dtype=torch.uint8, device=dev)
    native.check(fn(vp(X), n, p, p, vp(mu32), vp(P32), vp(Sinv), C.byref(prm), vp(T2), vp(Q), vp(acc), *extra, st))
    torch.cuda.synchronize(); return T2, Q, acc
dbg = torch.zeros((128, 1024), device=dev)
lib.ocm_simca_score_tc_debug(vp(dbg))
Xd = X.double(); xc = Xd - pca.mean; t = xc @ pca.P.T; r = xc - t @ pca.P
fro = lambda m: m.pow(2).sum().sqrt().item()
if hasattr(lib, "ocm_simca_score_tc_variant"):
    lib.ocm_simca_score_tc_variant.argtypes = [C.c_uint]
    for v in range(8):
        lib.ocm_simca_score_tc_variant(v)
        dbg.zero_()
        run(lib.ocm_simca_score_tc_f32)
        D = dbg.double(); xc0 = xc[:128]; xh0 = (t @ pca.P)[:128]
        print("variant %d (noneg=%d swap=%d kmajor=%d): |D-r| %.3e |D-xc| %.3e |D-(xc+xh)| %.3e  status %d" % (
            v, v & 1, (v >> 1) & 1, (v >> 2) & 1, fro(D - r[:128]), fro(D - xc0), fro(D - xc0 - xh0), lib.ocm_simca_score_tc_status()))
    lib.ocm_simca_score_tc_variant(0)
dbg.zero_()
T2a, Qa, acca = run(lib.ocm_simca_score_tc_f32)
lib.ocm_simca_score_tc_debug(None)
print("tc status", lib.ocm_simca_score_tc_status())
T2b, Qb, accb = run(lib.ocm_simca_score_simt_f32, (None,))
Xd = X.double(); xc = Xd - pca.mean; t = xc @ pca.P.T; r = xc - t @ pca.P
T2r = (t * t / torch.as_tensor(pca.eigs_all[:A], device=dev)).sum(1); Qr = (r * r).sum(1)
D = dbg.double(); xc0 = xc[:128]; xh0 = (t @ pca.P)[:128]; r0 = r[:128]
fro = lambda m: m.pow(2).sum().sqrt().item()
print("tile0: |D-r| %.3e |D-xc| %.3e |D-(xc+xh)| %.3e |r| %.3e |xc| %.3e |xh| %.3e" % (fro(D - r0), fro(D - xc0), fro(D - xc0 - xh0), fro(r0), fro(xc0), fro(xh0)))
E = xc0 - D   # what the MMA subtracted; should equal xh0 = T P
# least squares: E ~ T * M  -> M (16 x 1024) tells which matrix was applied
T0 = t[:128]
M = torch.linalg.lstsq(T0, E).solution   # [A, 1024]
print("fit residual |E - T M| %.3e ; |M - P| %.3e |M| %.3e" % (fro(E - T0 @ M), fro(M - pca.P), fro(M)))
Pm = pca.P
best = [(int((M[a:a+1] @ Pm.T).abs().argmax()), float((M[a:a+1] @ Pm.T).abs().max())) for a in range(A)]
print("row a of M correlates with P row:", best)
c0 = M[0] / Pm[0]
print("M[0][:16]/P[0][:16]", [round(float(v), 3) for v in c0[:16]])
print("M[0][:12]", [round(float(v), 4) for v in M[0][:12]], "P[0][:12]", [round(float(v), 4) for v in Pm[0][:12]])
rel = lambda a, b: ((a.double() - b).abs() / b.abs()).max().item()
print("T2 rel err tc %.3e simt %.3e" % (rel(T2a, T2r), rel(T2b, T2r)))
print("Q  rel err tc %.3e simt %.3e" % (rel(Qa, Qr), rel(Qb, Qr)))
print("Q samples tc", Qa[:4].tolist(), "ref", Qr[:4].tolist(), "tail tc", Qa[-3:].tolist(), "ref", Qr[-3:].tolist())
bad = ((Qa.double() - Qr).abs() / Qr > 1e-4).nonzero().flatten()
print("bad Q rows:", bad.numel(), bad[:10].tolist(), "rows mod 128:", (bad[:10] % 128).tolist())
print("mean signed rel err T2 %.3e Q %.3e" % (((T2a.double() - T2r) / T2r).mean().item(), ((Qa.double() - Qr) / Qr).mean().item()))
print("decisions differ tc vs simt:", int((acca != accb).sum()))
for name, fn, extra in (("tc", lib.ocm_simca_score_tc_f32, ()), ("rowteam", lib.ocm_simca_score_rowteam_f32, ())):
    T2 = torch.empty(n, device=dev); Q = torch.empty(n, device=dev); acc = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(2): fn(vp(X), n, p, p, vp(mu32), vp(P32), vp(Sinv), C.byref(prm), vp(T2), vp(Q), vp(acc), *extra, st)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn(vp(X), n, p, p, vp(mu32), vp(P32), vp(Sinv), C.byref(prm), vp(T2), vp(Q), vp(acc), *extra, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {ms:.3f} ms  {n / ms / 1e6:.1f} M rows/s  {n * 4105 / ms / 1e6:.0f} GB/s")
