"""bring-up harness of the cluster scoring kernel: one launch, error string and protocol status word"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocm_vae_simca_b200 import native
lib = native.lib()
n, p, A = int(sys.argv[1]), int(sys.argv[2]), 10
dev = torch.device("cuda")
X = torch.randn(n, p, device=dev) * 0.1 + 1
mu = X[:1000].mean(0).contiguous(); P = torch.linalg.qr(torch.randn(p, A, device=dev))[0].T.contiguous()
S = torch.eye(A, device=dev).contiguous()
prm = native.ScoreParams(A=A, rule=1, snv=0, snv_eps=1e-8, ct=0.05, cq=10.0, dlim=1.41)
T2 = torch.zeros(n, device=dev); Q = torch.zeros(n, device=dev); acc = torch.zeros(n, dtype=torch.uint8, device=dev)
vp = lambda t: C.c_void_p(t.data_ptr())
lib.ocm_simca_score_tc_inject_stall.restype = C.c_int
if len(sys.argv) > 3: lib.ocm_simca_score_tc_inject_stall(int(sys.argv[3]))
rc = lib.ocm_simca_score_tc_f32(vp(X), n, p, p, vp(mu), vp(P), vp(S), C.byref(prm), vp(T2), vp(Q), vp(acc), None)
print("rc", rc, lib.ocm_last_error())
try:
    torch.cuda.synchronize(); print("sync ok")
    xc = (X - mu).double(); t = xc @ P.double().T; r = xc - t @ P.double()
    Qr = (r * r).sum(1); d = (Q.double() - Qr)
    bad = torch.nonzero(d.abs() > 1e-4 * Qr).flatten()
    print("bad rows", bad.numel(), "of", n, "first", bad[:12].tolist(), "tiles", sorted(set((bad // 128).tolist()))[:20])
    if bad.numel():
        i = int(bad[0]); boxq = (r[i] * r[i]).view(-1, 32).sum(1)
        xq = (xc[i] * xc[i]).view(-1, 32).sum(1)
        print("row", i, "diff", float(d[i]), "Q", float(Qr[i]), "box r2", [round(float(v), 3) for v in boxq[:8]], "box x2", [round(float(v), 3) for v in xq[:4]])
        print("diffs of first bad rows", [round(float(d[j]), 4) for j in bad[:16]])
    print("max rel err T2", float(((t * t).sum(1) - T2.double()).abs().max() / (t * t).sum(1).max()), "Q", float(((r * r).sum(1) - Q.double()).abs().max() / (r * r).sum(1).max()))
except Exception as e:
    print("sync failed:", str(e)[:80])
print("status", lib.ocm_simca_score_tc_status())
w = (C.c_uint32 * 16)(); lib.ocm_simca_score_tc_debug_words(w)
names = ["status", "C1", "C2g0", "C2g1", "CT", "TMA1", "TMA2", "MMAP", "MMAR"]
print("trace", {names[i]: (w[i] >> 8, w[i] & 255) for i in range(1, 9)})
os._exit(0)
