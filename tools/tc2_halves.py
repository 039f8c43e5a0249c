"""bring-up: which cluster rank's partial Q goes wrong?  (OCM_TC2_POL=1024 makes rank 0 write its own partial Q into Q
and the peer's into T2)"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocm_vae_simca_b200 import native
lib = native.lib()
n, p, A = int(sys.argv[1]), 1024, 10
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(1)
w = torch.linspace(0, 1, p, device=dev)
basis = torch.stack([torch.exp(-((w - c) / 0.05) ** 2) for c in torch.linspace(0.05, 0.95, 30).tolist()])
X = torch.empty(n, p, device=dev)
for i in range(0, n, 500_000):
    m = min(500_000, n - i)
    c = torch.randn(m, 30, device=dev, generator=g) * (0.8 ** torch.arange(30, device=dev))
    X[i:i + m] = 1 + 0.2 * c @ basis + 1e-3 * torch.randn(m, p, device=dev, generator=g)
mu = X[:5000].mean(0).contiguous()
_, _, Vt = torch.linalg.svd(X[:5000] - mu, full_matrices=False)
P = Vt[:A].contiguous()
S = torch.eye(A, device=dev).contiguous()
prm = native.ScoreParams(A=A, rule=1, snv=0, snv_eps=1e-8, ct=0.05, cq=10.0, dlim=1.41)
T2 = torch.zeros(n, device=dev); Q = torch.zeros(n, device=dev); acc = torch.zeros(n, dtype=torch.uint8, device=dev)
vp = lambda t: C.c_void_p(t.data_ptr())
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    rc = lib.ocm_simca_score_tc_f32(vp(X), n, p, p, vp(mu), vp(P), vp(S), C.byref(prm), vp(T2), vp(Q), vp(acc), None)
    torch.cuda.synchronize()
    bad0 = bad1 = 0; ex = []
    for i in range(0, n, 500_000):
        xc = (X[i:i + 500_000] - mu).double(); t = xc @ P.double().T; r = xc - t @ P.double()
        mode = int(os.environ.get("OCM_TC2_POL", "0"))
        if mode & 8192: r = xc
        if mode & 16384: r = t @ P.double()
        q0 = (r[:, :512] ** 2).sum(1); q1 = (r[:, 512:] ** 2).sum(1)
        b0 = torch.nonzero((Q[i:i + 500_000].double() - q0).abs() > 1e-4 * q0).flatten()
        b1 = torch.nonzero((T2[i:i + 500_000].double() - q1).abs() > 1e-4 * q1).flatten()
        bad0 += b0.numel(); bad1 += b1.numel()
        for b, qq, out, nm in ((b0, q0, Q, "r0"), (b1, q1, T2, "r1")):
            for j in b[:3].tolist():
                if len(ex) < 12:
                    # which 32-column box of the half explains the error?  residual energy of the row if one box used a stale x
                    ex.append((nm, i + j, round(float(out[i + j] / qq[j]), 5)))
    print("rep", rep, "rc", rc, "bad rows: rank 0 half", bad0, " rank 1 half", bad1, ex, flush=True)

# ---- which 16-byte chunk is wrong, and whose data does it hold?  (only meaningful with OCM_TC2_POL = 8192 + 1024) ----
if int(os.environ.get("OCM_TC2_POL", "0")) & 8192:
    ncl = 74
    xcf = lambda rows, cols: (X[rows][:, cols] - mu[cols]).double()
    true_half = lambda i: float((xcf(torch.tensor([i], device=dev), torch.arange(0, 512, device=dev)) ** 2).sum())
    qtrue = torch.empty(n, dtype=torch.float64, device=dev)
    for i in range(0, n, 500_000):
        qtrue[i:i + 500_000] = ((X[i:i + 500_000, :512] - mu[:512]).double() ** 2).sum(1)
    bad = torch.nonzero((Q.double() - qtrue).abs() > 1e-4 * qtrue).flatten()[:12].tolist()
    for i in bad:
        meas, tru = float(Q[i]), float(qtrue[i])
        best = []
        for j in range(16):
            for c in range(8):
                cols = torch.arange(32 * j + 4 * c, 32 * j + 4 * c + 4, device=dev)
                e_true = float(((X[i, cols] - mu[cols]).double() ** 2).sum())
                for name, (ri, shift) in {"prev box same tile": (i, -128), "next box same tile": (i, 128),
                                          "prev tile": (i - ncl * 128, 512 - 128), "next tile": (i + ncl * 128, 128 - 512)}.items():
                    sc = cols + shift
                    if ri < 0 or ri >= n or int(sc.min()) < 0 or int(sc.max()) >= 512: continue
                    e_st = float(((X[ri, sc] - mu[cols]).double() ** 2).sum())
                    best.append((abs(tru - e_true + e_st - meas) / tru, name, j, c))
        best.sort()
        print("row", i, "tile", i // 128, "lane", i % 128, "meas/true", round(meas / tru, 5), "best explanations", [(f"{b[0]:.1e}", b[1], b[2], b[3]) for b in best[:3]], flush=True)
