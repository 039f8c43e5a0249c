"""A/B timing of ocm_simca_score_tc_f32 across library builds (ctypes straight on the .so, nothing else of the
package is bound, so libraries built from older commits load too).

    python tools/ab_score.py --libs ab/lib_x.so ab/lib_y.so [--steps 20] [--rounds 2] [--rows 10000000] [--p 1024]

Per library and round: `steps` back-to-back launches after 3 warm-up launches, device-timed per launch; prints
first / min / median / last.  Libraries are interleaved so that box-to-box and thermal differences cancel."""
import argparse, ctypes as C, json, os
import torch


class ScoreParams(C.Structure):
    _fields_ = [("A", C.c_int32), ("rule", C.c_int32), ("snv", C.c_int32), ("snv_eps", C.c_float),
                ("ct", C.c_float), ("cq", C.c_float), ("dlim", C.c_float), ("x_rms", C.c_float)]      # ABI v2 (older libraries ignore the tail)


ap = argparse.ArgumentParser()
ap.add_argument("--libs", nargs="+", required=True)
ap.add_argument("--steps", type=int, default=20); ap.add_argument("--rounds", type=int, default=2)
ap.add_argument("--rows", type=int, default=10_000_000); ap.add_argument("--p", type=int, default=1024)
ap.add_argument("--A", type=int, default=10); ap.add_argument("--snv", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
n, p, A = a.rows, a.p, a.A
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
lam = ((X[:5000] - mu) @ P.T).var(0)
Sinv = torch.diag(1.0 / lam).contiguous()
T2 = torch.empty(n, device=dev); Q = torch.empty(n, device=dev); acc = torch.empty(n, dtype=torch.uint8, device=dev)
x_rms = float(((X[:5000] - mu) ** 2).mean().sqrt())
hp = ScoreParams(A, 1, a.snv, 1e-8, 0.05, 100.0, 1.4142135, 0.0)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
ref = None
for rnd in range(a.rounds):
    for path in a.libs[rnd % len(a.libs):] + a.libs[:rnd % len(a.libs)]:      # rotate: the first of a round meets the coolest GPU
        os.environ['OCM_SCORE_TMA2D'] = '1' if 'tma2d' in os.path.basename(path) else '0'
        os.environ['OCM_SCORE_TC2'] = '1' if 'tc2' in os.path.basename(path) else '0'
        hp.x_rms = x_rms if 'f16' in os.path.basename(path) else 0.0
        os.environ['OCM_SCORE_BF16A'] = '1' if 'bf16' in os.path.basename(path) else '0'   # read once per library at its first call
        lib = C.CDLL(os.path.abspath(path))
        fn = lib.ocm_simca_score_tc_f32
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(ScoreParams),
                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        run = lambda: fn(X.data_ptr(), n, p, p, mu.data_ptr(), P.data_ptr(), Sinv.data_ptr(), C.byref(hp), T2.data_ptr(),
                         Q.data_ptr(), acc.data_ptr(), stream)
        for _ in range(3):
            rc = run()
            assert rc == 0, (path, rc)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
        ev[0].record()
        for i in range(a.steps):
            run(); ev[i + 1].record()
        torch.cuda.synchronize()
        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.steps)]
        chk = (float(T2[:100000].double().sum()), float(Q[:100000].double().sum()), int(acc.sum()))
        if ref is None:
            ref = chk
            T2r, Qr = T2.clone(), Q.clone()
        bad = torch.nonzero(((T2 - T2r).abs() > 1e-4 * T2r.abs()) | ((Q - Qr).abs() > 1e-4 * Qr.abs()) | ~torch.isfinite(Q)).flatten()
        if bad.numel():
            tiles = torch.unique(bad // 128)
            print("  rows that differ from the first run:", bad.numel(), "in", tiles.numel(), "tiles; first tiles", tiles[:10].tolist(),
                  "rows in first tile", (bad[bad // 128 == tiles[0]] % 128).tolist()[:40], "relQ", float(((Q - Qr).abs() / Qr)[bad].max()), "relT2", float(((T2 - T2r).abs() / T2r)[bad].max()),
                  "Q/Qr of first rows", [round(float(Q[i] / Qr[i]), 2) for i in bad[:8]], "T2/T2r", [round(float(T2[i] / T2r[i]), 3) for i in bad[:8]], flush=True)
        print(json.dumps({"lib": os.path.basename(path), "round": rnd, "first": round(ms[0], 3), "min": round(min(ms), 3),
                          "median": round(sorted(ms)[len(ms) // 2], 3), "last": round(ms[-1], 3),
                          "rel_T2": abs(chk[0] - ref[0]) / ref[0], "rel_Q": abs(chk[1] - ref[1]) / ref[1], "acc": chk[2]}), flush=True)
