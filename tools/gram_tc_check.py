"""GPU check + timing of the integer tcgen05 Gram (csrc/gram_tc.cu) against numpy and the DMMA kernel.
usage: python tools/gram_tc_check.py [--big]"""
import json, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from ocm_vae_simca_b200 import engine, native, synth

dev = torch.device("cuda")


def exact_quantised_gram(X, shift, mn, mx, products):
    """numpy restatement of the kernel's arithmetic (digits in float64 matmuls are exact: |d| <= 128, n < 2^20)."""
    m = np.maximum(mx.astype(np.float64) - shift, shift - mn.astype(np.float64)) * (1 + 1e-6)
    sc = np.zeros(len(m), dtype=np.int64)
    ok = m > 0
    sc[ok] = 30 - np.frexp(m[ok])[1]
    v = np.rint((X.astype(np.float64) - shift) * np.ldexp(1.0, sc)).astype(np.int64)
    u = v + 0x808080
    d = [((u >> (8 * k)) & 0xFF) - 128 for k in range(3)] + [u >> 24]
    assert all(np.abs(x).max() <= 128 for x in d)
    assert np.array_equal(sum(d[k] << (8 * k) for k in range(4)), v)
    dmin = {10: 3, 13: 2, 16: 0}[products]
    G = np.zeros((X.shape[1], X.shape[1]), dtype=np.longdouble)
    for k in range(4):
        for l in range(4):
            if k + l >= dmin:
                G += (d[k].astype(np.float64).T @ d[l].astype(np.float64)).astype(np.longdouble) * np.longdouble(256.0) ** (k + l)
    inv = np.ldexp(1.0, -sc).astype(np.longdouble)
    Gq = np.asarray(G * inv[:, None] * inv[None, :], dtype=np.float64)
    np.fill_diagonal(Gq, np.asarray((v.astype(np.longdouble) ** 2).sum(0) * inv * inv, dtype=np.float64))   # exact diagonal
    return Gq, (v.sum(0) * np.ldexp(1.0, -sc))


def check(n, p, dtype, products, gather=False, seed=0):
    rng = np.random.default_rng(seed)
    X = (rng.normal(1.0, 0.3, (n, p)) * np.linspace(0.5, 2.0, p)).astype(dtype)
    rows = np.sort(rng.choice(n, size=(2 * n) // 3, replace=False)) if gather else None
    Xs = X if rows is None else X[rows]
    Xd = torch.as_tensor(X, device=dev)
    idx = None if rows is None else engine.index_tensor(rows, dev)
    s, mn, mx, cnt = engine.colstats_minmax(Xd, idx)
    np.testing.assert_allclose(s.cpu().numpy(), Xs.astype(np.float64).sum(0), rtol=1e-12)
    assert np.all(mn.cpu().numpy() <= Xs.min(0)) and np.all(mx.cpu().numpy() >= Xs.max(0))
    shift = s / cnt
    G, cs, flags = engine.gram_tc(Xd, idx, shift, mn, mx, products=products)
    torch.cuda.synchronize()
    assert flags.cpu().tolist() == [0, 0], flags
    got = np.triu(G.cpu().numpy())
    ref_q, cs_q = exact_quantised_gram(Xs, shift.cpu().numpy(), mn.cpu().numpy(), mx.cpu().numpy(), products)
    ref_q = np.triu(ref_q)
    Y = Xs.astype(np.float64) - shift.cpu().numpy()
    ref = np.triu(Y.T @ Y)
    scale = np.sqrt(np.outer(np.diag(ref), np.diag(ref)))
    e_q = np.abs(got - ref_q).max() / np.abs(ref_q).max()
    e_r = (np.abs(got - ref) / scale).max()
    e_cs = np.abs(cs.cpu().numpy() - cs_q).max()
    print(json.dumps({"n": n, "p": p, "dtype": np.dtype(dtype).name, "products": products, "gather": gather,
                      "err_vs_quantised_model": float(e_q), "err_vs_fp64_gram_rel_diag": float(e_r), "colsum_err": float(e_cs)}), flush=True)
    return e_q, e_r


def timing(n, p, products):
    X = torch.empty((n, p), dtype=torch.float32, device=dev)
    synth.make_spectra_device(n, p, dev, seed=5, out=X)
    s, mn, mx, cnt = engine.colstats_minmax(X)
    shift = s / cnt
    G = torch.zeros((p, p), dtype=torch.float64, device=dev)
    cs = torch.zeros(p, dtype=torch.float64, device=dev)
    for _ in range(2):
        engine.gram_tc(X, None, shift, mn, mx, products=products, out=G, colsum=cs)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    engine.colstats_minmax(X)
    ev[1].record()
    reps = 3
    for _ in range(reps):
        G.zero_(); cs.zero_()
        _, _, flags = engine.gram_tc(X, None, shift, mn, mx, products=products, out=G, colsum=cs)
    ev[2].record()
    torch.cuda.synchronize()
    t_stats = ev[0].elapsed_time(ev[1])
    t_gram = ev[1].elapsed_time(ev[2]) / reps
    flop = n * p * (p + 1.0)
    out = {"n": n, "p": p, "products": products, "colstats_minmax_ms": t_stats, "gram_tc_ms": t_gram,
           "algorithmic_tflops": flop / (t_gram * 1e-3) / 1e12, "flags": flags.cpu().tolist()}
    # against the DMMA kernel on a slice of the same rows
    m = min(n, 400_000)
    Gd = engine.gram(X[:m], None, shift)
    Gt, _, _ = engine.gram_tc(X[:m], None, shift, mn, mx, products=products)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); Gd = engine.gram(X[:m], None, shift); torch.cuda.synchronize(); t_d = time.perf_counter() - t0
    d = torch.triu(Gt - Gd).abs().max().item() / torch.triu(Gd).abs().max().item()
    sym = lambda G_: (torch.triu(G_) + torch.triu(G_, 1).T) / (m - 1)
    ev_d = np.sort(engine.eigvals_sym(sym(Gd)).cpu().numpy())[::-1][:40]
    ev_t = np.sort(engine.eigvals_sym(sym(Gt)).cpu().numpy())[::-1][:40]
    out["eig_rel_diff_top20_vs_dmma"] = float(np.max(np.abs(ev_t[:20] - ev_d[:20]) / ev_d[:20]))
    out["eig_rel_diff_21_40_vs_dmma"] = float(np.max(np.abs(ev_t[20:] - ev_d[20:]) / ev_d[20:]))
    out["lambda_1_20_40"] = [float(ev_d[0]), float(ev_d[19]), float(ev_d[39])]
    out.update({"dmma_rows": m, "dmma_ms": t_d * 1e3, "dmma_tflops": m * p * (p + 1.0) / t_d / 1e12, "max_diff_vs_dmma_rel": d})
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    worst = 0.0
    for n, p, dt, g in ((300, 128, np.float32, False), (1000, 256, np.float32, False), (2000, 700, np.float64, True),
                        (5000, 1050, np.float32, True), (40000, 384, np.float32, False), (70000, 256, np.float32, True)):
        for pr in (16, 13, 10):
            e_q, e_r = check(n, p, dt, pr, g)
            worst = max(worst, e_q)
    print("worst error against the quantised integer model:", worst)
    if "--big" in sys.argv:
        for pr in (10, 13, 16):
            timing(2_000_000, 2048, pr)
        timing(4_000_000, 1024, 13)
