"""Generate tests/golden/*.npz by running the LIVE reference from /root/reference.

TEST INFRASTRUCTURE.  Runs only in the build container (the GPU box has no
/root/reference); the fixtures it writes are committed.  The reference is imported
unmodified: `utils/SIMCA.py:8-9` imports matplotlib/plotly at module level, which are
not installed, so empty stub modules are registered first.  The
`VAESIMCA` class lives in a script that opens Windows paths at import
(`VAE_SIMCA.py:388`), so its class body (lines 215-382) is exec'd from the file.

    python oracle/make_golden.py            # small fixtures (seconds)
    python oracle/make_golden.py --c2       # + the full 2000x1050 CV table (~3 min)
"""
from __future__ import annotations

import argparse
import importlib.util
import io
import os
import sys
import types
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def _load_synth():
    spec = importlib.util.spec_from_file_location(
        "_ocm_synth", os.path.join(ROOT, "ocm-vae-simca_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "plotly", "plotly.graph_objects"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["plotly"].graph_objects = sys.modules["plotly.graph_objects"]
    sys.path.insert(0, REF)
    import utils as ref_utils  # noqa: the reference's package
    import vae_model as ref_vae
    sys.path.pop(0)
    return ref_utils, ref_vae


def quiet(fn, *a, **k):
    with redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def golden_c1(ref, synth):
    """Config 1: 200 target + 100 other x 700, SNV, A=5, default limits."""
    X, y = synth.make_class_dataset(200, 100, 700, seed=0)
    np.random.seed(0)
    m = ref.SIMCA(n_components=5, model_class=0, verbose=False)
    quiet(m.fit, X, y)
    info = m._model[0]
    pred = quiet(m.predict, X)
    T2, T2red, Q, Qred = m.transform(X)
    np.savez_compressed(
        os.path.join(GOLD, "c1_fit_predict.npz"),
        seed=0, n_target=200, n_other=100, p=700, A=5,
        xmean=info["xmean"], P=info["P"], eigs_all=info["eigs_all"], invcovT=info["invcovT"],
        T2_train=info["T2"], Q_train=info["Q"],
        T2_limit=info["T2_limit"], Q_limit=info["Q_limit"], D_limit=info["D_limit"],
        pred=pred, T2=T2, T2red=T2red, Q=Q, Qred=Qred,
        pred_P=info["pca_model"].components_, pred_mean=info["pca_model"].mean_)


def golden_limits(ref, synth):
    """All 4 decision rules x 5 T2 limits x 4 Q limits on a 160x90 problem, A=4."""
    X, y = synth.make_class_dataset(120, 40, 90, seed=11)
    kinds = ["sim", "alt", "ci", "dd"]
    t2s = ["perc", "Fdistrig", "Fdist", "chi2", "chi2pom"]
    qs = ["perc", "jm", "chi2box", "chi2pom"]
    rows, preds, names = [], [], []
    for k in kinds:
        for t in t2s:
            for q in qs:
                np.random.seed(0)
                m = ref.SIMCA(n_components=4, model_class=0, type=k, t2lim=t, qlim=q, verbose=False)
                quiet(m.fit, X, y)
                info = m._model[0]
                rows.append([info["T2_limit"], info["Q_limit"], info["D_limit"]])
                preds.append(quiet(m.predict, X)[:, 0])
                names.append(f"{k}/{t}/{q}")
    np.savez_compressed(os.path.join(GOLD, "limits_grid.npz"), seed=11, n_target=120,
                        n_other=40, p=90, A=4, names=np.array(names),
                        limits=np.array(rows, dtype=np.float64),
                        preds=np.array(preds, dtype=np.uint8))


def golden_ncomp_sweep(ref, synth):
    """A = 2..20 on 400x300: T2/Q of held-out rows per A (nested-loadings check)."""
    X, y = synth.make_class_dataset(300, 100, 300, seed=21)
    Xte, yte = synth.make_class_dataset(60, 60, 300, seed=22)
    T2s, Qs, lims = [], [], []
    for A in range(2, 21):
        np.random.seed(0)
        m = ref.SIMCA(n_components=A, model_class=0, verbose=False)
        quiet(m.fit, X, y)
        T2, _, Q, _ = m.transform(Xte)
        T2s.append(T2); Qs.append(Q)
        info = m._model[0]
        lims.append([info["T2_limit"], info["Q_limit"], info["D_limit"]])
    np.savez_compressed(os.path.join(GOLD, "ncomp_sweep.npz"), seed=21, seed_test=22,
                        A_values=np.arange(2, 21), T2=np.array(T2s), Q=np.array(Qs),
                        limits=np.array(lims))


def golden_cv_small(ref, synth):
    """cross_validate_simca_grid on 360x150, venetian blinds 6 folds, LV 2..8."""
    from sklearn.model_selection import PredefinedSplit
    X, y = synth.make_class_dataset(270, 90, 150, seed=31)
    cv = PredefinedSplit(np.arange(X.shape[0]) % 6)
    np.random.seed(0)
    res = quiet(ref.cross_validate_simca_grid, ref.SIMCA(model_class=0, verbose=False), X, y, cv,
                LV_min=2, LV_max=8, class_index=0, print_summary=False, store_predictions=True)
    table = np.array([[r["LV"], r["spec"], r["sens"], r["eff"]] for r in res["results"]])
    pooled = np.array([b["prediction"] for b in res["by_combo"]], dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLD, "cv_small.npz"), seed=31, n_target=270, n_other=90,
                        p=150, folds=6, table=table, pooled=pooled,
                        best_LV=res["best_LV"], best_score=res["best_score"])
    # ClasswiseKFoldWithExternalVal splitter indices
    sp = ref.ClasswiseKFoldWithExternalVal(n_splits=4, cls_label=0)
    idx = [(tr, te) for tr, te in sp.split(X, y)]
    np.savez_compressed(os.path.join(GOLD, "classwise_kfold.npz"), seed=31,
                        **{f"tr{i}": t for i, (t, _) in enumerate(idx)},
                        **{f"te{i}": t for i, (_, t) in enumerate(idx)})


def golden_c2(ref, synth):
    """Config 2 at full size: 2000x1050, 10-fold venetian blinds, LV 2..20 (LV=1 crashes
    in the reference, SIMCA.py:69).  ~3 minutes of reference CPU time."""
    from sklearn.model_selection import PredefinedSplit
    import time
    X, y = synth.make_class_dataset(1500, 500, 1050, seed=1234)
    cv = PredefinedSplit(np.arange(X.shape[0]) % 10)
    np.random.seed(0)
    t0 = time.perf_counter()
    res = quiet(ref.cross_validate_simca_grid, ref.SIMCA(model_class=0, verbose=False), X, y, cv,
                LV_min=2, LV_max=20, class_index=0, print_summary=False, store_predictions=True)
    dt = time.perf_counter() - t0
    table = np.array([[r["LV"], r["spec"], r["sens"], r["eff"]] for r in res["results"]])
    pooled = np.packbits(np.array([b["prediction"] for b in res["by_combo"]], dtype=np.uint8), axis=1)
    info = res["best_estimator"]._model[0]
    np.savez_compressed(os.path.join(GOLD, "c2_cv_grid.npz"), seed=1234, n_target=1500,
                        n_other=500, p=1050, folds=10, table=table, pooled_packed=pooled,
                        best_LV=res["best_LV"], best_score=res["best_score"],
                        best_T2_limit=info["T2_limit"], best_Q_limit=info["Q_limit"],
                        ref_seconds=dt, ref_cores=os.cpu_count())
    print(f"c2 reference grid: {dt:.1f} s on {os.cpu_count()} cores")


def golden_vae(ref_vae):
    """compute_q_h_f (vae_model.py:162-181) and VAESIMCA (VAE_SIMCA.py:215-382)."""
    import torch
    from scipy import special
    torch.manual_seed(0)
    B, L, d = 256, 96, 6
    x = torch.randn(B, L)
    x_rec = x + 0.05 * torch.randn(B, L)
    z = torch.randn(B, d) @ torch.randn(d, d)
    q, h, f, qc, hc, fc = ref_vae.compute_q_h_f(x, x_rec, z)
    np.savez_compressed(os.path.join(GOLD, "vae_qhf.npz"), x=x.numpy(), x_rec=x_rec.numpy(),
                        z=z.numpy(), q=q.numpy(), h=h.numpy(), f=f.numpy(),
                        crit=np.array([qc, hc, fc]))

    src = open(os.path.join(REF, "VAE_SIMCA.py")).read().splitlines()
    ns = {"np": np, "torch": torch, "special": special}
    exec("\n".join(src[214:382]), ns)  # class VAESIMCA only
    VAESIMCA = ns["VAESIMCA"]
    torch.manual_seed(1)
    mean = np.zeros(L, dtype=np.float32); std = np.ones(L, dtype=np.float32)
    vae = ref_vae.ConvVAE1D(L, d, mean, std, conv_blocks=2, n_filters=4, kernel_size=5,
                            hidden_fc=16)
    vae.eval()
    Xcal = torch.randn(200, L); Xte = torch.randn(80, L) * 1.3
    out = {"state_keys": np.array(list(vae.state_dict().keys()))}
    for k, v in vae.state_dict().items():
        out["sd::" + k] = v.numpy()
    combos = [("alt", "Fdist", "jm"), ("sim", "perc", "perc"), ("dd", "chi2pom", "chi2pom"),
              ("alt", "chi2", "jm")]
    for kind, t2l, ql in combos:
        vs = VAESIMCA(vae, type=kind, t2lim=t2l, qlim=ql, device="cpu", verbose=False)
        vs.fit_thresholds([(Xcal,)], class_label=0)
        yp, T2, Q = vs.predict([(Xte,)])
        mi = vs._model[0]
        tag = f"{kind}_{t2l}_{ql}"
        out[tag + "::T2"] = T2; out[tag + "::Q"] = Q; out[tag + "::pred"] = yp.astype(np.uint8)
        out[tag + "::limits"] = np.array([mi["T2_limit"], mi["Q_limit"], mi["D_limit"]], dtype=np.float64)
        out[tag + "::T2_cal"] = mi["T2"]; out[tag + "::Q_cal"] = mi["Q"]
    np.savez_compressed(os.path.join(GOLD, "vae_simca.npz"), Xcal=Xcal.numpy(), Xte=Xte.numpy(),
                        L=L, d=d, **out)


def golden_splits(ref, synth):
    """object_aware_splits of the live reference on a small acquisition-shaped set (3 nut types x 8 objects, p=120;
    every per-nut matrix has max(shape) <= 500 so the reference's PCA(svd_solver='auto') is the exact 'full' solver and
    the fixture is deterministic)."""
    kw = dict(nuts=("almond", "hazelnut", "walnut"), n_objects=8, rows=(30, 60), p=120, seed=5)
    data = synth.make_object_dataset(**kw)
    assert all(sum(o["spectral_data"].shape[0] for o in v) <= 500 for v in data.values())
    splits, Xts, yts, Xcal, Xval, Xtin, Xtout = quiet(ref.object_aware_splits, data, list(kw["nuts"]), "almond", 120)
    np.savez_compressed(os.path.join(GOLD, "object_splits.npz"), seed=5, Xts=Xts, yts=yts, Xcal=Xcal, Xval=Xval,
                        Xtin=Xtin, Xtout=Xtout,
                        shapes=np.array([[splits[n][k].shape[0] for k in ("cal", "val", "test")] for n in kw["nuts"]]))
    # the preprocessing + distance step on its own (data_utils.py:57-77 re-run here line by line with the same
    # library calls) so that the kernels have a numeric target too
    from scipy.signal import savgol_filter
    from sklearn.decomposition import PCA
    X = np.vstack([o["spectral_data"] for o in data["hazelnut"]])
    X = X[np.isfinite(X).all(axis=1)]
    Xp = (X - np.mean(X, axis=1, keepdims=True)) / (np.std(X, axis=1, keepdims=True) + 1e-8)
    Xp = savgol_filter(Xp, window_length=5, polyorder=2, deriv=1, axis=1)
    T = PCA(n_components=10).fit_transform(Xp)
    ci = np.linalg.pinv(np.cov(T, rowvar=False))
    mahal = np.array([np.sqrt((t - T.mean(0)) @ ci @ (t - T.mean(0)).T) for t in T])
    np.savez_compressed(os.path.join(GOLD, "snv_savgol_mahal.npz"), X=X, Xp=Xp, mahal=mahal,
                        thr=np.percentile(mahal, 95))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c2", action="store_true")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    ref, ref_vae = import_reference()
    synth = _load_synth()
    golden_c1(ref, synth)
    golden_limits(ref, synth)
    golden_ncomp_sweep(ref, synth)
    golden_cv_small(ref, synth)
    golden_vae(ref_vae)
    golden_splits(ref, synth)
    if args.c2:
        golden_c2(ref, synth)
    print("fixtures written to", GOLD)


if __name__ == "__main__":
    main()
