"""CPU ORACLE for the SIMCA / DD-SIMCA hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file; the product (ocm-vae-simca_b200/) never does and fails
loudly when its CUDA library is missing.

What it is: a numpy / scikit-learn / scipy restatement of the algorithm of
/root/reference/utils/SIMCA.py and /root/reference/utils/CVSIMCA.py, written as plain
functions over dicts.  The arithmetic of the reference lives in third-party
libraries that are NOT under /root/reference and are not version-pinned by it (no
requirements file): sklearn.decomposition.PCA (1.9.0 here), numpy 2.3.5
einsum/cov/pinv/percentile, scipy 1.18.1 stats.f/chi2.ppf and special.erfinv.  The
oracle calls the same library routines at the same call sites, so that it is both the
parity checker and a faithful CPU timing stand-in.

Parity pin: the reference holds NO tests, golden vectors or fixtures for this path
(SURVEY.md §4).  The oracle is therefore pinned against outputs of the reference
itself, imported live from /root/reference by oracle/make_golden.py in the build
container and committed under tests/golden/*.npz; tests/test_oracle_golden.py
replays every fixture through this file.

Reference quirks restated on purpose:
  * the predict-time model is a SECOND PCA fit with svd_solver='auto' (randomized
    SVD at the benchmark shapes, unseeded) -- `faithful_solver=True` reproduces
    that for timing; parity runs pin svd_solver='full';
  * n_components == 1 crashes in the reference (np.cov of one column is 0-d);
    the oracle extends it the obvious way (1x1 covariance);
  * type='dd' forces both limits to 'chi2pom'.
"""
from __future__ import annotations

import numpy as np
import scipy.stats as stats
from scipy.special import erfinv
from sklearn.decomposition import PCA


# --------------------------------------------------------------------------------------
# critical limits  (reference: utils/SIMCA.py:156-236)
# --------------------------------------------------------------------------------------
def t2_limit(T2, A, kind, cl):
    """Hotelling T^2 critical value; returns (limit, dof, scale).  SIMCA.py:156-182."""
    n = len(T2)
    dof = scale = None
    if kind == "perc":
        lim = np.percentile(T2, cl * 100)
    elif kind == "Fdistrig":
        lim = (A / n) * (n ** 2 - 1) / (n - A) * stats.f.ppf(cl, A, n - A)
    elif kind == "Fdist":
        lim = A * (n - 1) / (n - A) * stats.f.ppf(cl, A, n - A)
    elif kind == "chi2":
        lim = stats.chi2.ppf(cl, A)
    elif kind == "chi2pom":
        h0 = float(np.mean(T2))
        v = float(np.var(T2, ddof=1)) if n > 1 else 0.0
        dof = max(int(np.round(2 * h0 ** 2 / v)) if v > 0 else 1, 1)
        lim = h0 * stats.chi2.ppf(cl, dof) / dof
        scale = h0
    else:
        raise ValueError(kind)
    return lim, dof, scale


def q_limit(Q, eigs_all, A, kind, cl):
    """Residual (Q / SPE) critical value; returns (limit, dof, scale).  SIMCA.py:184-217."""
    dof = scale = None
    tail = np.asarray(eigs_all)[A:]
    if kind == "perc":
        lim = np.percentile(Q, cl * 100)
    elif kind == "jm":
        th1, th2, th3 = tail.sum(), (tail ** 2).sum(), (tail ** 3).sum()
        h0 = max(1 - (2 * th1 * th3) / (3 * th2 ** 2), 0.001)
        ca = np.sqrt(2) * erfinv(2 * cl - 1)
        h1 = ca * np.sqrt(2 * th2 * h0 ** 2) / th1
        h2 = th2 * h0 * (h0 - 1) / th1 ** 2
        lim = th1 * (h1 + 1 + h2) ** (1 / h0)
    elif kind == "chi2box":
        th1, th2 = tail.sum(), (tail ** 2).sum()
        lim = (th2 / th1) * stats.chi2.ppf(cl, th1 ** 2 / th2)
    elif kind == "chi2pom":
        v0 = np.mean(Q)
        dof = max(round(2 * v0 ** 2 / np.var(Q, ddof=1)), 1)
        lim = v0 * stats.chi2.ppf(cl, dof) / dof
        scale = v0
    else:
        raise ValueError(kind)
    return lim, dof, scale


def d_limit(kind, A, t2lim, qlim, eigs_all, dcl, t2dof=None, qdof=None):
    """Critical full distance.  SIMCA.py:219-236."""
    if kind == "sim":
        return 1
    if kind == "alt":
        return np.sqrt(2)
    if kind == "ci":
        tail = np.asarray(eigs_all)[A:]
        th1, th2 = tail.sum(), (tail ** 2).sum()
        tr1 = A / t2lim + th1 / qlim
        tr2 = A / t2lim ** 2 + th2 / qlim ** 2
        return (tr2 / tr1) * stats.chi2.ppf(dcl, tr1 ** 2 / tr2)
    if kind == "dd":
        return stats.chi2.ppf(dcl, t2dof + qdof)
    raise ValueError(kind)


# --------------------------------------------------------------------------------------
# fit / score  (reference: utils/SIMCA.py:62-99, :101-154)
# --------------------------------------------------------------------------------------
def fit_one_class(X, A, kind="alt", t2lim="Fdist", t2cl=0.95, qlim="jm", qcl=0.95,
                  dcl=0.95, faithful_solver=False):
    """One-class PCA model + limits.  SIMCA.py:62-99 (and :42-48 for the dd override)."""
    if kind == "dd":
        t2lim = qlim = "chi2pom"
    full = PCA(n_components=None, svd_solver="full")
    T = full.fit_transform(X)[:, :A]
    P = full.components_[:A]
    R = X - (T @ P + full.mean_)
    cov = np.atleast_2d(np.cov(T, rowvar=False))
    Sinv = np.linalg.pinv(cov)
    T2 = np.einsum("ij,jk,ik->i", T, Sinv, T)
    Q = np.sum(R ** 2, axis=1)
    tl, tdof, tsc = t2_limit(T2, A, t2lim, t2cl)
    ql, qdof, qsc = q_limit(Q, full.explained_variance_, A, qlim, qcl)
    dl = d_limit(kind, A, tl, ql, full.explained_variance_, dcl, tdof, qdof)
    # the reference keeps a second decomposition for predict (SIMCA.py:75)
    second = PCA(A, svd_solver="auto" if faithful_solver else "full").fit(X)
    return dict(kind=kind, n_components=A, mean=full.mean_, P=P, Sinv=Sinv,
                eigs_all=full.explained_variance_, T=T, T2=T2, Q=Q,
                T2_limit=tl, Q_limit=ql, D_limit=dl, t2dof=tdof, t2scale=tsc,
                qdof=qdof, qscale=qsc, n_samples=X.shape[0],
                pred_mean=second.mean_, pred_P=second.components_)


def score(model, X):
    """(T2, Q) of new rows with the predict-time decomposition.  SIMCA.py:127-130."""
    Xc = X - model["pred_mean"]
    T = Xc @ model["pred_P"].T
    R = X - (T @ model["pred_P"] + model["pred_mean"])
    Q = np.sum(R ** 2, axis=1)
    T2 = np.einsum("ij,jk,ik->i", T, model["Sinv"], T)
    return T2, Q


def reduced_distance(model, T2, Q):
    """Full reduced distance by rule.  SIMCA.py:131-144."""
    k = model["kind"]
    if k == "dd":
        return model["t2dof"] * T2 / model["t2scale"] + model["qdof"] * Q / model["qscale"]
    a, b = T2 / model["T2_limit"], Q / model["Q_limit"]
    if k == "sim":
        return np.maximum(a, b)
    if k == "alt":
        return np.sqrt(a ** 2 + b ** 2)
    if k == "ci":
        return a + b
    raise ValueError(k)


def predict(model, X):
    """0./1. acceptance (strict '<').  SIMCA.py:145."""
    T2, Q = score(model, X)
    return (reduced_distance(model, T2, Q) < model["D_limit"]).astype(np.float64)


def conformity(y_true, y_pred, class_index):
    """TP/TN/FP/FN and the percentage figures.  SIMCA.py:238-266."""
    tc = (np.asarray(y_true) == class_index).astype(int)
    yp = np.asarray(y_pred)
    TP = int(np.sum((yp == 1) & (tc == 1)))
    TN = int(np.sum((yp == 0) & (tc == 0)))
    FP = int(np.sum((yp == 1) & (tc == 0)))
    FN = int(np.sum((yp == 0) & (tc == 1)))
    with np.errstate(divide="ignore", invalid="ignore"):
        sens = np.float64(TP) / (TP + FN) * 100
        spec = np.float64(TN) / (TN + FP) * 100
        acc = np.float64(TP + TN) / (TP + TN + FP + FN) * 100
        eff = np.sqrt(sens * spec)
    return dict(sensitivity=sens, specificity=spec, accuracy=acc, efficiency=eff,
                TP=TP, TN=TN, FP=FP, FN=FN)


# --------------------------------------------------------------------------------------
# cross-validation grid  (reference: utils/CVSIMCA.py:103-269, single-class use)
# --------------------------------------------------------------------------------------
def cv_grid(X, y, splits, lv_values, model_class=0, class_index=None, refit=True,
            faithful_solver=False, **simca_kw):
    """Folds x n-components grid exactly as CVSIMCA.py:157-233 walks it for one parameter
    combination: per (lv, fold) a fresh fit on the target-class rows of the training part
    and a prediction of the test part; spec = mean over folds, sens from the pooled
    prediction vector, eff = sqrt(sens*spec); argmax over lv; optional refit."""
    ci = model_class if class_index is None else class_index
    records, pooled = [], {}
    for lv in lv_values:
        pred_vec = np.zeros(X.shape[0])
        spec_f = np.zeros(len(splits))
        for f, (tr, te) in enumerate(splits):
            Xm, Lm = X[tr], y[tr]
            m = fit_one_class(Xm[Lm == model_class], lv, faithful_solver=faithful_solver, **simca_kw)
            yp = predict(m, X[te])
            pred_vec[te] = yp
            spec_f[f] = conformity(y[te], yp, ci)["specificity"]
        spec = float(np.mean(spec_f))
        sens = float(conformity(y, pred_vec, ci)["sensitivity"])
        records.append(dict(LV=lv, spec=spec, sens=sens, eff=float(np.sqrt(sens * spec))))
        pooled[lv] = pred_vec
    best = int(np.argmax([r["eff"] for r in records]))
    out = dict(results=records, best_LV=records[best]["LV"], best_score=records[best]["eff"],
               pooled=pooled)
    if refit:
        out["best_model"] = fit_one_class(X[y == model_class], out["best_LV"],
                                          faithful_solver=faithful_solver, **simca_kw)
    return out


def snv(X, eps=1e-8):
    """Row SNV as the drivers apply it.  simca_nuts.py:48-50."""
    return (X - X.mean(axis=1, keepdims=True)) / (X.std(axis=1, keepdims=True) + eps)


def snv_savgol(X, window_length=5, polyorder=2, deriv=1):
    """SNV followed by the Savitzky-Golay derivative of the drivers (simca_nuts.py:48-52, utils/data_utils.py:57-59)."""
    from scipy.signal import savgol_filter
    return savgol_filter(snv(X), window_length=window_length, polyorder=polyorder, deriv=deriv, axis=1)


def mahalanobis_filter(X, n_components=10, percentile=95, faithful_solver=False):
    """PCA score-space outlier cut (simca_nuts.py:127-147, utils/data_utils.py:65-80): returns (keep mask, distances, threshold).
    The reference's PCA(n_components) runs svd_solver='auto', i.e. an UNSEEDED randomized SVD at these shapes (quirk Q1
    again: its distances move by ~5e-4 from call to call); parity pins the exact decomposition."""
    T = PCA(n_components=n_components, svd_solver="auto" if faithful_solver else "full").fit_transform(X)
    diff = T - T.mean(axis=0)
    cov_inv = np.linalg.pinv(np.cov(T, rowvar=False))
    d = np.sqrt(np.einsum("ij,jk,ik->i", diff, cov_inv, diff))
    thr = np.percentile(d, percentile)
    return d <= thr, d, thr


def object_aware_splits(data, nut_types, target_nut, p, cal_frac=0.7, val_frac=0.15, test_frac=0.15,
                        random_state=42, outlier_percentile=95, use_pca=True, faithful_solver=False):
    """Restatement of utils/data_utils.py:12-169: per nut type stack the objects, drop non-finite rows (:49-54), drop
    rows beyond the P-th percentile of the PCA(<=10) score-space Mahalanobis distance of the SNV + SavGol copy
    (:57-80), send whole objects to cal / val / test with two seeded train_test_split calls (:103-120), then assemble
    the pooled test set with labels 0 = target / 1 = other (:133-167)."""
    from sklearn.model_selection import train_test_split
    empty = lambda: np.empty((0, p), dtype=np.float32)
    splits = {}
    for nut in nut_types:
        objs = [np.asarray(o["spectral_data"], dtype=np.float32) for o in data[nut]]
        splits[nut] = dict(cal=empty(), val=empty(), test=empty())
        if not objs:
            continue
        X = np.vstack(objs)
        ids = np.concatenate([np.full(len(o), i, dtype=int) for i, o in enumerate(objs)])
        ok = np.isfinite(X).all(axis=1)
        X, ids = X[ok], ids[ok]
        n = X.shape[0]
        if use_pca and n > 3:
            k = min(10, X.shape[1], max(1, n - 1))
            if n > k:
                Xp = (X - np.mean(X, axis=1, keepdims=True)) / (np.std(X, axis=1, keepdims=True) + 1e-8)
                if X.shape[1] >= 5:
                    from scipy.signal import savgol_filter
                    Xp = savgol_filter(Xp, window_length=5, polyorder=2, deriv=1, axis=1)
                keep, _, _ = mahalanobis_filter(Xp, k, outlier_percentile, faithful_solver=faithful_solver)
                X, ids = X[keep], ids[keep]
        uniq = [int(i) for i in np.unique(ids)]
        if not uniq:
            continue
        if len(uniq) >= 3:
            cal, tmp = train_test_split(uniq, test_size=1.0 - cal_frac, random_state=random_state)
            rel = test_frac / (val_frac + test_frac) if (val_frac + test_frac) > 0 else 0.5
            val, tst = train_test_split(tmp, test_size=rel, random_state=random_state)
        else:
            cal, val, tst = [uniq[0]], [], ([uniq[1]] if len(uniq) == 2 else [])
        for name, sel in (("cal", cal), ("val", val), ("test", tst)):
            if sel:
                splits[nut][name] = np.vstack([X[ids == i] for i in sel])
    tests = [(nut, splits[nut]["test"]) for nut in nut_types if len(splits[nut]["test"])]
    Xts = np.vstack([t for _, t in tests]) if tests else empty()
    yts = (np.concatenate([np.full(len(t), int(nut != target_nut)) for nut, t in tests]) if tests
           else np.array([], dtype=int))
    out = [t for nut, t in tests if nut != target_nut]
    tg = splits[target_nut]
    return splits, Xts, yts, tg["cal"], tg["val"], tg["test"], (np.vstack(out) if out else empty())
