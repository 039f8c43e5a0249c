"""Cross-validated SIMCA grid -- drop-in for /root/reference/utils/CVSIMCA.py.

`cross_validate_simca_grid` keeps the reference's signature, result dictionary and
aggregation (spec = mean over folds, sens = pooled over all rows, eff = sqrt(sens*spec),
argmax, final refit; CVSIMCA.py:103-269).  Two execution paths:

* batched (the B200 path): when the estimator is this package's
  SIMCA itself with a single model class, every fold is fitted ONCE -- Gram of the fold's
  training rows through a gather index, one eigensolve per fold -- and ALL nested models
  A = 1..A_max are scored in one multi-A kernel launch per fold, because PCA loadings are
  nested (the reference re-runs a full SVD plus a randomized SVD for every (fold, A) cell);
  limit-only parameter combinations (type / t2lim / qlim / confidence levels) reuse the
  same decompositions.  Decisions and confusion counts stay on the device until the final
  tiny table comes back.
* generic: anything else (a Pipeline with preprocessing steps, several model classes)
  walks the reference's clone/fit/predict loop, each fit still running on the GPU.
"""
from __future__ import annotations

import numpy as np
from sklearn.base import clone
from sklearn.model_selection import BaseCrossValidator, KFold, ParameterGrid
from sklearn.pipeline import Pipeline

from .. import engine, limits
from .SIMCA import SIMCA

_LIMIT_PARAMS = {"type", "t2lim", "t2cl", "qlim", "qcl", "dcl", "verbose", "maxPC", "criteria"}


class ClasswiseKFoldWithExternalVal(BaseCrossValidator):
    """One-class cross-validation splitter with the reference's protocol (CVSIMCA.py:39-80): only the rows of the TARGET
    class are divided into `n_splits` folds; a training set is the target rows outside the fold, a test set is the fold
    plus EVERY row of the other classes (so the other classes are tested in each split).  The target rows are given
    either as row indices (`cls_idx`) or as a label to look up in `y` (`cls_label`; a scalar `cls_idx` is read as a label
    too)."""

    def __init__(self, n_splits=5, cls_idx=None, cls_label=None, shuffle=False, random_state=None):
        self.kf = KFold(n_splits=n_splits, shuffle=shuffle, random_state=random_state)     # attribute kept: callers read it
        self.cls_idx = cls_idx if cls_idx is None else np.asarray(cls_idx)
        self.cls_label = cls_label

    def get_n_splits(self, X=None, y=None, groups=None):
        return self.kf.get_n_splits()

    def _target_rows(self, y):
        """Row indices of the target class, validated against the number of folds."""
        label = None
        if self.cls_idx is not None and self.cls_idx.ndim > 0:
            rows = self.cls_idx
        else:
            label = self.cls_label if self.cls_idx is None else int(self.cls_idx)
            if label is None:
                rows = np.zeros(0, dtype=np.int64)
            elif y is None:
                raise ValueError("the target class is given as a label: split(X, y) needs y to find its rows.")
            else:
                rows = np.flatnonzero(np.asarray(y) == label)
        if rows.size == 0:
            raise ValueError("cls_idx is empty: no sample of the target class found.")
        if rows.size < self.kf.n_splits:
            raise ValueError(f"Too many splits ({self.kf.n_splits}) for {rows.size} target-class samples.")
        return rows

    def split(self, X, y=None, groups=None):
        rows = self._target_rows(y)
        outside = np.ones(X.shape[0], dtype=bool)
        outside[rows] = False
        others = np.flatnonzero(outside)                      # sorted, like the reference's setdiff1d
        for keep, held in self.kf.split(rows):
            yield rows[keep], np.concatenate((rows[held], others))


def _simca_steps(estimator):
    """[(parameter prefix, object)] of everything in `estimator` that speaks the SIMCA metric protocol
    (`_metrics_simca_conformity`, CVSIMCA.py:85-98): the estimator itself, or the matching steps of a Pipeline in order."""
    speaks = lambda obj: hasattr(obj, "_metrics_simca_conformity")
    if speaks(estimator):
        return [("", estimator)]
    if isinstance(estimator, Pipeline):
        return [(name + "__", step) for name, step in estimator.named_steps.items() if speaks(step)]
    return []


def _get_simca(estimator):
    """The SIMCA object inside `estimator` (the LAST matching step of a Pipeline)."""
    found = _simca_steps(estimator)
    if not found:
        raise AttributeError("no SIMCA object found in the estimator.")
    return found[-1][1]


def _find_ncomp_param_name(estimator):
    """Parameter path of n_components: '<step>__n_components' for the first SIMCA step of a Pipeline, else 'n_components'."""
    if not isinstance(estimator, Pipeline):
        return "n_components"
    found = _simca_steps(estimator)
    if not found:
        raise AttributeError("Pipeline without a SIMCA step: cannot locate n_components.")
    return found[0][0] + "n_components"


def _percent_metrics(TP, TN, FP, FN):
    with np.errstate(divide="ignore", invalid="ignore"):
        sens = np.float64(TP) / (TP + FN) * 100
        spec = np.float64(TN) / (TN + FP) * 100
    return sens, spec


# ----------------------------------------------------------------------------------------------
# batched path
# ----------------------------------------------------------------------------------------------
def _batched_eligible(estimator, param_grid):
    if not isinstance(estimator, SIMCA):
        return False
    mc = estimator.model_class
    if mc is None or (isinstance(mc, (list, tuple, np.ndarray)) and len(mc) != 1):
        return False
    nc = estimator.n_components
    if isinstance(nc, (list, tuple)) and len(nc) != 1:
        return False
    return all(k in _LIMIT_PARAMS or k == "n_components" for k in param_grid)


class _FoldModels:
    """Per-fold decompositions at A_max (one Gram + one eigensolve per fold) and the
    device-resident test rows; training-row scores only when a limit needs them."""

    def __init__(self, Xd, y, splits, cls, A_max, need_train_scores):
        torch = engine._torch()
        self.folds = []
        self.pool = None
        if not splits:                      # a rank without folds (more ranks than folds)
            return
        row_sets = []
        for tr, te in splits:
            tr = np.asarray(tr)
            rows = tr[y[tr] == cls]
            if A_max > min(int(rows.size) - 1, Xd.shape[1]):
                raise ValueError(f"n_components={A_max} too large for a training fold of {int(rows.size)} rows")
            row_sets.append(engine.index_tensor(rows, Xd.device))
        pcas = engine.fit_pca_folds(Xd, row_sets, A_max)       # one batched partial eigensolve for all folds
        self.pool = engine.fit_pca_folds.last_pool             # PCA of all training rows together (for the refit), or None
        # test rows and target flags of all folds in ONE upload each
        tes = [np.asarray(te) for _, te in splits]
        offs = np.concatenate([[0], np.cumsum([t.size for t in tes])])
        te_all = engine.index_tensor(np.concatenate(tes) if tes else np.zeros(0, np.int64), Xd.device)
        tgt_all = torch.as_tensor((y[np.concatenate(tes)] == cls).astype(np.uint8) if tes else np.zeros(0, np.uint8), device=Xd.device)
        for k_f, ((tr, te), idx, pca) in enumerate(zip(splits, row_sets, pcas)):
            te = tes[k_f]
            n = int(idx.numel())
            inv_eig = getattr(pca, "inv_eig_device", None)
            if inv_eig is None:
                inv_eig = torch.as_tensor(1.0 / pca.eigs_all[:A_max], device=Xd.device)
            te_idx = te_all[int(offs[k_f]):int(offs[k_f + 1])]
            f = dict(pca=pca, n=n, te=te, te_idx=te_idx, Xte=Xd.index_select(0, te_idx), inv_eig=inv_eig,
                     is_target=tgt_all[int(offs[k_f]):int(offs[k_f + 1])])
            if need_train_scores:
                T2tr, Qtr, _ = engine.score_multiA(Xd.index_select(0, idx), pca.mean, pca.P, inv_eig)
                f["T2tr"] = T2tr.double().cpu().numpy()
                f["Qtr"] = Qtr.double().cpu().numpy()
            self.folds.append(f)


def _run_batched(estimator, X, y, splits, combos, lv_of_combo, ncomp_in_grid, store_predictions, group=None):
    """`group`: torch.distributed process group (or None).  With more than one rank the folds are dealt round-robin
    (SURVEY.md §8(e)): every rank holds X, fits and scores only its folds, and the per-(fold, A) confusion counts and
    the pooled decisions are summed over the ranks, so every rank returns the same table."""
    torch = engine._torch()
    rank, world = 0, 1
    if group is not None:
        import torch.distributed as tdist
        rank, world = tdist.get_rank(group), tdist.get_world_size(group)
    n_folds = len(splits)
    mine = [i for i in range(n_folds) if i % world == rank]
    splits = [splits[i] for i in mine]
    base = clone(estimator)
    cls = base.model_class[0] if isinstance(base.model_class, (list, tuple, np.ndarray)) else base.model_class
    Xd = engine.to_device(X)
    A_max = max(lv for lvs in lv_of_combo for lv in lvs)

    def eff_params(combo):
        prm = {k: getattr(base, k) for k in ("type", "t2lim", "t2cl", "qlim", "qcl", "dcl")}
        prm.update({k: v for k, v in combo.items() if k in prm})
        if prm["type"] == "dd":
            prm["t2lim"] = prm["qlim"] = "chi2pom"
        return prm

    need_train = any(limits.needs_training_scores(eff_params(c)["t2lim"], eff_params(c)["qlim"]) for c in combos)
    fm = _FoldModels(Xd, y, splits, cls, A_max, need_train)
    n_samples = X.shape[0]
    tgt_all = torch.as_tensor((y == cls).astype(np.uint8), device=Xd.device)
    records, by_combo = [], []
    for combo, lvs in zip(combos, lv_of_combo):
        prm = eff_params(combo)
        cols = torch.as_tensor(np.asarray(lvs) - 1, device=Xd.device)
        pooled = torch.zeros((n_samples, len(lvs)), dtype=torch.uint8, device=Xd.device)
        writer = torch.zeros((n_samples, len(lvs)), dtype=torch.int32, device=Xd.device) if world > 1 else None
        cnt_folds = torch.zeros((n_folds, len(lvs), 4), dtype=torch.int64, device=Xd.device)
        # limits of every requested nested model of every fold (host scalars), uploaded in ONE copy; then one launch
        # per fold scores its test rows and decides for all A at once
        lim_h = np.zeros((max(len(fm.folds), 1), 3, A_max), np.float32)
        for k_f, f in enumerate(fm.folds):
            for A in lvs:
                lim = limits.all_limits(prm["type"], prm["t2lim"], prm["t2cl"], prm["qlim"], prm["qcl"], prm["dcl"],
                                        A, f["n"], f["pca"].eigs_all,
                                        f["T2tr"][:, A - 1] if need_train else None,
                                        f["Qtr"][:, A - 1] if need_train else None)
                lim_h[k_f, 0, A - 1], lim_h[k_f, 1, A - 1] = lim.reduced_scales(prm["type"])
                lim_h[k_f, 2, A - 1] = lim.D_limit
        lim_d = torch.as_tensor(lim_h, device=Xd.device)
        for k_f, (fi, f) in enumerate(zip(mine, fm.folds)):
            _, _, acc = engine.score_multiA(f["Xte"], f["pca"].mean, getattr(f["pca"], "P32", f["pca"].P),
                                            getattr(f["pca"], "inv_eig32", f["inv_eig"]), rule=prm["type"],
                                            ct=lim_d[k_f, 0], cq=lim_d[k_f, 1], dlim=lim_d[k_f, 2])
            acc = acc.index_select(1, cols).contiguous()
            pooled[f["te_idx"]] = acc
            if writer is not None:
                writer[f["te_idx"]] = acc.to(torch.int32) + 2 * (fi + 1)
            cnt_folds[fi] = engine.conformity_counts(acc, f["is_target"])
        if world > 1:
            # A row can sit in the test set of SEVERAL folds (ClasswiseKFoldWithExternalVal puts every non-target row
            # in all of them); the reference keeps the decision of the LAST fold that tested it (pred_vec[te] = y_pred,
            # CVSIMCA.py:186).  Every rank tags its decisions with 2 * (fold + 1); the maximum over the ranks is the
            # last writer, its low bit the decision (0 = never tested).
            tdist.all_reduce(writer, op=tdist.ReduceOp.MAX, group=group)
            tdist.all_reduce(cnt_folds, group=group)
            pooled = (writer & 1).to(torch.uint8)
        cnt = cnt_folds.cpu().numpy()
        spec_folds = np.stack([_percent_metrics(cnt[fi, :, 0], cnt[fi, :, 1], cnt[fi, :, 2], cnt[fi, :, 3])[1]
                               for fi in range(n_folds)])
        cnt_all = engine.conformity_counts(pooled, tgt_all).cpu().numpy()
        sens_all, _ = _percent_metrics(cnt_all[:, 0], cnt_all[:, 1], cnt_all[:, 2], cnt_all[:, 3])
        pooled_h = pooled.cpu().numpy().astype(float) if store_predictions else None
        for j, A in enumerate(lvs):
            spec = float(np.mean(spec_folds[:, j]))
            sens = float(sens_all[j])
            rec = {"params": combo.copy(), "LV": (combo.get("n_components") if ncomp_in_grid else A),
                   "spec": spec, "sens": sens, "eff": float(np.sqrt(sens * spec))}
            records.append(rec)
            if store_predictions:
                by_combo.append({"params": combo.copy(), "LV": rec["LV"], "prediction": pooled_h[:, j].copy()})
    _run_batched.last_prefit = None if fm.pool is None else dict(X_host=X, Xd=Xd, rows=fm.pool["rows"], model=fm.pool["model"])
    return records, by_combo


# ----------------------------------------------------------------------------------------------
# public entry point
# ----------------------------------------------------------------------------------------------
def cross_validate_simca_grid(estimator, X, y, cv, LV_min=2, LV_max=10, param_grid=None, refit_metric="eff",
                              class_index=None, print_summary=True, store_predictions=False, batched="auto",
                              group=None):
    """Grid x n-components x folds cross-validation of a SIMCA estimator (CVSIMCA.py:103-269).

    Returns {'results': [{params, LV, spec, sens, eff}], 'best_params', 'best_LV', 'best_score',
    'best_estimator'[, 'by_combo']}.  Two arguments the reference does not have: `batched`
    ('auto' | True | False) selects the execution path; `group` (a torch.distributed process group, batched path
    only) deals the folds over the ranks -- every rank must make the same call and gets the same result."""
    if param_grid is None:
        param_grid = {}
    y = np.asarray(y)
    base_est = clone(estimator)
    ncomp_key = _find_ncomp_param_name(base_est)
    grid_includes_ncomp = any(k.endswith("n_components") for k in param_grid.keys())
    lv_values = None if grid_includes_ncomp else list(range(LV_min, LV_max + 1))

    use_batched = batched is True or (batched == "auto" and _batched_eligible(base_est, param_grid))
    if use_batched and not _batched_eligible(base_est, param_grid):
        raise ValueError("batched=True needs a bare SIMCA estimator with one model class and a limit-only grid")
    cls_of = lambda s: s.model_class[0] if isinstance(s.model_class, (list, tuple, np.ndarray)) else s.model_class
    if use_batched and class_index is not None and class_index != cls_of(base_est):
        use_batched = False   # metrics against a class other than the modelled one: generic path

    results_records, by_combo_predictions = [], []
    if use_batched:
        splits = list(cv.split(X, y))
        combos_full = list(ParameterGrid(param_grid))
        if grid_includes_ncomp:
            # group combos that differ only in n_components so they share decompositions
            combos, lv_of_combo = [], []
            for c in combos_full:
                combos.append(c); lv_of_combo.append([int(c["n_components"])])
        else:
            combos, lv_of_combo = combos_full, [lv_values] * len(combos_full)
        results_records, by_combo_predictions = _run_batched(base_est, X, y, splits, combos, lv_of_combo,
                                                             grid_includes_ncomp, store_predictions, group)
    else:
        for combo in ParameterGrid(param_grid):
            for lv in ([None] if grid_includes_ncomp else lv_values):
                est_lv = clone(base_est)
                est_lv.set_params(**combo)
                if not grid_includes_ncomp:
                    est_lv.set_params(**{ncomp_key: lv})
                n_folds = cv.get_n_splits(X, y)
                pred_vec = np.zeros(X.shape[0], dtype=float)
                step_spec = np.zeros(n_folds, dtype=float)
                last = None
                for i, (tr, te) in enumerate(cv.split(X, y)):
                    est_fold = clone(est_lv)
                    est_fold.fit(X[tr, :], y[tr])
                    try:
                        y_pred = est_fold.predict(X[te, :])
                    except TypeError:
                        y_pred = est_fold.predict(X[te, :], y[te])
                    y_pred = np.ravel(y_pred)
                    pred_vec[te] = y_pred
                    last = _get_simca(est_fold)
                    ci = class_index if class_index is not None else getattr(last, "model_class", 1)
                    with np.errstate(divide="ignore", invalid="ignore"):
                        step_spec[i] = last._metrics_simca_conformity(y_true=y[te], y_pred=y_pred, class_index=ci)['specificity']
                spec = float(np.mean(step_spec))
                ci = class_index if class_index is not None else getattr(last, "model_class", 1)
                with np.errstate(divide="ignore", invalid="ignore"):
                    sens = float(last._metrics_simca_conformity(y_true=y, y_pred=pred_vec, class_index=ci)['sensitivity'])
                rec = {"params": combo.copy(), "LV": (combo.get(ncomp_key) if grid_includes_ncomp else lv),
                       "spec": spec, "sens": sens, "eff": float(np.sqrt(sens * spec))}
                results_records.append(rec)
                if store_predictions:
                    by_combo_predictions.append({"params": combo.copy(), "LV": rec["LV"], "prediction": pred_vec.copy()})

    metric_key = {"eff": "eff", "spec": "spec", "sens": "sens"}[refit_metric]
    best_idx = int(np.argmax([r[metric_key] for r in results_records]))
    best_score = results_records[best_idx][metric_key]
    best_params = results_records[best_idx]["params"].copy()
    best_LV = results_records[best_idx]["LV"]

    if print_summary:
        pstr = lambda p: ", ".join(f"{k}={v}" for k, v in sorted(p.items()))
        current = None
        for r in sorted(results_records, key=lambda r: (pstr(r["params"]), r["LV"])):
            if pstr(r["params"]) != current:
                current = pstr(r["params"])
                print("\nPARAMS:", current)
            print(f"  LV={r['LV']:>2} | SPEC={r['spec']:.4f} | SENS={r['sens']:.4f} | EFF={r['eff']:.4f}")
        print(f"\n[best @ {refit_metric}] LV={best_LV} | score={best_score:.4f} | params={best_params}")

    best_estimator = clone(estimator)
    best_estimator.set_params(**best_params)
    if not grid_includes_ncomp:
        best_estimator.set_params(**{_find_ncomp_param_name(best_estimator): best_LV})
    if use_batched and getattr(_run_batched, "last_prefit", None) is not None:
        # the batched path has decomposed the pooled training rows already (same batched eigensolve as the folds)
        best_estimator._prefit = _run_batched.last_prefit
        _run_batched.last_prefit = None
    best_estimator.fit(X, y)

    out = {"results": results_records, "best_params": best_params, "best_LV": best_LV,
           "best_score": best_score, "best_estimator": best_estimator}
    if store_predictions:
        out["by_combo"] = by_combo_predictions
    return out


def plot_cv(res, metric="eff", params=None, show_best=True, title=None):
    """CV curve of one parameter combination (CVSIMCA.py:274-331); needs matplotlib."""
    import matplotlib.pyplot as plt
    if params is None:
        params = res.get("best_params") or {}
    rows = [r for r in res["results"] if all(r["params"].get(k, object()) == v for k, v in params.items())]
    if not rows:
        raise ValueError("no record matches the requested parameters.")
    rows.sort(key=lambda r: r["LV"])
    lv = np.array([r["LV"] for r in rows]); val = np.array([r[metric] for r in rows])
    plt.figure(figsize=(8, 5))
    plt.plot(lv, val, marker="o", color="C0", label=f"Mean CV {metric.upper()}")
    if show_best and "best_LV" in res:
        plt.axvline(res["best_LV"], color="r", linestyle="--",
                    label=f"Best LV = {res['best_LV']} ({metric} = {res['best_score']:.3f})")
    plt.xlabel("Number of latent variables (LVs)"); plt.ylabel(metric.upper())
    plt.title(title or f"Cross-validation {metric.upper()} vs LV")
    plt.grid(True, linestyle="--", alpha=0.5); plt.legend(); plt.show()
