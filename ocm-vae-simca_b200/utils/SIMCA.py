"""SIMCA / DD-SIMCA one-class estimator -- drop-in for /root/reference/utils/SIMCA.py.

Same constructor, same `fit / predict / transform / score / _metrics_simca_conformity`,
same `_model[cls]` dictionary keys, same prints, same sklearn estimator behaviour
(`clone`, `get_params/set_params`, last step of a `Pipeline`).  What differs is where the
arithmetic runs: the PCA fit is a fp64 Gram accumulation + symmetric eigensolve on the
GPU and predict/transform is ONE fused CUDA kernel per call that reads each spectrum once
and writes back only T2, Q and the decision (engine.py -> include/ocm_simca.h).

Deliberate deviations from the reference (each documented in DESIGN.md):
  * n_components == 1 works (the reference crashes in np.cov, SIMCA.py:69);
  * the predict-time decomposition IS the fit-time one (the reference refits a second,
    randomized-SVD PCA, SIMCA.py:75 -- identical up to its own run-to-run noise);
  * `invcovT` is exactly diag(1/lambda_a) (the reference's pinv(cov(T)) equals it to 1e-15);
  * CUDA tensors are accepted and, for them, results stay on the device.
"""
from __future__ import annotations

import numpy as np
from sklearn.base import BaseEstimator, ClassifierMixin

from .. import engine, limits


class _DevicePCA:
    """What `_model[cls]["pca_model"]` exposes: the sklearn-PCA attributes the reference's
    callers touch (`components_`, `mean_`, `explained_variance_`, `n_components_`) plus
    transform / inverse_transform for plotting code.  Host copies are made lazily."""

    def __init__(self, pca: engine.PCAModel):
        self._pca = pca
        self.n_components_ = int(pca.P.shape[0])
        self.n_samples_ = pca.n_samples
        self.explained_variance_ = pca.eigs_all[: self.n_components_].copy()
        self._P = self._mean = None

    @property
    def components_(self):
        if self._P is None:
            self._P = self._pca.P.cpu().numpy()
        return self._P

    @property
    def mean_(self):
        if self._mean is None:
            self._mean = self._pca.mean.cpu().numpy()
        return self._mean

    def transform(self, X):
        return (np.asarray(X) - self.mean_) @ self.components_.T

    def inverse_transform(self, T):
        return np.asarray(T) @ self.components_ + self.mean_


class _ModelDict(dict):
    """`_model[cls]` of the reference (SIMCA.py:83-99) with lazy entries.  'eigs_all' holds the leading eigenvalues and
    the tail power sums the limits were computed from and turns into the full eigenvalue array (one extra symmetric
    eigenvalue solve) the first time it is read; after a BULK fit the per-row arrays 'T', 'T2', 'Q', 'T2red', 'Qred'
    are thunks too (the distances stay on the device, the scores are not computed at all until somebody asks)."""

    def __init__(self, *a, lazy=None, **k):
        super().__init__(*a, **k)
        self._lazy = dict(lazy or {})

    def _materialise(self, key=None):
        keys = list(self._lazy) if key is None else ([key] if key in self._lazy else [])
        for kk in keys:
            dict.__setitem__(self, kk, self._lazy.pop(kk)())
        if key in (None, "eigs_all"):
            v = dict.get(self, "eigs_all")
            if hasattr(v, "full"):
                dict.__setitem__(self, "eigs_all", v.full())

    def __getitem__(self, key):
        self._materialise(key)
        return dict.__getitem__(self, key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._lazy

    def get(self, key, default=None):
        self._materialise(key)
        return dict.get(self, key, default)

    def keys(self):
        return list(dict.keys(self)) + list(self._lazy)

    def items(self):
        self._materialise()
        return dict.items(self)

    def values(self):
        self._materialise()
        return dict.values(self)

    def copy(self):
        self._materialise()
        return _ModelDict(dict.copy(self))


BULK_FIT_BYTES = 256 << 20     # class matrices above this take the bulk fit path (no copies, lazy per-row outputs)


def _score_rows(Xd, idx, mean, P, Sinv, **kw):
    """T2, Q, accept of the rows `idx` of Xd (None = all rows) through the dispatcher's bulk kernels; a row subset is
    gathered in bounded chunks (<= 512 MB) instead of materialising the class matrix."""
    torch = engine._torch()
    if idx is None:
        return engine.score(Xd, mean, P, Sinv, **kw)
    n = int(idx.numel())
    chunk = max(65536, (512 << 20) // (Xd.shape[1] * Xd.element_size()))
    T2 = torch.empty(n, dtype=torch.float32, device=Xd.device)
    Q = torch.empty(n, dtype=torch.float32, device=Xd.device)
    acc = torch.empty(n, dtype=torch.uint8, device=Xd.device)
    for r0 in range(0, n, chunk):
        r1 = min(n, r0 + chunk)
        t2, q, a = engine.score(Xd.index_select(0, idx[r0:r1]), mean, P, Sinv, **kw)
        T2[r0:r1], Q[r0:r1], acc[r0:r1] = t2, q, a
    return T2, Q, acc


class SIMCA(BaseEstimator, ClassifierMixin):
    def __init__(self, n_components=2, model_class=None, type: str = 'alt', t2lim='Fdist', t2cl=0.95,
                 qlim='jm', qcl=0.95, dcl=0.95, maxPC=20, criteria='compl', verbose=True):
        self.n_components = n_components
        self.model_class = model_class
        self.type = type
        self.t2lim = t2lim
        self.t2cl = t2cl
        self.qlim = qlim
        self.qcl = qcl
        self.criteria = criteria
        self.dcl = dcl
        self.maxPC = maxPC
        self.metrics = {}
        self.verbose = verbose

    # ------------------------------------------------------------------ fit ------------
    def fit(self, X, classes):
        # argument normalisation mutates the constructor parameters exactly like the
        # reference does (SIMCA.py:29-48); CVSIMCA relies on clone() before each fit.
        classes_np = classes.detach().cpu().numpy() if hasattr(classes, "detach") else np.asarray(classes)
        if self.model_class is None:
            self.model_class = np.unique(classes_np)
        elif isinstance(self.model_class, (int, np.integer)):
            self.model_class = [self.model_class]
        if not isinstance(self.n_components, list):
            self.n_components = [self.n_components]
        if len(self.n_components) == 1:
            self.n_components = [self.n_components[0]] * len(self.model_class)
        elif len(self.n_components) != len(self.model_class):
            raise ValueError("n_components length must match number of classes")
        if self.type == 'dd' and self.t2lim != 'chi2pom':
            print('t2lim set as chi2pom')
            self.t2lim = 'chi2pom'
        if self.type == 'dd' and self.qlim != 'chi2pom':
            print('qlim set as chi2pom')
            self.qlim = 'chi2pom'

        # CV refit: cross_validate_simca_grid hands over the device copy of X and the decomposition of the pooled
        # training rows it already has (engine.fit_pca_folds.last_pool); used only for exactly this X and these rows
        pre = self.__dict__.pop("_prefit", None)
        Xd = pre["Xd"] if pre is not None and pre["X_host"] is X else engine.to_device(X)
        self._model = {}
        for i, cls in enumerate(self.model_class):
            rows = np.flatnonzero(classes_np == cls)
            self._model[cls] = self._fit_one_class(Xd, rows, int(self.n_components[i]),
                                                   pre if pre is not None and pre["X_host"] is X else None)
        self.n_features_in_ = Xd.shape[1]
        self.is_fitted_ = True
        return self

    def _fit_one_class(self, Xd, rows, A, pre=None):
        """Device restatement of SIMCA.py:62-99.  Small class matrices keep every `_model[cls]` entry eager (scores T
        included); above BULK_FIT_BYTES nothing is copied: the Gram runs through the gather index (or over all rows),
        the training distances come from the bulk scoring kernel and stay on the device, 'perc' / 'chi2pom' limits
        are reduced there, and the per-row entries of `_model[cls]` are produced on first access."""
        import weakref
        n_all = Xd.shape[0]
        all_rows = rows.size == n_all and (rows.size == 0 or (rows[0] == 0 and rows[-1] == n_all - 1 and np.all(np.diff(rows) == 1)))
        idx = None if all_rows else engine.index_tensor(rows, Xd.device)
        n = int(rows.size)
        if not 1 <= A <= min(n - 1, Xd.shape[1], 64):
            raise ValueError(f"n_components={A} must be in 1..min(n_class-1, n_features, 64)")
        pm = pre["model"] if pre is not None else None
        if pm is not None and pre["Xd"] is Xd and A <= pm.P.shape[0] and np.array_equal(rows, pre["rows"]):
            pca = engine.PCAModel(mean=pm.mean, P=pm.P[:A].contiguous(), eigs_all=pm.eigs_all, n_samples=pm.n_samples, x_rms=pm.x_rms)
        else:
            pca = engine.fit_pca(Xd, idx, A)
        lam = pca.eigs_all[:A]
        torch = engine._torch()
        inv_eig = torch.as_tensor(1.0 / lam, device=Xd.device)
        Sinv = torch.diag(inv_eig)
        bulk = n * Xd.shape[1] * Xd.element_size() > BULK_FIT_BYTES
        lazy = {}
        if not bulk:
            Xcls = Xd if idx is None else Xd.index_select(0, idx)
            T2, Q, _, T = engine.score(Xcls, pca.mean, pca.P, Sinv, want_scores=True)
            T2s = T2h = T2.double().cpu().numpy()
            Qs = Qh = Q.double().cpu().numpy()
            Th = T.double().cpu().numpy()
        else:
            T2, Q, _ = _score_rows(Xd, idx, pca.mean, pca.P, Sinv)
            T2s, Qs = limits.DeviceSample(T2), limits.DeviceSample(Q)
            T2h = Qh = Th = None
            xref = weakref.ref(Xd)

            def lazy_scores():
                X_ = xref()
                if X_ is None:
                    raise RuntimeError("the training matrix of this bulk fit has been released; '_model[cls][\"T\"]' "
                                       "can no longer be computed (read it while the data is alive, or refit)")
                out = np.empty((n, A), dtype=np.float64)
                chunk = max(65536, (256 << 20) // (X_.shape[1] * X_.element_size()))
                for r0 in range(0, n, chunk):
                    r1 = min(n, r0 + chunk)
                    blk = X_[r0:r1] if idx is None else X_.index_select(0, idx[r0:r1])
                    out[r0:r1] = engine.score(blk, pca.mean, pca.P, Sinv, want_scores=True)[3].double().cpu().numpy()
                return out
            lazy["T"] = lazy_scores
        lim = limits.all_limits(self.type, self.t2lim, self.t2cl, self.qlim, self.qcl, self.dcl, A, n,
                                pca.eigs_all, T2s, Qs, verbose=True)
        if self.type == 'dd':
            # the reference keeps these on the instance (last class wins, SIMCA.py:179-180,:215-216)
            self._t2dof, self._t2scfact = lim.t2dof, lim.t2scale
            self._qdof, self._qscfact = lim.qdof, lim.qscale
            t2s, qs = lim.t2dof / lim.t2scale, lim.qdof / lim.qscale
        else:
            if self.t2lim == 'chi2pom':
                self._t2dof, self._t2scfact = lim.t2dof, lim.t2scale
            if self.qlim == 'chi2pom':
                self._qdof, self._qscfact = lim.qdof, lim.qscale
            t2s, qs = 1.0 / lim.T2_limit, 1.0 / lim.Q_limit
        entries = {
            "pca_model": _DevicePCA(pca),
            "n_components": A,
            "xmean": pca.mean.cpu().numpy(),
            "invcovT": np.diag(1.0 / lam),
            "eigs_all": pca.eigs_all,
            "P": pca.P.cpu().numpy(),
            "T2_limit": lim.T2_limit,
            "Q_limit": lim.Q_limit,
            "D_limit": lim.D_limit,
            "n_samples": n,
            # device-side handles (not in the reference)
            "_pca": pca,
            "_Sinv": Sinv,
            "_limits": lim,
        }
        if not bulk:
            if self.type == 'dd':
                T2red, Qred = lim.t2dof * T2h / lim.t2scale, lim.qdof * Qh / lim.qscale
            else:
                T2red, Qred = T2h / lim.T2_limit, Qh / lim.Q_limit
            entries.update({"T": Th, "T2": T2h, "Q": Qh, "T2red": T2red, "Qred": Qred})
        else:
            entries.update({"_T2_device": T2, "_Q_device": Q})
            lazy["T2"] = lambda: T2.double().cpu().numpy()
            lazy["Q"] = lambda: Q.double().cpu().numpy()
            lazy["T2red"] = lambda: T2.double().cpu().numpy() * t2s
            lazy["Qred"] = lambda: Q.double().cpu().numpy() * qs
        return _ModelDict(entries, lazy=lazy)

    # ------------------------------------------------------------------ scoring ---------
    def _score_class(self, Xd, cls, snv=False):
        info = self._model[cls]
        lim = info["_limits"]
        ct, cq = lim.reduced_scales(self.type)
        return engine.score(Xd, info["_pca"].mean, info["_pca"].P, info["_Sinv"], rule=self.type,
                            ct=ct, cq=cq, dlim=lim.D_limit, snv=snv, x_rms=getattr(info["_pca"], "x_rms", 0.0))

    def _score_class_host(self, X, cls, snv=False):
        """Same as _score_class for a large host matrix: streamed over the link chunk by chunk (engine.score_host),
        results as numpy arrays."""
        info = self._model[cls]
        lim = info["_limits"]
        ct, cq = lim.reduced_scales(self.type)
        return engine.score_host(X, info["_pca"].mean, info["_pca"].P, info["_Sinv"], rule=self.type,
                                 ct=ct, cq=cq, dlim=lim.D_limit, snv=snv, x_rms=getattr(info["_pca"], "x_rms", 0.0))

    def _score_all_classes_host(self, X):
        """[(T2, Q, accept)] per model class for a large host matrix.  One class: the overlapped host-buffer call.
        Several classes: every chunk crosses the link ONCE (engine.to_device: pinned bounce buffers) and is scored
        against each class while it is on the device."""
        if len(self.model_class) == 1:
            return [self._score_class_host(X, self.model_class[0])]
        n = X.shape[0]
        outs = [(np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.uint8)) for _ in self.model_class]
        rows = max(65536, engine.HOST_STREAM_BYTES // max(1, X.strides[0]))
        rows = -(-n // -(-n // rows))                      # equal chunks, none below the bulk kernels' row gate
        for r0 in range(0, n, rows):
            Xd = engine.to_device(X[r0:r0 + rows])
            for k, cls in enumerate(self.model_class):
                T2, Q, acc = self._score_class(Xd, cls)
                for dst, src in zip(outs[k], (T2, Q, acc)):
                    dst[r0:r0 + rows] = src.cpu().numpy()
        return outs

    def transform(self, X):
        """(T2, T2red, Q, Qred) of the LAST model class, as the reference returns (SIMCA.py:101-117)."""
        if engine.host_streamable(X):
            cls = self.model_class[-1]                     # the reference's loop overwrites: only the last class comes back
            T2, Q, _ = self._score_class_host(X, cls)
            ct, cq = self._model[cls]["_limits"].reduced_scales(self.type)
            T2h, Qh = T2.astype(np.float64), Q.astype(np.float64)
            return T2h, T2h * ct, Qh, Qh * cq
        on_device = _is_cuda_tensor(X)
        Xd = engine.to_device(X)
        for cls in self.model_class:
            T2, Q, _ = self._score_class(Xd, cls)
            lim = self._model[cls]["_limits"]
            ct, cq = lim.reduced_scales(self.type)
            T2red, Qred = T2 * ct, Q * cq
        if on_device:
            return T2, T2red, Q, Qred
        f = lambda t: t.double().cpu().numpy()
        T2h, Qh = f(T2), f(Q)
        return T2h, T2h * ct, Qh, Qh * cq

    def predict(self, X, y_true=None):
        """0./1. acceptance per model class, shape [N, n_classes] (SIMCA.py:120-154)."""
        on_device = _is_cuda_tensor(X)
        if engine.host_streamable(X):
            # a large numpy matrix is streamed over the link (never resident as a whole); decisions come back per class
            predictions = out = np.empty((X.shape[0], len(self.model_class)), dtype=np.float64)
            for i, (_, _, acc) in enumerate(self._score_all_classes_host(X)):
                out[:, i] = acc
        else:
            Xd = engine.to_device(X)
            torch = engine._torch()
            out = torch.empty((Xd.shape[0], len(self.model_class)), dtype=torch.float64, device=Xd.device)
            for i, cls in enumerate(self.model_class):
                _, _, acc = self._score_class(Xd, cls)
                out[:, i] = acc
            predictions = out if on_device else out.cpu().numpy()
        if y_true is not None:
            yt = y_true.detach().cpu().numpy() if hasattr(y_true, "detach") else np.asarray(y_true)
            pr = out.cpu().numpy() if on_device else predictions
            for i, cls in enumerate(self.model_class):
                self.metrics[cls] = self._metrics_simca_conformity(yt, pr[:, i], cls)
                if self.verbose:
                    m = self.metrics[cls]
                    print(f"Sample class {cls} = {np.sum(yt == cls)}")
                    print(f"Confusion Matrix for class {cls}:\nTP: {m['TP']}, TN: {m['TN']}, FP: {m['FP']}, FN: {m['FN']}")
                    print(f"Class {cls} - Sensitivity: {m['sensitivity']}, Specificity: {m['specificity']:.4f}, "
                          f"Accuracy: {m['accuracy']:.4f}, Efficiency: {m['efficiency']:.4f}")
        return predictions

    # ------------------------------------------------------------------ metrics ---------
    def _metrics_simca_conformity(self, y_true, y_pred, class_index):
        """Confusion counts and percentages (SIMCA.py:238-266); integer arithmetic on the host
        (the on-device variant for the CV grid is engine.conformity_counts)."""
        y_true = np.asarray(y_true)
        y_pred = np.asarray(y_pred)
        true_class = (y_true == class_index).astype(int)
        TP = np.sum((y_pred == 1) & (true_class == 1))
        TN = np.sum((y_pred == 0) & (true_class == 0))
        FP = np.sum((y_pred == 1) & (true_class == 0))
        FN = np.sum((y_pred == 0) & (true_class == 1))
        sensitivity = TP / (TP + FN) * 100
        specificity = TN / (TN + FP) * 100
        accuracy = (TP + TN) / (TP + TN + FP + FN) * 100
        efficiency = np.sqrt(sensitivity * specificity)
        return {'sensitivity': sensitivity, 'specificity': specificity, 'accuracy': accuracy,
                'efficiency': efficiency, 'TP': TP, 'TN': TN, 'FP': FP, 'FN': FN}

    def score(self, X, y):
        """Specificity of the prediction, as in the reference (SIMCA.py:268-278).  The
        reference broadcasts (N,1) predictions against (N,) labels; here the single
        model class is compared column-wise, which is what that code intends."""
        y_pred = self.predict(X, y_true=y)
        y_pred = y_pred.cpu().numpy() if _is_cuda_tensor(y_pred) else y_pred
        cls = self.model_class[0] if len(self.model_class) == 1 else self.model_class
        return self._metrics_simca_conformity(np.asarray(y), y_pred[:, 0], cls)['specificity']

    # ------------------------------------------------------------------ plots -----------
    def toplotT2Q(self, X, y_test):
        """T2red-vs-Qred acceptance plot (SIMCA.py:280-307); needs matplotlib."""
        import matplotlib.pyplot as plt
        for cls in self._model:
            _, T2red, _, Qred = self.transform(X)
            Dlim = self._model[cls]['D_limit']
            a = np.arange(0, Dlim + 0.0001, 0.0001)
            plt.figure(figsize=(6, 6))
            sc = plt.scatter(T2red, Qred, c=y_test, cmap='viridis', s=40, edgecolor='k', linewidth=0.5, alpha=0.7)
            plt.plot(a, np.sqrt(np.maximum(Dlim ** 2 - a ** 2, 0)), 'b-', lw=2, label=f'class {cls} limit')
            plt.xlabel(r"$T^2_{red}$"); plt.ylabel(r"$Q_{red}$")
            plt.legend(*sc.legend_elements(), title="Class")
            plt.xlim(left=0); plt.ylim(bottom=0); plt.grid(True, alpha=0.3); plt.tight_layout()
            plt.show()
            return plt

    def toplotT2Q_iterative(self, X, y_test):
        """Plotly version of the acceptance plot (SIMCA.py:310-381); needs plotly."""
        import plotly.graph_objects as go
        figs = []
        labels = np.asarray(y_test).astype(str)
        for cls in self._model:
            _, T2red, _, Qred = self.transform(X)
            Dlim = float(self._model[cls]['D_limit'])
            a = np.linspace(0, Dlim, 1200)
            fig = go.Figure()
            for c in np.unique(labels):
                mk = labels == c
                fig.add_trace(go.Scatter(x=T2red[mk], y=Qred[mk], mode='markers', name=f"Class {c}"))
            fig.add_trace(go.Scatter(x=a, y=np.sqrt(np.maximum(Dlim ** 2 - a ** 2, 0.0)), mode='lines',
                                     name="Decision Limit"))
            fig.update_layout(width=600, height=600, xaxis_title="T2 red", yaxis_title="Q red")
            figs.append(fig)
        return figs[0] if len(figs) == 1 else figs


def _is_cuda_tensor(x):
    return hasattr(x, "is_cuda") and bool(x.is_cuda)
