"""Critical limits of a SIMCA model: O(1) host arithmetic per model.

These need special functions (F / chi-square quantiles with non-integer dof, erfinv) and
run once per fitted model on scalars that the device produced: (n, A), the eigenvalue
spectrum, and -- only for 'perc' / 'chi2pom' -- the training T2 / Q vectors.  The bulk
work (T2, Q of every row) never comes through here.

Behaviour follows /root/reference/utils/SIMCA.py:156-236 (cited per function).
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import lru_cache
from typing import Optional

import numpy as np
from scipy import stats
from scipy.special import erfinv

RULES = {"sim": 0, "alt": 1, "ci": 2, "dd": 3}


@lru_cache(maxsize=4096)
def _f_ppf(cl, d1, d2):
    return float(stats.f.ppf(cl, d1, d2))


@lru_cache(maxsize=4096)
def _chi2_ppf(cl, dof):
    return float(stats.chi2.ppf(cl, dof))


T2_KINDS = ("perc", "Fdistrig", "Fdist", "chi2", "chi2pom")
Q_KINDS = ("perc", "jm", "chi2box", "chi2pom")


@dataclass
class Limits:
    T2_limit: float
    Q_limit: float
    D_limit: float
    t2dof: Optional[int] = None
    t2scale: Optional[float] = None
    qdof: Optional[int] = None
    qscale: Optional[float] = None

    def reduced_scales(self, rule: str):
        """(ct, cq) such that the reduced distances are T2*ct and Q*cq (SIMCA.py:131-144)."""
        if rule == "dd":
            return self.t2dof / self.t2scale, self.qdof / self.qscale
        return 1.0 / self.T2_limit, 1.0 / self.Q_limit


class Spectrum:
    """Eigenvalue spectrum known as its k leading values plus the power sums of everything beyond them
    (what a partial eigensolve + traces of the deflated covariance give: engine.eig_topk_subspace)."""

    def __init__(self, top, tail):
        self.top = np.asarray(top, dtype=np.float64)
        self.tail = tuple(float(t) for t in tail)       # (sum, sum of squares, sum of cubes) of lambda_{k+1..}
        self._sfx = None

    def suffix_sums(self):
        """sfx[i][A] = sum_{k >= A} top[k]^(i+1), A = 0..len(top): the tail moments of every nested model in O(1)
        (the CV grid asks for all A of every fold)."""
        if self._sfx is None:
            z = np.zeros(1)
            self._sfx = tuple(np.concatenate([np.cumsum((self.top ** k)[::-1])[::-1], z]) for k in (1, 2, 3))
        return self._sfx

    def __getitem__(self, sl):
        return self.top[sl]

    def __len__(self):
        return len(self.top)


class LazySpectrum(Spectrum):
    """A Spectrum that can still produce the full eigenvalue array (descending, what the reference keeps under
    `_model[cls]['eigs_all']`) when somebody asks for it: the fit only needs the leading values and the tail sums, and
    a full 1050 x 1050 fp64 eigensolve is ~13 ms on the GPU -- 40 % of the whole CV grid when done for the refit."""

    def __init__(self, top, tail, full_fn):
        super().__init__(top, tail)
        self._full_fn, self._full = full_fn, None

    def full(self):
        if self._full is None:
            self._full = np.asarray(self._full_fn(), dtype=np.float64)
            self._full_fn = None
        return self._full

    def __array__(self, dtype=None, copy=None):
        a = self.full()
        return a.astype(dtype) if dtype is not None else a


def tail_moments(eigs_all, A):
    """theta_i = sum_{k>A} lambda_k^i, i=1..3 (SIMCA.py:189-191)."""
    if isinstance(eigs_all, Spectrum):
        if A > len(eigs_all.top):
            raise ValueError("partial spectrum shorter than n_components")
        s1, s2, s3 = eigs_all.suffix_sums()
        t1, t2, t3 = eigs_all.tail
        return float(t1 + s1[A]), float(t2 + s2[A]), float(t3 + s3[A])
    tail = np.asarray(eigs_all, dtype=np.float64)[A:]
    return float(tail.sum()), float((tail ** 2).sum()), float((tail ** 3).sum())


class DeviceSample:
    """Training distances that stay on the GPU (bulk fits): the order statistics / moments the 'perc' and 'chi2pom'
    limits need are reduced there (fp64), only scalars come back.  np.percentile's default (linear interpolation
    between order statistics) is restated on the sorted device array."""

    def __init__(self, values):
        self.v = values                      # 1-D CUDA tensor (fp32 distances)

    def __len__(self):
        return int(self.v.numel())

    def percentile(self, q):
        import torch
        srt = torch.sort(self.v.double()).values
        pos = (q / 100.0) * (srt.numel() - 1)
        lo = int(np.floor(pos)); hi = min(lo + 1, srt.numel() - 1)
        a, b = float(srt[lo]), float(srt[hi])
        return a + (b - a) * (pos - lo)

    def mean(self):
        return float(self.v.double().mean())

    def var1(self):
        return float(self.v.double().var(unbiased=True)) if self.v.numel() > 1 else 0.0


def _percentile(x, q):
    return x.percentile(q) if isinstance(x, DeviceSample) else float(np.percentile(x, q))


def _mean(x):
    return x.mean() if isinstance(x, DeviceSample) else float(np.mean(x))


def _var1(x):
    return x.var1() if isinstance(x, DeviceSample) else (float(np.var(x, ddof=1)) if len(x) > 1 else 0.0)


def hotelling_limit(kind, cl, A, n, T2=None):
    """SIMCA.py:156-182.  Returns (limit, dof, scale); dof/scale only for 'chi2pom'."""
    if kind == "Fdist":
        return A * (n - 1) / (n - A) * _f_ppf(cl, A, n - A), None, None
    if kind == "Fdistrig":
        return (A / n) * (n ** 2 - 1) / (n - A) * _f_ppf(cl, A, n - A), None, None
    if kind == "chi2":
        return _chi2_ppf(cl, A), None, None
    if kind == "perc":
        return _percentile(T2, cl * 100), None, None
    if kind == "chi2pom":
        h0 = _mean(T2)
        var = _var1(T2)
        dof = max(int(np.round(2 * h0 * h0 / var)) if var > 0 else 1, 1)
        return h0 * _chi2_ppf(cl, dof) / dof, dof, h0
    raise ValueError(f"unknown t2lim {kind!r}")


def residual_limit(kind, cl, thetas, Q=None, verbose=False):
    """SIMCA.py:184-217.  Returns (limit, dof, scale)."""
    th1, th2, th3 = thetas
    if kind == "jm":
        # Jackson-Mudholkar
        h0 = max(1.0 - (2.0 * th1 * th3) / (3.0 * th2 * th2), 0.001)
        ca = np.sqrt(2.0) * erfinv(2.0 * cl - 1.0)
        term = ca * np.sqrt(2.0 * th2 * h0 * h0) / th1 + 1.0 + th2 * h0 * (h0 - 1.0) / (th1 * th1)
        return float(th1 * term ** (1.0 / h0)), None, None
    if kind == "chi2box":
        g, Ng = th2 / th1, th1 * th1 / th2
        if verbose:  # the reference prints these unconditionally (SIMCA.py:207-208)
            print("here we are")
            print(th1, th2, g, Ng)
        return float(g * stats.chi2.ppf(cl, Ng)), None, None
    if kind == "perc":
        return _percentile(Q, cl * 100), None, None
    if kind == "chi2pom":
        v0 = _mean(Q)
        dof = max(round(2 * v0 * v0 / _var1(Q)), 1)
        return v0 * _chi2_ppf(cl, dof) / dof, dof, v0
    raise ValueError(f"unknown qlim {kind!r}")


def distance_limit(rule, dcl, A, t2lim, qlim, thetas, t2dof=None, qdof=None):
    """SIMCA.py:219-236."""
    if rule == "sim":
        return 1
    if rule == "alt":
        return float(np.sqrt(2))
    if rule == "ci":
        th1, th2, _ = thetas
        tr1 = A / t2lim + th1 / qlim
        tr2 = A / t2lim ** 2 + th2 / qlim ** 2
        return float((tr2 / tr1) * stats.chi2.ppf(dcl, tr1 * tr1 / tr2))
    if rule == "dd":
        return float(stats.chi2.ppf(dcl, t2dof + qdof))
    raise ValueError(f"unknown type {rule!r}")


def needs_training_scores(t2_kind, q_kind):
    return t2_kind in ("perc", "chi2pom") or q_kind in ("perc", "chi2pom")


def all_limits(rule, t2_kind, t2cl, q_kind, qcl, dcl, A, n, eigs_all, T2=None, Q=None,
               verbose=False) -> Limits:
    th = tail_moments(eigs_all, A)
    tl, tdof, tsc = hotelling_limit(t2_kind, t2cl, A, n, T2)
    ql, qdof, qsc = residual_limit(q_kind, qcl, th, Q, verbose=verbose)
    dl = distance_limit(rule, dcl, A, tl, ql, th, tdof, qdof)
    return Limits(float(tl), float(ql), dl, tdof, tsc, qdof, qsc)
