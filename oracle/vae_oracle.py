"""CPU ORACLE for the VAE-SIMCA decision utilities -- TEST INFRASTRUCTURE, NOT PRODUCT.

Restates, with numpy on host arrays, the arithmetic of
  * `compute_q_h_f`            /root/reference/vae_model.py:162-181
  * `VAESIMCA.fit_thresholds`  /root/reference/VAE_SIMCA.py:231-279
  * `VAESIMCA._compute_*_limit`/root/reference/VAE_SIMCA.py:281-346
  * `VAESIMCA.predict`         /root/reference/VAE_SIMCA.py:349-382
The encoder/decoder stay what they are in the reference (a torch module); the oracle
takes latent means `mu`, round-trip latents `mu_hat`, spectra `x` and reconstructions
`x_rec` as arrays, so it restates only the decision arithmetic.

Parity pin: tests/golden/vae_qhf.npz and vae_simca.npz, produced by the live reference
code (oracle/make_golden.py).  Note the percentile-flavoured limits of VAESIMCA: its
'Fdist' / 'chi2' / 'chi2pom' variants use np.percentile where SIMCA uses a distribution
quantile (VAE_SIMCA.py:286-297) -- restated as is.
"""
from __future__ import annotations

import numpy as np
from scipy import special
from scipy.stats import chi2


def q_h_f(x, x_rec, z):
    """vae_model.py:162-181: q = squared residual, h = leverage of the batch-standardised
    latent, chi-square moment matching, f = h/h0*Nh + q/q0*Nq."""
    x = np.asarray(x, np.float64); x_rec = np.asarray(x_rec, np.float64); z = np.asarray(z, np.float64)
    q = ((x - x_rec) ** 2).sum(1)
    q0, sq = q.mean(), q.std(ddof=1)
    Nq = 2 * (q0 / sq) ** 2
    zs = (z - z.mean(0)) / (z.std(0, ddof=1) + 1e-12)   # torch.std default is unbiased
    U, _, _ = np.linalg.svd(zs, full_matrices=False)
    h = (U ** 2).sum(1)
    h0, sh = h.mean(), h.std(ddof=1)
    Nh = 2 * (h0 / sh) ** 2
    f = h / h0 * Nh + q / q0 * Nq
    return q, h, f, chi2.ppf(0.95, Nq), chi2.ppf(0.95, Nh), chi2.ppf(0.95, Nh + Nq)


def latent_stats(mu_cal, eps=1e-12):
    """VAE_SIMCA.py:245-248: mean, pinv(cov + eps I)."""
    d = mu_cal.shape[1]
    m = mu_cal.mean(0)
    Sinv = np.linalg.pinv(np.cov(mu_cal, rowvar=False) + np.eye(d) * eps)
    return m, Sinv


def latent_t2(mu, m, Sinv):
    diff = mu - m[None, :]
    return np.einsum("ij,jk,ik->i", diff, Sinv, diff)


def t2_limit(T2, d, kind, cl):
    """VAE_SIMCA.py:281-300."""
    n = len(T2)
    dof = scale = None
    pc = np.percentile(T2, cl * 100)
    if kind in ("perc", "chi2"):
        lim = pc
    elif kind == "Fdist":
        lim = d * (n - 1) / (n - d) * pc
    elif kind == "chi2pom":
        h0 = float(np.mean(T2))
        v = float(np.var(T2, ddof=1)) if n > 1 else 0.0
        dof = max(int(np.round(2 * h0 ** 2 / v)) if v > 0 else 1, 1)
        lim = h0 * pc / dof
        scale = h0
    else:
        raise ValueError(kind)
    return lim, dof, scale


def q_limit(Q, kind, cl):
    """VAE_SIMCA.py:302-327 ('jm' takes its thetas from the Q sample itself)."""
    dof = scale = None
    if kind == "perc":
        lim = np.percentile(Q, cl * 100)
    elif kind == "jm":
        th1, th2, th3 = Q.sum(), (Q ** 2).sum(), (Q ** 3).sum()
        if th1 == 0:
            lim = 0
        else:
            h0 = max(1 - (2 * th1 * th3) / (3 * th2 ** 2), 1e-3)
            ca = np.sqrt(2) * special.erfinv(2 * cl - 1)
            h1 = ca * np.sqrt(2 * th2 * h0 ** 2) / th1
            h2 = th2 * h0 * (h0 - 1) / th1 ** 2
            lim = th1 * (1 + h1 + h2) ** (1 / h0)
    elif kind == "chi2pom":
        v0 = np.mean(Q)
        dof = max(round(2 * v0 ** 2 / np.var(Q, ddof=1)), 1)
        lim = v0 * np.percentile(Q, cl * 100) / dof
        scale = v0
    else:
        raise ValueError(kind)
    return lim, dof, scale


def d_limit(kind, T2_limit, Q_limit, Q, d, dcl, t2dof=None, qdof=None):
    """VAE_SIMCA.py:329-346."""
    if kind == "sim":
        return 1
    if kind == "alt":
        return np.sqrt(2)
    if kind == "ci":
        tr1 = d / T2_limit + Q.sum() / Q_limit
        tr2 = d / T2_limit ** 2 + (Q ** 2).sum() / Q_limit ** 2
        return (tr2 / tr1) * np.percentile(Q, dcl * 100)
    if kind == "dd":
        return t2dof + qdof
    raise ValueError(kind)


def fit_thresholds(mu_cal, mu_hat_cal, kind="alt", t2lim="Fdist", t2cl=0.95, qlim="jm",
                   qcl=0.95, dcl=0.95):
    """Calibration on latent means and their decode->encode round trip."""
    d = mu_cal.shape[1]
    m, Sinv = latent_stats(mu_cal)
    T2 = latent_t2(mu_cal, m, Sinv)
    Q = ((mu_cal.astype(np.float32) - mu_hat_cal.astype(np.float32)) ** 2).sum(1)
    tl, tdof, tsc = t2_limit(T2, d, t2lim, t2cl)
    ql, qdof, qsc = q_limit(Q, qlim, qcl)
    dl = d_limit(kind, tl, ql, Q, d, dcl, tdof, qdof)
    return dict(kind=kind, latent_mean=m, Sinv=Sinv, T2=T2, Q=Q, T2_limit=tl, Q_limit=ql,
                D_limit=dl, t2dof=tdof, t2scale=tsc, qdof=qdof, qscale=qsc, d=d)


def predict(model, mu, mu_hat):
    """VAE_SIMCA.py:361-382: 'alt' root-sum-square, 'dd' dof-scaled sum, anything else max."""
    T2 = latent_t2(mu, model["latent_mean"], model["Sinv"])
    Q = ((mu.astype(np.float32) - mu_hat.astype(np.float32)) ** 2).sum(1)
    if model["kind"] == "alt":
        D = np.sqrt((T2 / model["T2_limit"]) ** 2 + (Q / model["Q_limit"]) ** 2)
    elif model["kind"] == "dd":
        D = T2 * model["t2dof"] / model["t2scale"] + Q * model["qdof"] / model["qscale"]
    else:
        D = np.maximum(T2 / model["T2_limit"], Q / model["Q_limit"])
    return D < model["D_limit"], T2, Q
