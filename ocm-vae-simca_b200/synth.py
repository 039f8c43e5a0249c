"""Deterministic synthetic "nut-like" NIR spectra.

The reference ships no data (its drivers read private Windows paths, e.g.
/root/reference/simca_nuts.py:12), so every parity test, golden fixture and
benchmark in this repo runs on the generator below.

spectrum(w) = 1 + 0.5*exp(-((w-0.5)/0.3)^2) + 0.2 * sum_k C_k B_k(w) + offset + noise

* K = 30 Gaussian bands B_k, centres spread over [0.05, 0.95], widths U(0.02, 0.10)
* scores C_k ~ N(0, decay^k) (geometric spectrum keeps neighbouring eigenvalues
  separated, which keeps the reference's unseeded randomized SVD reproducible --
  SURVEY.md §8 quirk Q1)
* per-row additive offset N(0, 0.05^2), white noise N(0, noise^2)
* out-of-class rows: same recipe with the band centres shifted by `shift`

Two back ends produce the same *distribution* (not the same bits):
`make_spectra` (numpy, host, seeded -- used for fixtures and parity tests) and
`make_spectra_device` (torch on the GPU, chunked -- used for the 10 M row
bench workload that never exists on the host).
"""
from __future__ import annotations

import numpy as np

N_BANDS = 30


def band_basis(p: int, seed: int = 7, shift: float = 0.0, dtype=np.float64):
    """Return (baseline[p], B[K,p]) for a wavelength axis of p points on [0,1]."""
    rng = np.random.default_rng(seed)
    w = np.linspace(0.0, 1.0, p, dtype=np.float64)
    centres = np.linspace(0.05, 0.95, N_BANDS) + rng.uniform(-0.01, 0.01, N_BANDS)
    widths = rng.uniform(0.02, 0.10, N_BANDS)
    centres = centres + shift
    B = np.exp(-((w[None, :] - centres[:, None]) / widths[:, None]) ** 2)
    baseline = 1.0 + 0.5 * np.exp(-((w - 0.5) / 0.3) ** 2)
    return baseline.astype(dtype), B.astype(dtype)


def make_spectra(n: int, p: int, seed: int = 0, shift: float = 0.0, decay: float = 0.8,
                 noise: float = 1e-3, basis_seed: int = 7, dtype=np.float64):
    """Host generator: returns X[n,p] (C-contiguous, `dtype`)."""
    baseline, B = band_basis(p, seed=basis_seed, shift=shift)
    rng = np.random.default_rng(seed)
    scales = decay ** np.arange(N_BANDS)
    C = rng.standard_normal((n, N_BANDS)) * scales[None, :]
    X = baseline[None, :] + 0.2 * (C @ B)
    X += rng.normal(0.0, 0.05, size=(n, 1))
    X += rng.normal(0.0, noise, size=(n, p))
    return np.ascontiguousarray(X, dtype=dtype)


def snv(X):
    """Row-wise standard normal variate exactly as the reference drivers do it
    (/root/reference/simca_nuts.py:48-50, utils/data_utils.py:57): population std,
    eps added to the std (not the variance)."""
    m = X.mean(axis=1, keepdims=True)
    s = X.std(axis=1, keepdims=True)
    return (X - m) / (s + 1e-8)


def make_class_dataset(n_target: int, n_other: int, p: int, seed: int = 0,
                       shift: float = 0.02, apply_snv: bool = True, dtype=np.float64):
    """Target class (label 0) + out-of-class rows (label 1), rows shuffled by a seeded
    permutation (a periodic label pattern aliases with venetian-blind folds and yields
    folds without negatives -- SURVEY.md §8(d))."""
    Xt = make_spectra(n_target, p, seed=seed, shift=0.0)
    Xo = make_spectra(n_other, p, seed=seed + 1, shift=shift)
    X = np.vstack([Xt, Xo])
    y = np.concatenate([np.zeros(n_target, dtype=np.int64), np.ones(n_other, dtype=np.int64)])
    perm = np.random.default_rng(seed + 2).permutation(X.shape[0])
    X, y = X[perm], y[perm]
    if apply_snv:
        X = snv(X)
    return np.ascontiguousarray(X, dtype=dtype), y


def make_object_dataset(nuts=("almond", "hazelnut", "walnut"), n_objects=8, rows=(30, 60), p=120, seed=0,
                        n_bad=3, n_gross=4):
    """Acquisition-shaped input of `object_aware_splits` (/root/reference/utils/data_utils.py:12): per nut type a list
    of objects, each {'spectral_data': float32 [rows_i, p]} -- with a few NaN / inf rows and a few gross outliers
    planted so that both screens have something to remove.  Each object gets its own offset so that objects differ."""
    rng = np.random.default_rng(seed)
    data = {}
    for k, nut in enumerate(nuts):
        objs = []
        for o in range(n_objects):
            r = int(rng.integers(rows[0], rows[1] + 1))
            X = make_spectra(r, p, seed=seed * 1000 + k * 100 + o, shift=0.02 * k).astype(np.float32)
            X += np.float32(0.01 * rng.standard_normal())
            objs.append({"spectral_data": X})
        for _ in range(n_gross):
            o = int(rng.integers(n_objects)); X = objs[o]["spectral_data"]
            X[int(rng.integers(X.shape[0]))] += (0.2 * rng.standard_normal(p)).astype(np.float32)
        for b in range(n_bad):
            o = int(rng.integers(n_objects)); X = objs[o]["spectral_data"]
            X[int(rng.integers(X.shape[0])), int(rng.integers(p))] = np.nan if b % 2 == 0 else np.inf
        data[nut] = objs
    return data


def make_spectra_device(n: int, p: int, device, seed: int = 0, shift: float = 0.0,
                        decay: float = 0.8, noise: float = 1e-3, chunk: int = 1 << 18,
                        out=None):
    """Device generator for the bulk workloads (torch Philox); fills `out` (or a new
    fp32 [n,p] tensor) chunk by chunk so that peak extra memory is one chunk."""
    import torch

    baseline, B = band_basis(p, shift=shift)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    base_d = torch.as_tensor(baseline, dtype=torch.float32, device=device)
    B_d = torch.as_tensor(0.2 * B, dtype=torch.float32, device=device)
    scales = torch.as_tensor(decay ** np.arange(N_BANDS), dtype=torch.float32, device=device)
    if out is None:
        out = torch.empty((n, p), dtype=torch.float32, device=device)
    for r0 in range(0, n, chunk):
        r1 = min(n, r0 + chunk)
        m = r1 - r0
        C = torch.randn((m, N_BANDS), generator=g, device=device) * scales
        blk = out[r0:r1]
        torch.matmul(C, B_d, out=blk)
        blk += base_d
        blk += 0.05 * torch.randn((m, 1), generator=g, device=device)
        blk += noise * torch.randn((m, p), generator=g, device=device)
    return out
