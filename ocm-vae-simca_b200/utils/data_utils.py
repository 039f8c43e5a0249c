"""Object-aware calibration / validation / test splitting with the GPU preprocessing + outlier cut in front of it.

Mirror of /root/reference/utils/data_utils.py:12-169 (`object_aware_splits`): same arguments, same seven return values,
same numpy-in / numpy-out contract.  What moved to the B200:

  * the NaN / inf row screen                              (data_utils.py:49-54)   torch.isfinite on the device copy
  * SNV + Savitzky-Golay 1st derivative (w=5, poly 2)     (data_utils.py:57-61)   ocm_savgol_f32, one pass over HBM
  * PCA(<=10) score-space Mahalanobis distance + P95 cut  (data_utils.py:65-80)   Gram -> eigh -> fused T2 kernel

The object bookkeeping and sklearn's `train_test_split` over OBJECT ids stay on the host: a few dozen integers.
Only the boolean keep-mask travels back; the kept rows are gathered from the caller's own arrays, so the returned
spectra are bit-identical copies of the input rows.
"""
from __future__ import annotations

import numpy as np
from sklearn.model_selection import train_test_split

from .. import engine


def _empty(p):
    return np.empty((0, p), dtype=np.float32)


def _clean_rows(X, use_pca, outlier_percentile, tag):
    """keep-mask (host bool [n]) after the finite screen and the score-space outlier cut."""
    torch = engine._torch()
    Xd = engine.to_device(X, dtype=torch.float32)
    finite = torch.isfinite(Xd).all(dim=1)
    keep = finite.cpu().numpy()
    n_bad = int((~keep).sum())
    if n_bad:
        print(f"  WARNING: {tag}: Found {n_bad} NaN/inf samples. Removing them.")
        Xd = Xd[finite].contiguous()
    n, p = Xd.shape
    if not use_pca or n <= 3:
        return keep
    k = min(10, p, max(1, n - 1))
    if n <= k:
        print(f"  {tag}: not enough samples for PCA components, skipping outlier removal")
        return keep
    # reference: SavGol failure (window longer than the spectrum) silently leaves the SNV rows (data_utils.py:58-61)
    proc = engine.savgol(Xd, 5, 2, 1, snv=True) if p >= 5 else engine.snv(Xd)
    inlier, _, thr = engine.mahalanobis_outlier_mask(proc, k, float(outlier_percentile))
    inlier = inlier.cpu().numpy()
    if (~inlier).any():
        print(f"  {tag}: removed {int((~inlier).sum())} outliers (threshold {thr:.3f})")
    keep[np.flatnonzero(keep)[~inlier]] = False
    return keep


def _split_objects(ids, cal_frac, val_frac, test_frac, seed):
    """Whole objects to cal / val / test (data_utils.py:103-120)."""
    if len(ids) >= 3:
        cal, rest = train_test_split(ids, test_size=1.0 - cal_frac, random_state=seed)
        rel = test_frac / (val_frac + test_frac) if (val_frac + test_frac) > 0 else 0.5
        val, test = train_test_split(rest, test_size=rel, random_state=seed)
        return cal, val, test
    if len(ids) == 2:
        return [ids[0]], [], [ids[1]]
    return [ids[0]], [], []


def object_aware_splits(data, nut_types, target_nut, n_wavelengths, cal_frac=0.7, val_frac=0.15, test_frac=0.15,
                        random_state=42, outlier_percentile=95, use_pca=True):
    """See the reference docstring (data_utils.py:15-28).  Returns
    (splits, Xts_data, Xts_label, X_cal, X_val, X_test_in, X_test_out)."""
    assert abs(cal_frac + val_frac + test_frac - 1.0) < 1e-6, "Fractions must sum to 1.0"
    p = n_wavelengths
    splits = {}
    for nut in nut_types:
        blocks = [np.asarray(o["spectral_data"], dtype=np.float32) for o in data[nut]]
        if not blocks:
            print(f"  {nut}: no objects found, skipping")
            splits[nut] = {"cal": _empty(p), "val": _empty(p), "test": _empty(p)}
            continue
        X = np.vstack(blocks)
        owner = np.repeat(np.arange(len(blocks)), [b.shape[0] for b in blocks])
        keep = _clean_rows(X, use_pca, outlier_percentile, nut)
        X, owner = X[keep], owner[keep]
        alive = [int(i) for i in np.unique(owner)]
        if not alive:
            print(f"  {nut}: no objects remaining after cleaning, skipping")
            splits[nut] = {"cal": _empty(p), "val": _empty(p), "test": _empty(p)}
            continue
        parts = _split_objects(alive, cal_frac, val_frac, test_frac, random_state)
        rows = {name: (np.vstack([X[owner == i] for i in ids]) if len(ids) else _empty(p))
                for name, ids in zip(("cal", "val", "test"), parts)}
        splits[nut] = rows
        print(f"  {nut}: objects after cleaning={len(alive)}, raw samples after cleaning={X.shape[0]} -> "
              f"cal={rows['cal'].shape}, val={rows['val'].shape}, test={rows['test'].shape}")

    tests = [(nut, splits[nut]["test"]) for nut in nut_types if splits[nut]["test"].shape[0]]
    if tests:
        Xts_data = np.vstack([t for _, t in tests])
        Xts_label = np.concatenate([np.full(t.shape[0], 0 if nut == target_nut else 1, dtype=int) for nut, t in tests])
    else:
        Xts_data, Xts_label = _empty(p), np.array([], dtype=int)
    others = [t for nut, t in tests if nut != target_nut]
    X_test_out = np.vstack(others) if others else _empty(p)
    tgt = splits[target_nut]
    return splits, Xts_data, Xts_label, tgt["cal"], tgt["val"], tgt["test"], X_test_out
