"""Pin the CPU oracle (oracle/) against fixtures produced by the live reference
(oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
from sklearn.model_selection import PredefinedSplit

import ocm_vae_simca_b200 as ocm
from ocm_vae_simca_b200 import synth
from oracle import simca_oracle as so
from oracle import vae_oracle as vo


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def sign_norm(P):
    idx = np.argmax(np.abs(P), axis=1)
    s = np.sign(P[np.arange(P.shape[0]), idx])
    return P * s[:, None]


def test_c1_fit_predict(golden_dir):
    g = load(golden_dir, "c1_fit_predict.npz")
    X, y = synth.make_class_dataset(200, 100, 700, seed=0)
    m = so.fit_one_class(X[y == 0], 5)
    np.testing.assert_allclose(m["eigs_all"], g["eigs_all"], rtol=1e-10, atol=1e-22)
    np.testing.assert_allclose(sign_norm(m["P"]), sign_norm(g["P"]), atol=1e-10)
    np.testing.assert_allclose(m["mean"], g["xmean"], rtol=1e-13)
    np.testing.assert_allclose(m["T2"], g["T2_train"], rtol=1e-9)
    np.testing.assert_allclose(m["Q"], g["Q_train"], rtol=1e-9)
    for k in ("T2_limit", "Q_limit", "D_limit"):
        np.testing.assert_allclose(m[k], g[k], rtol=1e-12)
    T2, Q = so.score(m, X)
    np.testing.assert_allclose(T2, g["T2"], rtol=1e-8)
    np.testing.assert_allclose(Q, g["Q"], rtol=1e-8)
    assert np.array_equal(so.predict(m, X), g["pred"][:, 0])


def test_all_limit_combinations(golden_dir):
    g = load(golden_dir, "limits_grid.npz")
    X, y = synth.make_class_dataset(120, 40, 90, seed=11)
    Xc = X[y == 0]
    for i, name in enumerate(g["names"]):
        kind, t2l, ql = str(name).split("/")
        m = so.fit_one_class(Xc, 4, kind=kind, t2lim=t2l, qlim=ql)
        got = np.array([m["T2_limit"], m["Q_limit"], m["D_limit"]], dtype=np.float64)
        np.testing.assert_allclose(got, g["limits"][i], rtol=1e-9, err_msg=str(name))
        assert np.array_equal(so.predict(m, X).astype(np.uint8), g["preds"][i]), name


def test_ncomp_sweep(golden_dir):
    g = load(golden_dir, "ncomp_sweep.npz")
    X, y = synth.make_class_dataset(300, 100, 300, seed=21)
    Xte, _ = synth.make_class_dataset(60, 60, 300, seed=22)
    for i, A in enumerate(g["A_values"]):
        m = so.fit_one_class(X[y == 0], int(A))
        T2, Q = so.score(m, Xte)
        np.testing.assert_allclose(T2, g["T2"][i], rtol=1e-7)
        np.testing.assert_allclose(Q, g["Q"][i], rtol=1e-7)
        np.testing.assert_allclose([m["T2_limit"], m["Q_limit"], m["D_limit"]], g["limits"][i], rtol=1e-10)


def test_ncomp_one_extension():
    # the reference crashes at A == 1 (SIMCA.py:69); the oracle extends it with a 1x1 covariance
    X, y = synth.make_class_dataset(80, 20, 40, seed=5)
    m = so.fit_one_class(X[y == 0], 1)
    assert m["Sinv"].shape == (1, 1)
    np.testing.assert_allclose(m["Sinv"][0, 0], 1.0 / m["eigs_all"][0], rtol=1e-10)


def test_cv_small(golden_dir):
    g = load(golden_dir, "cv_small.npz")
    X, y = synth.make_class_dataset(270, 90, 150, seed=31)
    splits = list(PredefinedSplit(np.arange(X.shape[0]) % 6).split())
    res = so.cv_grid(X, y, splits, range(2, 9), model_class=0, refit=False)
    table = np.array([[r["LV"], r["spec"], r["sens"], r["eff"]] for r in res["results"]])
    np.testing.assert_allclose(table, g["table"], rtol=1e-12)
    assert res["best_LV"] == int(g["best_LV"])
    for i, lv in enumerate(range(2, 9)):
        assert np.array_equal(res["pooled"][lv].astype(np.uint8), g["pooled"][i])


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "golden", "c2_cv_grid.npz")),
                    reason="full-size C2 fixture not generated")
def test_c2_single_column(golden_dir):
    # one LV column of the full 2000x1050 grid keeps the CPU suite short
    g = load(golden_dir, "c2_cv_grid.npz")
    X, y = synth.make_class_dataset(1500, 500, 1050, seed=1234)
    splits = list(PredefinedSplit(np.arange(X.shape[0]) % 10).split())[:2]
    pooled = np.unpackbits(g["pooled_packed"], axis=1)[:, : X.shape[0]]
    lv = 10
    col = list(g["table"][:, 0].astype(int)).index(lv)
    for tr, te in splits:
        m = so.fit_one_class(X[tr][y[tr] == 0], lv)
        assert np.array_equal(so.predict(m, X[te]).astype(np.uint8), pooled[col][te])


def test_object_aware_splits(golden_dir):
    """Oracle restatement of object_aware_splits and of its SNV + SavGol + Mahalanobis step vs the live reference."""
    g = load(golden_dir, "object_splits.npz")
    data = synth.make_object_dataset(seed=int(g["seed"]))
    nuts = ["almond", "hazelnut", "walnut"]
    splits, Xts, yts, Xcal, Xval, Xtin, Xtout = so.object_aware_splits(data, nuts, "almond", 120)
    for got, key in ((Xts, "Xts"), (yts, "yts"), (Xcal, "Xcal"), (Xval, "Xval"), (Xtin, "Xtin"), (Xtout, "Xtout")):
        assert np.array_equal(got, g[key]), key
    assert [[splits[n][k].shape[0] for k in ("cal", "val", "test")] for n in nuts] == g["shapes"].tolist()
    h = load(golden_dir, "snv_savgol_mahal.npz")
    Xp = so.snv_savgol(h["X"])
    np.testing.assert_allclose(Xp, h["Xp"], rtol=0, atol=1e-12)
    keep, d, thr = so.mahalanobis_filter(Xp, 10, 95)
    np.testing.assert_allclose(d, h["mahal"], rtol=1e-9)
    np.testing.assert_allclose(thr, h["thr"], rtol=1e-9)
    # degenerate inputs the reference handles: empty nut, one and two objects
    tiny = {"a": [], "b": data["almond"][:1], "c": data["walnut"][:2]}
    sp, Xts, yts, Xcal, *_ = so.object_aware_splits(tiny, ["a", "b", "c"], "b", 120)
    assert sp["a"]["cal"].shape == (0, 120) and sp["b"]["test"].shape[0] == 0 and sp["c"]["val"].shape[0] == 0
    assert set(yts.tolist()) == {1} and Xcal.shape[0] > 0


def test_conformity_matches_counts():
    y = np.array([0, 0, 1, 1, 0, 1]); p = np.array([1, 0, 0, 1, 1, 0.0])
    m = so.conformity(y, p, 0)
    assert (m["TP"], m["FN"], m["TN"], m["FP"]) == (2, 1, 2, 1)
    np.testing.assert_allclose(m["efficiency"], np.sqrt(m["sensitivity"] * m["specificity"]))


def test_vae_qhf(golden_dir):
    g = load(golden_dir, "vae_qhf.npz")
    q, h, f, qc, hc, fc = vo.q_h_f(g["x"], g["x_rec"], g["z"])
    np.testing.assert_allclose(q, g["q"], rtol=2e-5)
    np.testing.assert_allclose(h, g["h"], rtol=2e-4)
    np.testing.assert_allclose(f, g["f"], rtol=5e-4)
    np.testing.assert_allclose([qc, hc, fc], g["crit"], rtol=5e-4)


def test_vae_simca(golden_dir):
    import torch
    from ocm_vae_simca_b200.vae import ConvVAE1D
    g = load(golden_dir, "vae_simca.npz")
    L, d = int(g["L"]), int(g["d"])
    vae = ConvVAE1D(L, d, np.zeros(L, np.float32), np.ones(L, np.float32), conv_blocks=2,
                    n_filters=4, kernel_size=5, hidden_fc=16)
    vae.load_state_dict({str(k): torch.from_numpy(g["sd::" + str(k)]) for k in g["state_keys"]})
    vae.eval()

    def latents(X):
        with torch.no_grad():
            x = torch.from_numpy(X)
            mu, _ = vae.encode((x - vae.spec_mean) / vae.spec_std)
            mu_hat, _ = vae.encode((vae.decode(mu) - vae.spec_mean) / vae.spec_std)
        return mu.numpy(), mu_hat.numpy()

    mu_c, mh_c = latents(g["Xcal"]); mu_t, mh_t = latents(g["Xte"])
    for tag in ("alt_Fdist_jm", "sim_perc_perc", "dd_chi2pom_chi2pom", "alt_chi2_jm"):
        kind, t2l, ql = tag.split("_")
        m = vo.fit_thresholds(mu_c, mh_c, kind=kind, t2lim=t2l, qlim=ql)
        np.testing.assert_allclose([m["T2_limit"], m["Q_limit"], m["D_limit"]], g[tag + "::limits"], rtol=1e-4)
        yp, T2, Q = vo.predict(m, mu_t, mh_t)
        np.testing.assert_allclose(T2, g[tag + "::T2"], rtol=1e-4)
        np.testing.assert_allclose(Q, g[tag + "::Q"], rtol=1e-4, atol=1e-7)
        assert np.mean(yp.astype(np.uint8) != g[tag + "::pred"]) < 0.02
