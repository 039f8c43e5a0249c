"""CPU-only tests: C-ABI library loads and exports every declared symbol, host-side logic
(estimator protocol, splitter, limits, row sharding, gloo all-reduce of moments)."""
import os
import re

import numpy as np
import pytest

import ocm_vae_simca_b200 as ocm
from ocm_vae_simca_b200 import dist as odist
from ocm_vae_simca_b200 import limits, native, synth
from ocm_vae_simca_b200.utils import SIMCA, ClasswiseKFoldWithExternalVal, cross_validate_simca_grid
from ocm_vae_simca_b200.utils.CVSIMCA import _find_ncomp_param_name, _get_simca
from oracle import simca_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ocm_simca.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ocm_[a-z0-9_A-Z]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    native.build()
    import ctypes
    handle = ctypes.CDLL(native.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in include/ocm_simca.h but not exported"
    assert native.lib().ocm_abi_version() == 2
    # every bound signature is declared in the header as well
    for s in native.SIGNATURES:
        assert s in syms, f"{s} bound in native.py but not declared in the header"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    X, y = synth.make_class_dataset(40, 10, 30, seed=1)
    with pytest.raises(native.NativeError):
        SIMCA(n_components=2, model_class=0).fit(X, y)


def test_estimator_protocol():
    from sklearn.base import clone
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler
    m = SIMCA(n_components=3, model_class=0, type="ci", qlim="chi2box")
    c = clone(m)
    assert c.get_params()["type"] == "ci" and c.get_params()["n_components"] == 3
    c.set_params(n_components=7)
    assert c.n_components == 7
    pipe = Pipeline([("sc", StandardScaler()), ("simca", SIMCA(model_class=0))])
    assert _get_simca(pipe) is pipe.named_steps["simca"]
    assert _find_ncomp_param_name(pipe) == "simca__n_components"
    assert _find_ncomp_param_name(m) == "n_components"
    with pytest.raises(AttributeError):
        _get_simca(StandardScaler())


def test_metrics_match_oracle():
    rng = np.random.default_rng(0)
    y = rng.integers(0, 2, 200); p = rng.integers(0, 2, 200).astype(float)
    a = SIMCA()._metrics_simca_conformity(y, p, 0)
    b = so.conformity(y, p, 0)
    for k in ("TP", "TN", "FP", "FN"):
        assert int(a[k]) == int(b[k])
    for k in ("sensitivity", "specificity", "accuracy", "efficiency"):
        np.testing.assert_allclose(a[k], b[k])


def test_classwise_kfold_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "classwise_kfold.npz"))
    X, y = synth.make_class_dataset(270, 90, 150, seed=31)
    sp = ClasswiseKFoldWithExternalVal(n_splits=4, cls_label=0)
    for i, (tr, te) in enumerate(sp.split(X, y)):
        assert np.array_equal(tr, g[f"tr{i}"]) and np.array_equal(te, g[f"te{i}"])
    with pytest.raises(ValueError):
        list(ClasswiseKFoldWithExternalVal(n_splits=4, cls_label=0).split(X))
    with pytest.raises(ValueError):
        list(ClasswiseKFoldWithExternalVal(n_splits=4, cls_label=7).split(X, y))


def test_limits_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "limits_grid.npz"))
    X, y = synth.make_class_dataset(120, 40, 90, seed=11)
    base = so.fit_one_class(X[y == 0], 4)
    for i, name in enumerate(g["names"]):
        kind, t2l, ql = str(name).split("/")
        if kind == "dd":
            t2l = ql = "chi2pom"
        lim = limits.all_limits(kind, t2l, 0.95, ql, 0.95, 0.95, 4, base["n_samples"], base["eigs_all"],
                                base["T2"], base["Q"])
        np.testing.assert_allclose([lim.T2_limit, lim.Q_limit, lim.D_limit], g["limits"][i], rtol=1e-9,
                                   err_msg=str(name))


def test_limits_property_random_spectra():
    """Host limit arithmetic against the oracle's restatement of SIMCA.py:156-236 on random spectra / distances, and
    the partial-spectrum form (leading values + tail power sums, what the partial eigensolve provides) against the
    full eigenvalue array."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(2, 12), st.integers(30, 400), st.integers(0, 10_000), st.floats(0.80, 0.999),
           st.sampled_from(["sim", "alt", "ci", "dd"]), st.sampled_from(["Fdist", "Fdistrig", "chi2", "chi2pom", "perc"]),
           st.sampled_from(["jm", "chi2box", "chi2pom", "perc"]))
    def check(A, n, seed, cl, kind, t2l, ql):
        rng = np.random.default_rng(seed)
        eigs = np.sort(rng.gamma(1.0, 1.0, size=A + 25) * 0.7 ** np.arange(A + 25))[::-1].copy()
        T2 = rng.chisquare(A, size=n) * 1.1
        Q = rng.chisquare(6, size=n) * eigs[A:].sum() / 6
        if kind == "dd":
            t2l = ql = "chi2pom"
        tl, tdof, tsc = so.t2_limit(T2, A, t2l, cl)
        qlm, qdof, qsc = so.q_limit(Q, eigs, A, ql, cl)
        dl = so.d_limit(kind, A, tl, qlm, eigs, cl, tdof, qdof)
        lim = limits.all_limits(kind, t2l, cl, ql, cl, cl, A, n, eigs, T2, Q)
        np.testing.assert_allclose([lim.T2_limit, lim.Q_limit, lim.D_limit], [tl, qlm, dl], rtol=1e-10)
        tail = eigs[A + 3:]
        part = limits.Spectrum(eigs[:A + 3], (tail.sum(), (tail ** 2).sum(), (tail ** 3).sum()))
        lim2 = limits.all_limits(kind, t2l, cl, ql, cl, cl, A, n, part, T2, Q)
        np.testing.assert_allclose([lim2.T2_limit, lim2.Q_limit, lim2.D_limit], [tl, qlm, dl], rtol=1e-9)
    check()


def test_lazy_spectrum_and_model_dict():
    """`_model[cls]['eigs_all']` turns into the full eigenvalue array on first read, exactly once."""
    from ocm_vae_simca_b200.utils.SIMCA import _ModelDict
    calls = []
    full = np.array([5.0, 3.0, 1.0, 0.5, 0.2])
    sp = limits.LazySpectrum(full[:2], (full[2:].sum(), (full[2:] ** 2).sum(), (full[2:] ** 3).sum()),
                             lambda: (calls.append(1), full)[1])
    assert limits.tail_moments(sp, 1) == pytest.approx(limits.tail_moments(full, 1)) and not calls
    d = _ModelDict({"eigs_all": sp, "D_limit": 1.0})
    assert d["D_limit"] == 1.0 and not calls
    assert isinstance(d["eigs_all"], np.ndarray) and d["eigs_all"].shape == (5,) and len(calls) == 1
    assert np.array_equal(d.get("eigs_all"), full) and np.array_equal(dict(d.items())["eigs_all"], full) and len(calls) == 1
    assert np.array_equal(np.asarray(sp), full)


def test_shard_rows_partition():
    for n, w in [(10, 3), (7, 8), (10_000_000, 8), (0, 2)]:
        spans = [odist.shard_rows(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    odist.init("gloo")
    X = synth.make_spectra(301, 40, seed=3)
    lo, hi = odist.shard_rows(X.shape[0], rank, world)
    shift = X[:50].mean(0)
    Xs = X[lo:hi] - shift
    G = torch.from_numpy(Xs.T @ Xs)
    s = torch.from_numpy(Xs.sum(0))
    n = odist.allreduce_moments(G, s, hi - lo)
    d = s.numpy() / n
    C = (G.numpy() - n * np.outer(d, d)) / (n - 1)
    tabs = odist.allgather_counts(torch.tensor([[rank, 1, 2, 3]], dtype=torch.int64))
    mx = odist.max_over_ranks(float(rank), "cpu")
    if rank == 0:
        q.put((n, float(np.abs(C - np.cov(X, rowvar=False)).max()), [int(t[0, 0]) for t in tabs], mx))
    dist.destroy_process_group()


def test_gloo_world2_moment_allreduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    n, err, ranks, mx = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert n == 301 and err < 1e-12 and ranks == [0, 1] and mx == 1.0


def test_convvae_loads_both_reference_checkpoint_families():
    """vae_model.py checkpoints (threshold, threshold_q, threshold_h, threshold_f) and utils/final_vaesimca.py:96
    checkpoints (a single `q_threshold` buffer) both load with strict=True."""
    import torch
    from ocm_vae_simca_b200.vae import ConvVAE1D
    mk = lambda: ConvVAE1D(64, 4, np.zeros(64, np.float32), np.ones(64, np.float32), conv_blocks=2, n_filters=4, kernel_size=5, hidden_fc=8)
    a = mk()
    sd = {k: v.clone() for k, v in a.state_dict().items()}
    sd["threshold_q"] = torch.tensor(3.5)
    b = mk(); b.load_state_dict(sd, strict=True)
    assert float(b.threshold_q) == 3.5
    other = {k: v for k, v in sd.items() if not k.startswith("threshold")}
    other["q_threshold"] = torch.tensor(7.25)
    c = mk(); c.load_state_dict(other, strict=True)
    assert float(c.threshold_q) == 7.25 and float(c.threshold_h) == 0.0
    for k in a.state_dict():
        if not k.startswith("threshold"):
            assert torch.equal(c.state_dict()[k], a.state_dict()[k])


def test_frozen_scorer_network_equals_the_module():
    """vae.FrozenScorer (BatchNorm folded, 2-D channels-last convolutions) computes the module's encoder / deterministic
    decoder: fp32 on the CPU, trained-looking BatchNorm statistics, a length the strides do not divide."""
    import torch
    from ocm_vae_simca_b200.vae import ConvVAE1D, FrozenScorer
    torch.manual_seed(0)
    for L, blocks in ((200, 3), (131, 2)):
        vae = ConvVAE1D(L, 8, np.linspace(0, 1, L).astype(np.float32), np.full(L, 0.5, np.float32), conv_blocks=blocks,
                        n_filters=8, hidden_fc=32, dropout=0.1)
        vae.train()
        for _ in range(3):
            vae(torch.randn(16, L))
        vae.eval()
        fs = FrozenScorer(vae, dtype=torch.float32)
        x = torch.randn(7, L)
        mu, xr = fs.encode_decode(x)
        with torch.no_grad():
            xr0, mu0, _ = vae(x, sample=False)
        assert xr.shape == xr0.shape
        np.testing.assert_allclose(mu.numpy(), mu0.numpy(), atol=5e-6)
        np.testing.assert_allclose(xr.numpy(), xr0.numpy(), atol=5e-6)


def test_convvae_channels_last_mode_is_the_same_network():
    """ConvVAE1D.channels_last = True (2-D channels-last execution of the same parameters) against the module as the
    reference defines it: forward in train mode (batch statistics, running-stat updates), backward, eval mode."""
    import torch
    from ocm_vae_simca_b200.vae import ConvVAE1D
    L = 96
    mk = lambda: ConvVAE1D(L, 4, np.zeros(L, np.float32), np.ones(L, np.float32), conv_blocks=3, n_filters=4, hidden_fc=16)
    torch.manual_seed(3)
    a = mk(); b = mk(); b.load_state_dict(a.state_dict()); b.channels_last = True
    x = torch.randn(12, L)
    outs = []
    for m in (a, b):
        m.train()
        mu, lv = m.encode(x)
        xr = m.decode(mu)
        (xr.square().mean() + mu.square().mean() + lv.mean()).backward()
        outs.append((mu.detach(), xr.detach(), m.encoder_conv[0].weight.grad.clone(), m.decoder_conv[0].weight.grad.clone(),
                     m.encoder_conv[1].running_mean.clone(), int(m.encoder_conv[1].num_batches_tracked)))
    for u, v in zip(outs[0][:5], outs[1][:5]):
        np.testing.assert_allclose(v.numpy(), u.numpy(), rtol=1e-4, atol=1e-5)
    assert outs[0][5] == outs[1][5] == 1
    a.eval(); b.eval()
    with torch.no_grad():
        ra, rb = a(x, sample=False), b(x, sample=False)
    np.testing.assert_allclose(rb[0].numpy(), ra[0].numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(rb[1].numpy(), ra[1].numpy(), rtol=1e-4, atol=1e-5)
    assert set(a.state_dict()) == set(b.state_dict())


def test_host_streamable_gate():
    """engine.host_streamable: which host matrices SIMCA.predict / transform stream through the host-buffer call
    instead of uploading whole (shape / dtype / stride logic only; no device needed)."""
    from ocm_vae_simca_b200 import engine
    big = np.zeros((70_000, 1024), np.float32)                    # 287 MB > HOST_STREAM_BYTES
    assert big.nbytes > engine.HOST_STREAM_BYTES and engine.host_streamable(big)
    assert engine.host_streamable(np.zeros((40_000, 1024), np.float64))
    assert engine.host_streamable(np.zeros((70_000, 1100), np.float32)[:, :1024])      # row stride > p: fine
    assert not engine.host_streamable(big[:1000])                                      # small: one upload, one launch
    assert not engine.host_streamable(big.astype(np.float16))
    assert not engine.host_streamable(np.asfortranarray(big))                          # column-major
    assert not engine.host_streamable(big[:, ::2])                                     # strided columns
    assert not engine.host_streamable(big.reshape(70_000, 32, 32))
    assert not engine.host_streamable(big.tolist()[:2])
    assert not engine.host_streamable(np.zeros((0, 1024), np.float32))


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm): ONE JSON line with the GPU arm's
    metric / unit / config, `impl: reference`, a cpu_baseline describing the run and an e2e with no device traffic.
    Needs no GPU; one bounded step."""
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "spectra scored/s (fused T2+Q)" and d["unit"] == "spectra/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert "10000000 x 1024" in d["config"]["workload"] and d["dtype"] == "f32"
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
