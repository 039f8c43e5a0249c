"""Secondary measurements -- NOT the contract bench (that is /bench.py).

One JSON line per kernel family, device-timed with CUDA events after warm-up, inputs larger than L2:
  preprocess (SNV)            HBM, 8p B/spectrum
  latent / recon scoring      HBM, 4p + (4p | 2p) + 4d + 8 B/spectrum   (VAE-SIMCA epilogue, config 5 shape)
  Gram accumulation           FP64 pipe / tensor pipe, n p (p+1) flop
  covariance eigensolve       latency
  ConvVAE1D forward bf16      PyTorch (cuDNN / cuBLAS), reported separately as the north star asks

    python tools/bench_extra.py [--rows 1000000]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ocm_vae_simca_b200 import engine, native, synth
from ocm_vae_simca_b200.vae import ConvVAE1D


def peak_hbm():
    try:
        return json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def timed(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--gram-rows", type=int, default=500_000)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    pk = peak_hbm()
    n, p, d = args.rows, 1024, 16
    X = synth.make_spectra_device(n, p, dev, seed=1)
    out = torch.empty_like(X)
    lines = []

    ms = timed(lambda: engine.preprocess(X, "snv", out=out))
    gbs = n * 8 * p / ms / 1e6
    lines.append({"kernel": "preprocess_kernel (SNV)", "rows": n, "p": p, "ms": ms, "spectra_per_s": n / ms * 1e3,
                  "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"]}})

    ms = timed(lambda: engine.savgol(X, 5, 2, 1, snv=True, out=out))
    gbs = n * 8 * p / ms / 1e6
    lines.append({"kernel": "savgol_kernel (SNV + SavGol w5 p2 d1)", "rows": n, "p": p, "ms": ms, "spectra_per_s": n / ms * 1e3,
                  "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"]}})

    # VAE-SIMCA epilogue at the config-5 shape: latent 16, reconstruction fp32 and bf16
    z = torch.randn(n, d, device=dev)
    m = torch.zeros(d, device=dev)
    Sinv = torch.eye(d, device=dev)
    for name, xr in (("fp32", out), ("bf16", out.to(torch.bfloat16))):
        ms = timed(lambda: engine.latent_score(z=z, m=m, Sinv=Sinv, x=X, xrec=xr))
        bytes_per = 4 * p + (4 if name == "fp32" else 2) * p + 4 * d + 8
        gbs = n * bytes_per / ms / 1e6
        lines.append({"kernel": f"latent_score_kernel (x fp32, xrec {name}, d=16)", "rows": n, "ms": ms, "spectra_per_s": n / ms * 1e3,
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"]}})
    del out, z

    # Gram accumulation (p = 2048 as in config 4, bounded rows)
    gn, gp = args.gram_rows, 2048
    Xg = synth.make_spectra_device(gn, gp, dev, seed=2)
    shift = Xg[:4096].double().mean(0)
    G = torch.zeros((gp, gp), dtype=torch.float64, device=dev)
    ms = timed(lambda: engine.gram(Xg, None, shift, out=G), warm=1, reps=3)
    fl = gn * gp * (gp + 1)
    lines.append({"kernel": "gram_dmma_kernel (fp64, FP64 tensor pipe: mma.sync m8n8k4)", "rows": gn, "p": gp, "ms": ms,
                  "roofline": {"bound": "tensor", "achieved": fl / ms / 1e9, "peak": pk.get("bf16_tflops"), "unit": "TFLOP/s",
                               "frac": fl / ms / 1e9 / pk.get("bf16_tflops", 1590.0),
                               "note": "algorithmic flops n*p*(p+1); split-precision passes are not credited"}})
    s0, mn, mx, cnt = engine.colstats_minmax(Xg)
    G2 = torch.zeros((gp, gp), dtype=torch.float64, device=dev)
    for products in (10, 13, 16):
        ms = timed(lambda: engine.gram_tc(Xg, None, shift, mn, mx, products=products, out=G2), warm=1, reps=3)
        G2.zero_(); engine.gram_tc(Xg, None, shift, mn, mx, products=products, out=G2)
        G.zero_(); engine.gram(Xg, None, shift, out=G)
        err = float(((torch.triu(G2) - torch.triu(G)).abs().max() / torch.triu(G).abs().max()).item())
        lines.append({"kernel": f"gram_tc_kernel (tcgen05 kind::i8, {products} digit products, int32 TMEM -> fp64)", "rows": gn, "p": gp,
                      "ms": ms, "max_rel_diff_vs_dmma": err,
                      "roofline": {"bound": "tensor", "achieved": fl / ms / 1e9, "peak": pk.get("bf16_tflops"), "unit": "TFLOP/s",
                                   "frac": fl / ms / 1e9 / pk.get("bf16_tflops", 1590.0),
                                   "note": "algorithmic flops n*p*(p+1); digit products are not credited"}})
    Cm, _, _ = engine.covariance(Xg[:100_000].contiguous())
    ms = timed(lambda: engine.eigh_topk(Cm, 20), warm=1, reps=3)
    lines.append({"kernel": "eigh_topk (p = 2048, fp64, top-20 vectors + all values)", "ms": ms})
    del Xg, G

    # ConvVAE1D forward (PyTorch) in bf16 autocast + fused epilogue = config-5 scoring step
    L = 1024
    vae = ConvVAE1D(L, d, np.zeros(L, np.float32), np.ones(L, np.float32)).to(dev).eval()
    from ocm_vae_simca_b200.vae.vae_simca import latent_recon_scores
    B = 8192
    xb = X[:B]

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return latent_recon_scores(vae, xb, sample=False, recon_dtype=torch.bfloat16)
    ms = timed(step, warm=2, reps=5)
    lines.append({"kernel": "ConvVAE1D bf16 forward + fused latent/recon epilogue (config-5 scoring step)", "batch": B, "ms": ms,
                  "spectra_per_s": B / ms * 1e3, "note": "encoder/decoder are PyTorch (cuDNN/cuBLAS); epilogue is ocm_latent_score_f32"})
    # config 5, training side: one bf16-autocast optimisation step of the reference architecture (PyTorch; not rewritten)
    from ocm_vae_simca_b200.vae.convvae import beta_vae_cosine_loss
    vae.train()
    opt = torch.optim.Adam(vae.parameters(), lr=1e-3, fused=True)
    Bt = 4096
    xt = X[:Bt]

    def train_step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            xr, mu_, lv_ = vae(xt)[:3]
            loss = beta_vae_cosine_loss(xt, xr.float(), mu_.float(), lv_.float())[0]
        loss.backward()
        opt.step()
    ms = timed(train_step, warm=3, reps=10)
    lines.append({"kernel": "ConvVAE1D bf16-autocast training step (fwd + bwd + fused Adam, cosine beta-VAE loss)", "batch": Bt, "ms": ms,
                  "spectra_per_s": Bt / ms * 1e3, "note": "PyTorch cuDNN/cuBLAS; 17 M parameters; reference trains in fp32 at batch 512"})
    for ln in lines:
        print(json.dumps(ln))


if __name__ == "__main__":
    main()
