"""How fast can PyTorch drive the ConvVAE1D scoring pass? (config 5; the network stays in torch)
variants: autocast as bench.py has it / cudnn.benchmark / frozen bf16 copy with BatchNorm folded / + 2-D channels_last."""
import json, sys, time
import torch
from torch import nn
sys.path.insert(0, ".")
from ocm_vae_simca_b200 import engine, synth
from ocm_vae_simca_b200.vae import ConvVAE1D
from ocm_vae_simca_b200.vae.vae_simca import latent_recon_scores

dev = torch.device("cuda")
L, d, n, Bs = 1024, 16, 262144, 8192
X = synth.make_spectra_device(n, L, dev, seed=3000)
mean = X[:65536].mean(0); std = X[:65536].std(0) + 1e-8
torch.manual_seed(0)
vae = ConvVAE1D(L, d, mean.cpu().numpy(), std.cpu().numpy()).to(dev).eval()
out = {}


def timed(f, reps=2):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record(); torch.cuda.synchronize()
    return n * reps / (e0.elapsed_time(e1) * 1e-3)


def pass_autocast():
    for r in range(0, n, Bs):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            latent_recon_scores(vae, X[r:r + Bs], sample=False, recon_dtype=torch.bfloat16)


out["autocast"] = timed(pass_autocast)
torch.backends.cudnn.benchmark = True
out["autocast+cudnn.benchmark"] = timed(pass_autocast)


def fold(seq, transpose):
    from torch.nn.utils.fusion import fuse_conv_bn_eval
    mods, outm = list(seq), []
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, (nn.Conv1d, nn.ConvTranspose1d)) and i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm1d):
            outm.append(fuse_conv_bn_eval(m, mods[i + 1], transpose=isinstance(m, nn.ConvTranspose1d)))
            i += 2
        else:
            outm.append(m); i += 1
    return nn.Sequential(*outm)


import copy
fz = copy.deepcopy(vae)
fz.encoder_conv = fold(fz.encoder_conv, False)
fz.decoder_conv = fold(fz.decoder_conv, True)
fz = fz.to(torch.bfloat16)
fz.spec_mean = vae.spec_mean.clone(); fz.spec_std = vae.spec_std.clone()
fz.latent_mean = vae.latent_mean.clone(); fz.latent_cov_inv = vae.latent_cov_inv.clone()


def frozen_scores(model, xb):
    with torch.no_grad():
        xs = ((xb - model.spec_mean) / model.spec_std).to(torch.bfloat16)
        mu, _ = model.encode(xs)
        xr = model.decode(mu) * model.spec_std.to(torch.bfloat16) + model.spec_mean.to(torch.bfloat16)
        return engine.latent_score(z=mu.float(), m=model.latent_mean, Sinv=model.latent_cov_inv, x=xb, xrec=xr)


def pass_frozen():
    for r in range(0, n, Bs):
        frozen_scores(fz, X[r:r + Bs])


out["frozen bf16, BN folded"] = timed(pass_frozen)
# agreement of the frozen copy with the autocast path
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    d2a, qa, _ = latent_recon_scores(vae, X[:Bs], sample=False, recon_dtype=torch.bfloat16)
d2f, qf = frozen_scores(fz, X[:Bs])
out["frozen_vs_autocast_rel_q"] = float(((qf - qa).abs() / qa).median())
out["frozen_vs_autocast_rel_d2"] = float(((d2f - d2a).abs() / (d2a + 1e-6)).median())

# channels_last through 2-D convolutions (H = 1)
def to2d(seq):
    mods = []
    for m in seq:
        if isinstance(m, nn.Conv1d):
            c = nn.Conv2d(m.in_channels, m.out_channels, (1, m.kernel_size[0]), stride=(1, m.stride[0]), padding=(0, m.padding[0]))
            c.weight.data = m.weight.data.unsqueeze(2).clone(); c.bias.data = m.bias.data.clone(); mods.append(c)
        elif isinstance(m, nn.ConvTranspose1d):
            c = nn.ConvTranspose2d(m.in_channels, m.out_channels, (1, m.kernel_size[0]), stride=(1, m.stride[0]), padding=(0, m.padding[0]),
                                   output_padding=(0, m.output_padding[0]))
            c.weight.data = m.weight.data.unsqueeze(2).clone(); c.bias.data = m.bias.data.clone(); mods.append(c)
        else:
            mods.append(m)
    return nn.Sequential(*mods).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)


enc2, dec2 = to2d(fz.encoder_conv), to2d(fz.decoder_conv)


def cl_scores(xb):
    with torch.no_grad():
        xs = ((xb - fz.spec_mean) / fz.spec_std).to(torch.bfloat16)
        h = enc2(xs.view(-1, 1, 1, L).contiguous(memory_format=torch.channels_last))
        h = h.squeeze(2).flatten(1)                      # [N, C, L'] logical order, as fc expects
        h = fz.fc(h)
        mu = fz.fc_mu(h)
        g = fz.fc_dec(mu).view(-1, fz._enc_out_channels, 1, fz._enc_out_length).contiguous(memory_format=torch.channels_last)
        xr = dec2(g).reshape(xb.shape[0], -1)[:, :L]
        xr = xr * fz.spec_std.to(torch.bfloat16) + fz.spec_mean.to(torch.bfloat16)
        return engine.latent_score(z=mu.float(), m=fz.latent_mean, Sinv=fz.latent_cov_inv, x=xb, xrec=xr.contiguous())


def pass_cl():
    for r in range(0, n, Bs):
        cl_scores(X[r:r + Bs])


try:
    out["frozen bf16 + channels_last 2-D"] = timed(pass_cl)
    d2c, qc = cl_scores(X[:Bs])
    out["cl_vs_frozen_rel_q"] = float(((qc - qf).abs() / qf).median())
except Exception as e:
    out["channels_last_error"] = repr(e)[:300]
print(json.dumps(out))
