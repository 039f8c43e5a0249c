"""ConvVAE1D bf16-autocast training step: module as it is vs the same parameters driven as 2-D channels-last convolutions."""
import json, sys
import torch
import torch.nn.functional as F
from torch import nn
sys.path.insert(0, ".")
from ocm_vae_simca_b200 import synth
from ocm_vae_simca_b200.vae import ConvVAE1D
from ocm_vae_simca_b200.vae.convvae import beta_vae_cosine_loss

dev = torch.device("cuda")
L, d, Bt = 1024, 16, 4096
X = synth.make_spectra_device(65536, L, dev, seed=3000)
mean = X.mean(0); std = X.std(0) + 1e-8
out = {}


def run_seq_cl(seq, x4):
    for m in seq:
        if isinstance(m, nn.Conv1d):
            x4 = F.conv2d(x4, m.weight.unsqueeze(2), m.bias, stride=(1, m.stride[0]), padding=(0, m.padding[0]))
        elif isinstance(m, nn.ConvTranspose1d):
            x4 = F.conv_transpose2d(x4, m.weight.unsqueeze(2), m.bias, stride=(1, m.stride[0]), padding=(0, m.padding[0]),
                                    output_padding=(0, m.output_padding[0]))
        elif isinstance(m, nn.BatchNorm1d):
            x4 = F.batch_norm(x4, m.running_mean, m.running_var, m.weight, m.bias, m.training, m.momentum, m.eps)
        else:
            x4 = m(x4)
    return x4


def forward_cl(vae, x):
    xs = (x - vae.spec_mean) / vae.spec_std
    h = run_seq_cl(vae.encoder_conv, xs.view(-1, 1, 1, L).contiguous(memory_format=torch.channels_last))
    h = vae.fc(h.squeeze(2).flatten(1))
    mu, logvar = vae.fc_mu(h), vae.fc_logvar(h)
    z = vae.reparameterize(mu, logvar)
    g = vae.fc_dec(z).view(-1, vae._enc_out_channels, 1, vae._enc_out_length).contiguous(memory_format=torch.channels_last)
    xr = run_seq_cl(vae.decoder_conv, g).reshape(x.shape[0], -1)[:, :L]
    return xr * vae.spec_std + vae.spec_mean, mu, logvar


def fwd_flag(v, x):
    v.channels_last = True
    return v(x)


for name, fwd in (("module", lambda v, x: v(x)), ("channels_last_2d (script)", forward_cl), ("ConvVAE1D.channels_last", fwd_flag)):
    torch.manual_seed(0)
    vae = ConvVAE1D(L, d, mean.cpu().numpy(), std.cpu().numpy()).to(dev).train()
    opt = torch.optim.Adam(vae.parameters(), lr=1e-4, fused=True)

    def step(i):
        xb = X[(i * Bt) % (X.shape[0] - Bt):][:Bt]
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            xr, mu_, lv_ = fwd(vae, xb)
            loss = beta_vae_cosine_loss(xb, xr.float(), mu_.float(), lv_.float(), as_floats=False)[0]
        loss.backward()
        opt.step()
        return loss
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        loss = step(3 + i)
    e1.record(); torch.cuda.synchronize()
    out[name] = {"ms_per_step": e0.elapsed_time(e1) / 10, "spectra_per_s": Bt * 10 / (e0.elapsed_time(e1) * 1e-3), "loss": float(loss)}
print(json.dumps(out))
