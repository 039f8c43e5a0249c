"""ConvVAE1D -- the encoder/decoder that feeds VAE-SIMCA scoring (PyTorch; north-star item 5
keeps the network itself in torch, cuDNN/cuBLAS do the convolutions).

Architecture, buffer names and state-dict keys follow /root/reference/vae_model.py:6-129
so that a checkpoint trained with the reference loads here unchanged
(`encoder_conv`, `fc`, `fc_mu`, `fc_logvar`, `fc_dec`, `decoder_conv`; buffers
`threshold{,_q,_h,_f}`, `spec_mean`, `spec_std`, `latent_mean`, `latent_cov_inv`).
Differences, all opt-in: `forward(x, sample=False)` gives the deterministic
reconstruction dec(mu) (the reference always samples z, even in eval -- SURVEY.md Q6),
and `encode_decode_scores` returns what the fused scoring kernel consumes.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


def _conv_out_len(n, k, stride, pad):
    return (n + 2 * pad - (k - 1) - 1) // stride + 1


def _maybe(layer_or_none):
    return [] if layer_or_none is None else [layer_or_none]


class ConvVAE1D(nn.Module):
    def __init__(self, input_length, latent_dim, mean, std, conv_blocks=3, n_filters=32,
                 kernel_size=9, stride=2, hidden_fc=256, activation="elu", dropout=0.0,
                 use_batchnorm=True, beta=1.0):
        super().__init__()
        self.input_length, self.latent_dim = input_length, latent_dim
        self.beta, self.dropout, self.use_batchnorm = beta, dropout, use_batchnorm
        for name in ("threshold", "threshold_q", "threshold_h", "threshold_f"):
            self.register_buffer(name, torch.tensor(0.0))

        Act = nn.ELU if activation == "elu" else nn.GELU
        pad = kernel_size // 2
        drop = (lambda: nn.Dropout(dropout)) if dropout > 0 else None

        # encoder: first block keeps the length, later ones stride down; channels double (cap 1024)
        widths = [min(n_filters * 2 ** b, 1024) for b in range(conv_blocks)]
        layers, cin, length = [], 1, input_length
        for b, cout in enumerate(widths):
            s = 1 if b == 0 else stride
            layers += [nn.Conv1d(cin, cout, kernel_size, stride=s, padding=pad)]
            layers += _maybe(nn.BatchNorm1d(cout) if use_batchnorm else None)
            layers += [Act()] + _maybe(drop() if drop else None)
            length = _conv_out_len(length, kernel_size, s, pad)
            cin = cout
        self.encoder_conv = nn.Sequential(*layers)
        self._enc_out_channels, self._enc_out_length = cin, length
        flat = cin * length

        self.fc = nn.Sequential(nn.Linear(flat, hidden_fc), Act(), drop() if drop else nn.Identity())
        self.fc_mu = nn.Linear(hidden_fc, latent_dim)
        self.fc_logvar = nn.Linear(hidden_fc, latent_dim)
        self.fc_dec = nn.Sequential(nn.Linear(latent_dim, hidden_fc), Act(),
                                    drop() if drop else nn.Identity(),
                                    nn.Linear(hidden_fc, flat), Act())

        # decoder: mirror; channels halve (floor n_filters); last block has stride 1
        layers, ch = [], cin
        for b in range(conv_blocks):
            nxt = max(ch // 2, n_filters)
            s = stride if b < conv_blocks - 1 else 1
            layers += [nn.ConvTranspose1d(ch, nxt, kernel_size, stride=s, padding=pad,
                                          output_padding=s - 1)]
            layers += _maybe(nn.BatchNorm1d(nxt) if use_batchnorm else None)
            layers += [Act()] + _maybe(drop() if drop else None)
            ch = nxt
        layers += [nn.Conv1d(ch, 1, kernel_size=1)]
        self.decoder_conv = nn.Sequential(*layers)

        self.register_buffer("spec_mean", torch.as_tensor(mean, dtype=torch.float32))
        self.register_buffer("spec_std", torch.as_tensor(std, dtype=torch.float32))
        self.register_buffer("latent_mean", torch.zeros(latent_dim))
        self.register_buffer("latent_cov_inv", torch.eye(latent_dim))
        self._init_weights()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        """Checkpoints of the reference's other VAE script (utils/final_vaesimca.py:96 registers ONE buffer called
        `q_threshold` and saves `model.state_dict()`): map it onto `threshold_q` and let the threshold buffers that
        script never had keep their defaults, so that both families of reference checkpoints load with strict=True."""
        old = prefix + "q_threshold"
        if old in state_dict:
            state_dict[prefix + "threshold_q"] = state_dict.pop(old)
        for name in ("threshold", "threshold_q", "threshold_h", "threshold_f", "latent_mean", "latent_cov_inv"):
            if prefix + name not in state_dict:
                state_dict[prefix + name] = getattr(self, name).detach().clone()
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def _init_weights(self):
        for m in self.modules():
            if isinstance(m, (nn.Conv1d, nn.ConvTranspose1d, nn.Linear)):
                nn.init.kaiming_normal_(m.weight, nonlinearity="linear")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    # -- channels-last execution (opt-in: `vae.channels_last = True`) --------------------
    # Same parameters, same state dict, same arithmetic: every Conv1d / ConvTranspose1d runs as a 2-D convolution with
    # H = 1 on channels-last activations (the weights are unsqueezed VIEWS, so autograd and the optimiser see the module's
    # own parameters).  That is the layout cuDNN's tensor-core kernels take without transposes: the bf16-autocast
    # training step of config 5 goes from 81.8 to 51.2 ms (batch 4096 x 1024, one B200, tools/vae_train_tune.py).
    channels_last = False

    @staticmethod
    def _run_channels_last(seq, x4):
        for m in seq:
            if isinstance(m, nn.ConvTranspose1d):
                x4 = F.conv_transpose2d(x4, m.weight.unsqueeze(2), m.bias, stride=(1, m.stride[0]), padding=(0, m.padding[0]),
                                        output_padding=(0, m.output_padding[0]))
            elif isinstance(m, nn.Conv1d):
                x4 = F.conv2d(x4, m.weight.unsqueeze(2), m.bias, stride=(1, m.stride[0]), padding=(0, m.padding[0]))
                if m.in_channels == 1 and not x4.is_contiguous(memory_format=torch.channels_last):
                    # a 1-channel input has no memory format of its own, so cuDNN answers in NCHW and every later layer
                    # (BatchNorm's generic kernels included) would follow: convert once, here
                    x4 = x4.contiguous(memory_format=torch.channels_last)
            elif isinstance(m, nn.BatchNorm1d):
                if m.training and m.track_running_stats and m.num_batches_tracked is not None:
                    m.num_batches_tracked.add_(1)
                x4 = F.batch_norm(x4, m.running_mean, m.running_var, m.weight, m.bias, m.training or not m.track_running_stats,
                                  m.momentum if m.momentum is not None else 0.0, m.eps)
            else:
                x4 = m(x4)
        return x4

    # -- reference-visible API ---------------------------------------------------------
    def encode(self, x):
        if self.channels_last:
            x4 = x.reshape(x.shape[0], 1, 1, -1).contiguous(memory_format=torch.channels_last)
            h = self._run_channels_last(self.encoder_conv, x4).squeeze(2).flatten(1)
        else:
            h = self.encoder_conv(x.unsqueeze(1)).flatten(1)
        h = self.fc(h)
        return self.fc_mu(h), self.fc_logvar(h)

    def reparameterize(self, mu, logvar):
        return mu + torch.randn_like(mu) * torch.exp(0.5 * logvar)

    def decode(self, z):
        if self.channels_last:
            h = self.fc_dec(z).view(z.size(0), self._enc_out_channels, 1, self._enc_out_length)
            out = self._run_channels_last(self.decoder_conv, h.contiguous(memory_format=torch.channels_last)).reshape(z.size(0), -1)
        else:
            h = self.fc_dec(z).view(z.size(0), self._enc_out_channels, self._enc_out_length)
            out = self.decoder_conv(h).squeeze(1)
        n = out.shape[-1]
        if n > self.input_length:
            out = out[..., : self.input_length]
        elif n < self.input_length:
            out = F.pad(out, (0, self.input_length - n))
        return out

    def forward(self, x, sample=True):
        mu, logvar = self.encode((x - self.spec_mean) / self.spec_std)
        z = self.reparameterize(mu, logvar) if sample else mu
        return self.decode(z) * self.spec_std + self.spec_mean, mu, logvar


def beta_vae_cosine_loss(x, x_recon, mu, logvar, beta=1.0, eps=1e-8, as_floats=True):
    """Chord-distance reconstruction + beta*KL (reference vae_model.py:136-145).  `as_floats=False` returns the two
    monitoring terms as device tensors (the reference's `.cpu().item()` is two host synchronisations per step)."""
    cos = F.cosine_similarity(x.flatten(1), x_recon.flatten(1), dim=1, eps=1e-12)
    cos = cos.clamp(-1.0 + eps, 1.0 - eps)
    rec = torch.sqrt(2.0 * (1.0 - cos)).mean()
    kl = -0.5 * (1 + logvar - mu.pow(2) - logvar.exp()).sum(1).mean()
    if not as_floats:
        return rec + beta * kl, rec.detach(), kl.detach()
    return rec + beta * kl, rec.detach().cpu().item(), kl.detach().cpu().item()


def beta_vae_bce_loss(x, x_recon, mu, logvar, beta=1.0, eps=1e-8):
    """Logit-BCE against the per-row min-max scaled target + beta*KL (vae_model.py:148-158)."""
    lo = x.amin(dim=1, keepdim=True)
    hi = x.amax(dim=1, keepdim=True)
    target = ((x - lo) / (hi - lo + eps)).clamp(0.0, 1.0)
    rec = F.binary_cross_entropy_with_logits(x_recon.flatten(1), target.flatten(1), reduction="mean")
    kl = -0.5 * (1 + logvar - mu.pow(2) - logvar.exp()).sum(1).mean()
    return rec + beta * kl, rec.detach().cpu().item(), kl.detach().cpu().item()
