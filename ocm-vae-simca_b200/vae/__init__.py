from .convvae import ConvVAE1D, beta_vae_bce_loss, beta_vae_cosine_loss  # noqa: F401


def __getattr__(name):
    if name in ("VAESIMCA", "compute_q_h_f", "latent_recon_scores", "FrozenScorer"):
        from . import vae_simca
        return getattr(vae_simca, name)
    raise AttributeError(name)
