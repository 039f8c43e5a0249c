"""ocm-vae-simca_b200 -- B200-native (sm_100a) SIMCA / DD-SIMCA hot path.

Drop-in for the reference's Python API on this path:
    from ocm_vae_simca_b200.utils import SIMCA, cross_validate_simca_grid, ClasswiseKFoldWithExternalVal
    from ocm_vae_simca_b200.vae import ConvVAE1D, VAESIMCA, compute_q_h_f
Everything numerical below the API runs in hand-written CUDA behind the C ABI declared
in include/ocm_simca.h (csrc/*.cu -> libocm_simca_b200.so).  There is no CPU fallback:
calling a compute entry point without the library or without a CUDA device raises.
"""
from . import synth  # noqa: F401  (pure numpy; safe on CPU)

__all__ = ["synth", "native", "engine", "utils", "vae"]
__version__ = "0.1.0"


def __getattr__(name):
    import importlib
    if name in ("native", "engine", "utils", "vae", "limits", "dist"):
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
