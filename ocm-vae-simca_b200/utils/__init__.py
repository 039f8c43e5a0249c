"""Reference-compatible namespace: mirrors /root/reference/utils/__init__.py for the hot path
(`from ocm_vae_simca_b200.utils import SIMCA, cross_validate_simca_grid, ...`)."""
from .SIMCA import SIMCA
from .CVSIMCA import ClasswiseKFoldWithExternalVal, cross_validate_simca_grid, plot_cv
from .data_utils import object_aware_splits

__all__ = ["SIMCA", "cross_validate_simca_grid", "plot_cv", "ClasswiseKFoldWithExternalVal", "object_aware_splits"]
