"""Drop-in for `src/hicdiff.py` (unconditional DDPM) of the reference; `Unet` defaults to self_condition=False here
(/root/reference/src/hicdiff.py:263), `GaussianDiffusion.sample` returns tiles, stacked traces on dim 1 (:617)."""
from functools import wraps

from .diffusion import GaussianDiffusionUncond as GaussianDiffusion
from .diffusion import ModelPrediction
from .nets import Unet as _Unet


class Unet(_Unet):
    @wraps(_Unet.__init__)
    def __init__(self, dim, *args, self_condition=False, **kwargs):
        super().__init__(dim, *args, self_condition=self_condition, **kwargs)


__all__ = ["Unet", "GaussianDiffusion", "ModelPrediction"]
