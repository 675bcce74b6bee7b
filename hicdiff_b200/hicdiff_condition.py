"""Drop-in for `src/hicdiff_condition.py` of the reference: same `Unet` / `GaussianDiffusion` names and constructor
arguments (/root/reference/src/hicdiff_condition.py:255-269, 429-445), sm_100a execution."""
from .diffusion import GaussianDiffusionCond as GaussianDiffusion
from .diffusion import ModelPrediction
from .nets import Unet

__all__ = ["Unet", "GaussianDiffusion", "ModelPrediction"]
