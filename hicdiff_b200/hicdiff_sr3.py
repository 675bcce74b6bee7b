"""Drop-in for `src/hicdiff_sr3.py` (SR3-style continuous noise-level conditioning) of the reference
(/root/reference/src/hicdiff_sr3.py:310-325, 491-505).  As in the reference, `Unet(noise_level_emb=True)` selects the
PositionalEncoding + additive FeatureWiseAffine variant."""
from .diffusion import GaussianDiffusionSR3 as GaussianDiffusion
from .diffusion import ModelPrediction
from .nets import Unet

__all__ = ["Unet", "GaussianDiffusion", "ModelPrediction"]
