"""hicdiff_b200 -- B200-native (sm_100a) implementation of HiCDiff's reverse-diffusion sampling hot path.

Public surface mirrors the reference's Python classes:
    hicdiff_b200.hicdiff_condition.{Unet, GaussianDiffusion}
    hicdiff_b200.hicdiff.{Unet, GaussianDiffusion}
    hicdiff_b200.hicdiff_sr3.{Unet, GaussianDiffusion}
    hicdiff_b200.model.hicedrn_Diff.hicedrn_Diff / hicdiff_b200.model.hicedrn_sr3_Diff.hicedrn_Diff
All compute goes through the C ABI in include/hicdiff_b200.h (libhicdiff_b200.so); there is no CPU fallback.
"""
__version__ = "0.1.0"
