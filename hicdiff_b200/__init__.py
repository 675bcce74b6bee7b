"""hicdiff_b200 -- B200-native (sm_100a) implementation of HiCDiff's reverse-diffusion sampling hot path.

Public surface mirrors the reference's Python classes:
    hicdiff_b200.hicdiff_condition.{Unet, GaussianDiffusion}
    hicdiff_b200.hicdiff.{Unet, GaussianDiffusion}
    hicdiff_b200.hicdiff_sr3.{Unet, GaussianDiffusion}
    hicdiff_b200.model.hicedrn_Diff.hicedrn_Diff / hicdiff_b200.model.hicedrn_sr3_Diff.hicedrn_Diff
    hicdiff_b200.functions.denoising.efficient_generalized_steps        (DDRM sampler, src/functions/denoising.py)
    hicdiff_b200.metrics.{ssim, SSIM, psnr, inverse_data_transform, get_metrics}   (src/Utils/loss/SSIM.py, metrics_cond.py)
    hicdiff_b200.prepare.{load_both_constraints, make_splits}            (processdata/PrepareData_linear.py)
    hicdiff_b200.genome.denoise_chromosomes, hicdiff_b200.ops.{tile_extract, tile_scatter}   (splitPieces and its inverse)
Training: `loss = diffusion(x); loss.backward()` runs forward + loss + backward on the device for every eps-net
(hicdiff_b200.train); `train.enable_gradient_allreduce(model)` adds the data-parallel gradient exchange.
All compute goes through the C ABI in include/hicdiff_b200.h (libhicdiff_b200.so); there is no CPU fallback.
"""
__version__ = "0.1.0"
