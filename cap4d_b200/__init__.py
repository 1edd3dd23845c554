"""cap4d_b200: B200-native (sm_100a) implementation of CAP4D's MMDM multi-view denoising hot path.

The compute lives in `libcap4d_b200.so` (hand-written CUDA behind the C ABI of
`include/cap4d_b200.h`); this package is the thin host-side mirror of the reference's Python
call conventions for that path.  Importing the package does not need a GPU; constructing
`B200MMDMUnet` does, and there is no CPU fallback.
"""
from .schedule import MMDMSchedule, ddim_factors, ddim_timesteps  # noqa: F401
from .unet import B200MMDMUnet, config_from_reference, install  # noqa: F401
from .sampler import B200MMLDM, B200StochasticIOSampler  # noqa: F401
from .vae import B200VAEDecoder, install_vae  # noqa: F401
from .conditioning import B200CAP4DConditioning, install_conditioning  # noqa: F401
from .output import convert_and_save_latent_images, save_flame_params, save_visualization  # noqa: F401

__all__ = ["B200MMDMUnet", "B200MMLDM", "B200StochasticIOSampler", "B200VAEDecoder", "install_vae", "B200CAP4DConditioning",
           "install_conditioning", "convert_and_save_latent_images", "save_flame_params", "save_visualization",
           "MMDMSchedule", "ddim_factors",
           "ddim_timesteps", "config_from_reference", "install"]
