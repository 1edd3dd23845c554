"""Hyper-parameters of the shipped MMDM (reference configs/mmdm/cap4d_mmdm_final.yaml:95-115) and of the
shipped generation workloads (configs/generation/{single_ref,multi_ref,debug}.yaml)."""

MMDM_UNET_CONFIG = dict(
    in_channels=4,
    out_channels=4,
    model_channels=320,
    condition_channels=50,
    num_res_blocks=2,
    channel_mult=(1, 2, 4, 4),
    attention_resolutions=(4, 2, 1),
    num_head_channels=64,
    time_steps=8,
)

# n_ddim_steps, cfg_scale, resolution, R_max, V, n_samples (+ number of reference views in the examples)
GENERATION_CONFIGS = {
    "single_ref": dict(n_ddim_steps=100, cfg_scale=2.0, resolution=512, seed=124, R_max=4, V=8, n_samples=840, n_ref=1),
    "multi_ref": dict(n_ddim_steps=100, cfg_scale=2.0, resolution=512, seed=124, R_max=4, V=8, n_samples=80, n_ref=10),
    "debug": dict(n_ddim_steps=10, cfg_scale=2.0, resolution=512, seed=124, R_max=4, V=8, n_samples=28, n_ref=1),
}
