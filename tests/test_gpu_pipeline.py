"""The four stages either side of the sampler wired together on the GPU behind the reference's call conventions
(generate_images.py:80-143): VAE encode + conditioning maps per frame (inference/utils.py:64-100) ->
StochasticIOSampler.sample -> decode + files
(inference/utils.py:125-137), every stage on libcap4d_b200.so, checked against the oracles run on the same inputs.

Tolerances are the per-stage ones: conditioning 2e-6, sampler >= 40 dB PSNR, decoded images >= 35 dB after both
(the latents fed to the decoder already differ by the sampler's bf16 error)."""
import os

import numpy as np
import pytest
import torch

from oracle import cond_oracle as CO
from oracle import mmdm_oracle as O
from oracle import vae_oracle as VO

pytestmark = pytest.mark.gpu


def test_conditioning_sampler_decode_files(cuda_device, tmp_path):
    from cap4d_b200 import (B200CAP4DConditioning, B200MMDMUnet, B200MMLDM, B200StochasticIOSampler, B200VAEDecoder,
                            convert_and_save_latent_images)
    from cap4d_b200 import output as OUT

    dev = cuda_device
    cfg = O.TINY_CONFIG  # V = 4, 50 conditioning channels
    S, n_ref, n_gen = 8, 1, 6
    tv, faces, fmask = CO.make_mesh(10, 12, seed=2)
    props = CO.normalize_props(tv)
    cond = B200CAP4DConditioning(torch.from_numpy(faces), torch.from_numpy(props), torch.from_numpy(fmask),
                                 image_size=S, super_resolution=2, use_crop_mask=True)
    n = n_ref + n_gen
    verts, offs = CO.make_views(tv, n, seed=3)
    g = torch.Generator().manual_seed(4)
    ray = torch.nn.functional.normalize(torch.randn(n, 3, S, S, generator=g), dim=1).numpy()
    ref_mask = np.zeros((n, S, S), np.float32)
    ref_mask[:n_ref] = 1
    crop = np.ones((n, S, S), np.float32)

    # -- stages 0 + 1: VAE encode of every frame + conditioning, driven like get_condition_from_dataloader
    #    (cap4d/inference/utils.py:64-100: one frame per batch, B = T = 1)
    from cap4d_b200.conditioning import concat_frames, get_condition_from_dataloader

    vsd = VO.init_vae_state_dict(VO.TINY_VAE, seed=0)
    vsd.update(VO.init_vae_encoder_state_dict(VO.TINY_VAE, seed=0))
    vae = B200VAEDecoder(VO.TINY_VAE, vsd, device=dev)
    imgs_in = torch.tanh(torch.randn(n, 8 * S, 8 * S, 3, generator=g))
    imgs_in[n_ref:] = 0  # frames to generate carry a zero image (inference_data.py:84)
    frames = [{"jpg": imgs_in[i][None, None],
               "hint": {"verts_2d": torch.from_numpy(verts[i])[None, None], "offsets_3d": torch.from_numpy(offs[i])[None, None],
                        "reference_mask": torch.from_numpy(ref_mask[i])[None, None], "ray_map": torch.from_numpy(ray[i])[None, None],
                        "out_crop_mask": torch.from_numpy(crop[i])[None, None]},
               "flame_params": {"fx": torch.full((1, 1, 1), 1000.0 + i)}} for i in range(n)]
    torch.manual_seed(5)
    data = get_condition_from_dataloader(cond, vae, frames, dev, to_cpu=True)
    torch.manual_seed(5)
    on_dev = get_condition_from_dataloader(cond, vae, frames[:2], dev)  # device-resident variant: same numbers
    assert on_dev["cond_frames"]["pos_enc"][1].is_cuda
    assert torch.equal(on_dev["cond_frames"]["z_input"][1].cpu(), data["cond_frames"]["z_input"][1])
    assert len(data["flame_params"]) == n and float(data["flame_params"][3]["fx"][0, 0]) == 1003.0
    all_c, all_u = concat_frames(data["cond_frames"]), concat_frames(data["uncond_frames"])
    cut = lambda d, lo, hi: {k: v[lo:hi] for k, v in d.items()}  # noqa: E731
    ref_c, ref_u, gen_c, gen_u = cut(all_c, 0, n_ref), cut(all_u, 0, n_ref), cut(all_c, n_ref, n), cut(all_u, n_ref, n)
    want_pe = CO.cond_pos_enc(verts, offs, faces, props, fmask, ray, ref_mask, crop, S, 2)
    got_pe = all_c["pos_enc"].numpy()
    assert got_pe.shape == (n, S, S, 50)
    assert np.all(np.abs(got_pe - want_pe) <= 2e-6 + 2e-6 * np.abs(want_pe))
    assert float(gen_u["pos_enc"].abs().max()) == 0 and float(all_u["z_input"].abs().max()) == 0
    # the latents of the frames against the oracle encoder with the reference's RNG use (one randn per frame)
    torch.manual_seed(5)
    want_z = torch.cat([VO.SCALE_FACTOR * VO.posterior_sample(
        VO.vae_encode_moments(vsd, VO.TINY_VAE, imgs_in[i].permute(2, 0, 1)[None]), torch.randn(1, 4, S, S)) for i in range(n)])
    assert all_c["z_input"].shape == (n, 4, S, S) and O.max_rel_err(all_c["z_input"], want_z) <= 2e-2
    z_ref = ref_c["z_input"]  # the sampler stages below are compared on the SAME reference latents

    # -- stage 2: the sampler on those dicts, against the oracle sampler on the ORACLE's conditioning
    sd = O.init_state_dict(cfg, seed=0)
    unet = B200MMDMUnet(cfg, sd, device=dev)
    torch.manual_seed(7)
    np.random.seed(7)
    z = B200StochasticIOSampler(B200MMLDM(unet), groups_per_call=2).sample(
        S=4, ref_cond=ref_c, ref_uncond=ref_u, gen_cond=gen_c, gen_uncond=gen_u, latent_shape=(4, S, S), V=4, R_max=4,
        cfg_scale=2.0)
    assert z.shape == (n_gen, 4, S, S)
    pe = torch.from_numpy(want_pe)
    rm = torch.from_numpy(ref_mask)[:, None]
    o_ref_c = {"z_input": z_ref, "ref_mask": rm[:n_ref], "pos_enc": pe[:n_ref]}
    o_gen_c = {"z_input": gen_c["z_input"], "ref_mask": rm[n_ref:], "pos_enc": pe[n_ref:]}
    zero = lambda c: {"z_input": c["z_input"] * 0, "ref_mask": c["ref_mask"], "pos_enc": c["pos_enc"] * 0}  # noqa: E731
    acp = O.mmdm_schedule()[1].astype(np.float32)
    torch.manual_seed(7)
    np.random.seed(7)
    z_want = O.stochastic_io_sample(lambda a, b, c: O.unet_forward(sd, cfg, a, b, c), acp, 4, o_ref_c, zero(o_ref_c),
                                    o_gen_c, zero(o_gen_c), (4, S, S), V=4, R_max=4, cfg_scale=2.0)
    assert O.psnr(z.cpu(), z_want) >= 40.0

    # -- stage 3: decode + files
    ref_dir, gen_dir = OUT.make_output_dirs(tmp_path / "out")
    assert convert_and_save_latent_images(z.cpu(), vae, dev, gen_dir, batch=4) == n_gen
    assert convert_and_save_latent_images(ref_c["z_input"], vae, dev, ref_dir) == n_ref
    imgs = OUT.read_output_images(gen_dir).astype(np.float64)
    assert imgs.shape == (n_gen, 8 * S, 8 * S, 3) and sorted(os.listdir(ref_dir / "images")) == ["00000.png"]
    want_img = OUT.to_uint8_bgr(VO.vae_decode(vsd, VO.TINY_VAE, z_want))[..., ::-1].astype(np.float64)
    mse = ((imgs - want_img) ** 2).mean()
    assert 10 * np.log10(255.0 ** 2 / max(mse, 1e-12)) >= 35.0
