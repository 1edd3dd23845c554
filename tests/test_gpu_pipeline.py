"""The three stages either side of the sampler wired together on the GPU behind the reference's call conventions
(generate_images.py:80-143): conditioning maps (cap4dcond.py) -> StochasticIOSampler.sample -> decode + files
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
    z_ref = torch.randn(n_ref, 4, S, S, generator=g)

    # -- stage 1: per-frame conditioning exactly as get_condition_from_dataloader drives it (B = T = 1 per frame)
    frames_c, frames_u = [], []
    for i in range(n):
        d = lambda x: torch.as_tensor(x[i:i + 1]).to(dev)[None]  # noqa: E731
        batch = {"verts_2d": d(verts), "offsets_3d": d(offs), "reference_mask": d(ref_mask), "ray_map": d(ray),
                 "out_crop_mask": d(crop),
                 "z": (z_ref[i:i + 1] if i < n_ref else torch.zeros(1, 4, S, S)).to(dev)[None]}
        frames_c.append({k: v[0].cpu() for k, v in cond(batch, unconditional=False).items()})
        frames_u.append({k: v[0].cpu() for k, v in cond(batch, unconditional=True).items()})
    cat = lambda fr, lo, hi: {k: torch.cat([f[k] for f in fr[lo:hi]], 0) for k in fr[0]}  # noqa: E731
    ref_c, ref_u = cat(frames_c, 0, n_ref), cat(frames_u, 0, n_ref)
    gen_c, gen_u = cat(frames_c, n_ref, n), cat(frames_u, n_ref, n)
    want_pe = CO.cond_pos_enc(verts, offs, faces, props, fmask, ray, ref_mask, crop, S, 2)
    got_pe = torch.cat([ref_c["pos_enc"], gen_c["pos_enc"]]).numpy()
    assert got_pe.shape == (n, S, S, 50)
    assert np.all(np.abs(got_pe - want_pe) <= 2e-6 + 2e-6 * np.abs(want_pe))
    assert float(gen_u["pos_enc"].abs().max()) == 0 and torch.equal(ref_c["z_input"], z_ref)

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
    o_gen_c = {"z_input": torch.zeros(n_gen, 4, S, S), "ref_mask": rm[n_ref:], "pos_enc": pe[n_ref:]}
    zero = lambda c: {"z_input": c["z_input"] * 0, "ref_mask": c["ref_mask"], "pos_enc": c["pos_enc"] * 0}  # noqa: E731
    acp = O.mmdm_schedule()[1].astype(np.float32)
    torch.manual_seed(7)
    np.random.seed(7)
    z_want = O.stochastic_io_sample(lambda a, b, c: O.unet_forward(sd, cfg, a, b, c), acp, 4, o_ref_c, zero(o_ref_c),
                                    o_gen_c, zero(o_gen_c), (4, S, S), V=4, R_max=4, cfg_scale=2.0)
    assert O.psnr(z.cpu(), z_want) >= 40.0

    # -- stage 3: decode + files
    vsd = VO.init_vae_state_dict(VO.TINY_VAE, seed=0)
    vae = B200VAEDecoder(VO.TINY_VAE, vsd, device=dev)
    ref_dir, gen_dir = OUT.make_output_dirs(tmp_path / "out")
    assert convert_and_save_latent_images(z.cpu(), vae, dev, gen_dir, batch=4) == n_gen
    assert convert_and_save_latent_images(ref_c["z_input"], vae, dev, ref_dir) == n_ref
    imgs = OUT.read_output_images(gen_dir).astype(np.float64)
    assert imgs.shape == (n_gen, 8 * S, 8 * S, 3) and sorted(os.listdir(ref_dir / "images")) == ["00000.png"]
    want_img = OUT.to_uint8_bgr(VO.vae_decode(vsd, VO.TINY_VAE, z_want))[..., ::-1].astype(np.float64)
    mse = ((imgs - want_img) ** 2).mean()
    assert 10 * np.log10(255.0 ** 2 / max(mse, 1e-12)) >= 35.0
