"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded
weights and inputs.  Runs only in the authoring container (the reference cannot travel); the
fixtures it writes are committed.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import mmdm_oracle as O  # noqa: E402
from oracle import ref_import as RI  # noqa: E402
from oracle import vae_oracle as VO  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def golden_unet(name, cfg, B, V, H, W, R, wseed, iseed, timestep):
    ref = RI.build_reference_unet(cfg)
    sd = O.init_state_dict(cfg, seed=wseed)
    ref.load_state_dict(sd)
    x, t, ctrl = O.make_inputs(cfg, B=B, V=V, H=H, W=W, R=R, seed=iseed, timestep=timestep)
    feats = {}
    hooks = []
    for blk_name in ["input_blocks.4", "middle_block"]:
        mod = ref.get_submodule(blk_name)
        hooks.append(mod.register_forward_hook(lambda m, i, o, n=blk_name: feats.__setitem__(n, o.detach().clone())))
    with torch.no_grad():
        y = ref(x, timesteps=t, context=None, control=ctrl)
    for h in hooks:
        h.remove()
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        cfg_keys=np.array(sorted(cfg.keys())),
        cfg_vals=np.array([str(cfg[k]) for k in sorted(cfg.keys())]),
        B=B, V=V, H=H, W=W, R=R, wseed=wseed, iseed=iseed, timestep=timestep,
        out=y.numpy(),
        **{"feat_" + k.replace(".", "_"): v.numpy() for k, v in feats.items()},
    )
    print(name, "out", tuple(y.shape), "absmax", float(y.abs().max()))


def golden_schedule():
    model = RI.build_reference_mmldm(O.TINY_CONFIG)
    _, Sampler, _ = RI.import_reference()
    out = dict(
        betas=model.betas.numpy(),
        alphas_cumprod=model.alphas_cumprod.numpy(),
        alphas_cumprod_prev=model.alphas_cumprod_prev.numpy(),
    )
    for S in (10, 100):
        s = Sampler(model)
        s.make_schedule(ddim_num_steps=S, ddim_eta=0.0, verbose=False)
        out[f"ddim_timesteps_{S}"] = np.asarray(s.ddim_timesteps)
        out[f"ddim_alphas_{S}"] = np.asarray(s.ddim_alphas)
        out[f"ddim_alphas_prev_{S}"] = np.asarray(s.ddim_alphas_prev)
        xs, es = [], []
        for index in range(S):
            # the reference's own coefficient arithmetic, cap4d/mmdm/sampler.py:215-229
            alpha_t = s.ddim_alphas.float()[index]
            somat = s.ddim_sqrt_one_minus_alphas[index]
            sigma_t = s.ddim_sigmas[index]
            alpha_prev_t = torch.tensor(s.ddim_alphas_prev).float()[index].double()
            somat = somat.double()
            alpha_t = alpha_t.double()
            e_f = -alpha_prev_t.sqrt() * somat / alpha_t.sqrt() + (1.0 - alpha_prev_t - sigma_t ** 2).sqrt()
            x_f = alpha_prev_t.sqrt() / alpha_t.sqrt()
            xs.append(float(x_f.float()))
            es.append(float(e_f.float()))
        out[f"x_factor_{S}"] = np.asarray(xs, dtype=np.float32)
        out[f"e_factor_{S}"] = np.asarray(es, dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "schedule.npz"), **out)
    print("schedule: acp[0], acp[500], acp[999] =", out["alphas_cumprod"][[0, 500, 999]])


def golden_sampler(name, n_ref, n_gen, S, R_max, cfg_scale, seed):
    cfg = O.TINY_CONFIG
    V, H, W = cfg["time_steps"], 8, 8
    model = RI.build_reference_mmldm(cfg)
    sd = O.init_state_dict(cfg, seed=3)
    model.model.diffusion_model.load_state_dict(sd)
    _, Sampler, _ = RI.import_reference()
    ref_cond, ref_unc, gen_cond, gen_unc = O.make_sampler_conditioning(cfg, n_ref, n_gen, H, W, seed=11)
    torch.manual_seed(seed)
    np.random.seed(seed)
    z = Sampler(model).sample(S=S, ref_cond=ref_cond, ref_uncond=ref_unc, gen_cond=gen_cond, gen_uncond=gen_unc,
                              latent_shape=(cfg["in_channels"], H, W), V=V, R_max=R_max, cfg_scale=cfg_scale)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), n_ref=n_ref, n_gen=n_gen, S=S, R_max=R_max,
                        cfg_scale=cfg_scale, seed=seed, H=H, W=W, V=V, wseed=3, cseed=11, out=z.numpy())
    print(name, tuple(z.shape), "absmax", float(z.abs().max()))


def golden_vae_encode(name, cfg, N, H, W, wseed, xseed):
    """AutoencoderKL.encode of the unmodified reference: posterior parameters, mode and a sample."""
    vae = RI.build_reference_vae(cfg)
    sd = VO.init_vae_state_dict(cfg, seed=wseed)
    sd.update(VO.init_vae_encoder_state_dict(cfg, seed=wseed))
    missing, unexpected = vae.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("loss") for k in missing), (missing, unexpected)
    x = torch.tanh(torch.randn(N, 3, H, W, generator=torch.Generator().manual_seed(xseed)))
    with torch.no_grad():
        post = vae.encode(x)
        torch.manual_seed(xseed + 1)
        sample = post.sample()
    mine = VO.vae_encode_moments(sd, cfg, x)
    print(name, "moments", tuple(post.parameters.shape), "absmax", float(post.parameters.abs().max()),
          "oracle err", float((mine - post.parameters).abs().max()))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), N=N, H=H, W=W, wseed=wseed, xseed=xseed,
                        moments=post.parameters.numpy(), mode=post.mode().numpy(), sample=sample.numpy())


def golden_vae(name, cfg, N, H, W, wseed, zseed):
    """decode_first_stage of the reference: z / scale_factor -> AutoencoderKL.decode (ddpm.py:822-830)."""
    vae = RI.build_reference_vae(cfg)
    sd = VO.init_vae_state_dict(cfg, seed=wseed)
    missing, unexpected = vae.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith(("encoder.", "quant_conv.", "loss.")) for k in missing), missing
    z = torch.randn(N, cfg["z_channels"], H, W, generator=torch.Generator().manual_seed(zseed)) * 0.8
    feats = {}
    hook = vae.decoder.mid.attn_1.register_forward_hook(lambda m, i, o: feats.__setitem__("attn", o.detach().clone()))
    with torch.no_grad():
        y = vae.decode(z / VO.SCALE_FACTOR)
    hook.remove()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), N=N, H=H, W=W, wseed=wseed, zseed=zseed,
                        cfg_keys=np.array(sorted(cfg.keys())), cfg_vals=np.array([str(cfg[k]) for k in sorted(cfg.keys())]),
                        out=y.numpy(), mid_attn=feats["attn"].numpy())
    print(name, tuple(y.shape), "std", float(y.std()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    golden_unet("unet_tiny_v4_h16", O.TINY_CONFIG, B=2, V=4, H=16, W=16, R=1, wseed=0, iseed=1, timestep=501)
    golden_unet("unet_tiny_v4_h8_r2", O.TINY_CONFIG, B=2, V=4, H=8, W=8, R=2, wseed=5, iseed=6, timestep=991)
    golden_schedule()
    golden_sampler("sampler_r1", n_ref=1, n_gen=6, S=4, R_max=4, cfg_scale=2.0, seed=124)
    golden_sampler("sampler_r2", n_ref=3, n_gen=4, S=5, R_max=2, cfg_scale=2.0, seed=7)
    golden_vae_encode("vae_enc_tiny_64", VO.TINY_VAE, N=2, H=64, W=64, wseed=0, xseed=2)
    golden_vae_encode("vae_enc_tiny_64x128", VO.TINY_VAE, N=1, H=64, W=128, wseed=5, xseed=6)
    golden_vae("vae_tiny_h8", VO.TINY_VAE, N=2, H=8, W=8, wseed=0, zseed=1)
    golden_vae("vae_tiny_h16x8", VO.TINY_VAE, N=1, H=16, W=8, wseed=3, zseed=4)
