"""The oracle (oracle/mmdm_oracle.py) against fixtures produced by the unmodified reference
(oracle/make_golden.py, run in the authoring container).  CPU only."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import mmdm_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _cfg(g):
    return {k: ast.literal_eval(v) for k, v in zip(g["cfg_keys"].tolist(), g["cfg_vals"].tolist())}


@pytest.mark.parametrize("name", ["unet_tiny_v4_h16", "unet_tiny_v4_h8_r2"])
def test_unet_forward_matches_reference(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = _cfg(g)
    sd = O.init_state_dict(cfg, seed=int(g["wseed"]))
    x, t, ctrl = O.make_inputs(cfg, B=int(g["B"]), V=int(g["V"]), H=int(g["H"]), W=int(g["W"]), R=int(g["R"]),
                               seed=int(g["iseed"]), timestep=int(g["timestep"]))
    taps = {}
    y = O.unet_forward(sd, cfg, x, t, ctrl, taps=taps)
    ref = torch.from_numpy(g["out"])
    assert y.shape == ref.shape
    # fp32 CPU vs fp32 CPU, same torch build: only summation-order noise is allowed
    assert O.max_rel_err(y, ref) < 2e-5
    R = int(g["R"])
    # reference views return exactly x - z_input (mmdm_unet.py:77,125)
    assert torch.equal(y[:, :R], (x - ctrl["z_input"])[:, :R])
    assert O.max_rel_err(taps["input_blocks.4"], torch.from_numpy(g["feat_input_blocks_4"])) < 2e-5
    assert O.max_rel_err(taps["middle_block"], torch.from_numpy(g["feat_middle_block"])) < 2e-5


def test_schedule_matches_reference_bit_exact():
    g = np.load(os.path.join(GOLD, "schedule.npz"))
    betas, acp, acp_prev = O.mmdm_schedule()
    assert np.array_equal(betas.astype(np.float32), g["betas"])
    assert np.array_equal(acp.astype(np.float32), g["alphas_cumprod"])
    assert np.array_equal(acp_prev.astype(np.float32), g["alphas_cumprod_prev"])
    # SURVEY.md 8a4 probe values
    assert abs(float(g["alphas_cumprod"][0]) - 0.997754) < 1e-6
    assert abs(float(g["alphas_cumprod"][500]) - 0.107163) < 1e-6
    for S in (10, 100):
        ts, a, ap, sg = O.ddim_schedule(g["alphas_cumprod"], S)
        assert np.array_equal(ts, g[f"ddim_timesteps_{S}"])
        assert np.array_equal(np.asarray(a), g[f"ddim_alphas_{S}"])
        assert np.array_equal(ap, g[f"ddim_alphas_prev_{S}"])
        for index in range(S):
            x_f, e_f = O.ddim_coefficients(a, ap, sg, index)
            assert float(x_f) == float(g[f"x_factor_{S}"][index])
            assert float(e_f) == float(g[f"e_factor_{S}"][index])


def test_ddim_timesteps_edge_case():
    # 1000 // 3 = 333 -> timesteps 1, 334, 667, 1000: the reference indexes out of range (util.py:65)
    acp = O.mmdm_schedule()[1].astype(np.float32)
    with pytest.raises(IndexError):
        O.ddim_schedule(acp, 3)


@pytest.mark.parametrize("name", ["sampler_r1", "sampler_r2"])
def test_sampler_matches_reference(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = O.TINY_CONFIG
    sd = O.init_state_dict(cfg, seed=int(g["wseed"]))
    H, W, V = int(g["H"]), int(g["W"]), int(g["V"])
    rc, ru, gc, gu = O.make_sampler_conditioning(cfg, int(g["n_ref"]), int(g["n_gen"]), H, W, seed=int(g["cseed"]))
    acp = O.mmdm_schedule()[1].astype(np.float32)
    torch.manual_seed(int(g["seed"]))
    np.random.seed(int(g["seed"]))
    z = O.stochastic_io_sample(lambda x, t, c: O.unet_forward(sd, cfg, x, t, c), acp, int(g["S"]), rc, ru, gc, gu,
                               (cfg["in_channels"], H, W), V=V, R_max=int(g["R_max"]), cfg_scale=float(g["cfg_scale"]))
    ref = torch.from_numpy(g["out"])
    assert z.shape == ref.shape
    assert O.max_rel_err(z, ref) < 1e-4
    assert O.psnr(z, ref) > 80.0


def test_param_census_production():
    shapes = O.unet_param_shapes(O.PRODUCTION_CONFIG)
    assert len(shapes) == 576
    assert sum(int(np.prod(s)) for s in shapes.values()) == 815_549_764
    zero = sum(int(np.prod(s)) for n, s in shapes.items() if O._is_zero_module(n))
    norm_bias = sum(int(np.prod(s)) for n, s in shapes.items()
                    if n.endswith("bias") and (n.startswith("out.0.") or any(
                        k in n for k in (".in_layers.0.", ".out_layers.0.", ".norm.", ".norm1.", ".norm3."))))
    # SURVEY.md note Z counts every all-zero parameter of a fresh model: zero_module()s + norm biases
    assert zero + norm_bias == 224_928_644


# ---------------------------------------------------------------------------------------------
# VAE decode: oracle/vae_oracle.py against the reference AutoencoderKL fixtures
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["vae_tiny_h8", "vae_tiny_h16x8"])
def test_vae_oracle_matches_reference_fixture(name):
    from oracle import vae_oracle as VO

    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = VO.TINY_VAE
    assert {k: str(cfg[k]) for k in sorted(cfg)} == dict(zip(g["cfg_keys"].tolist(), g["cfg_vals"].tolist()))
    sd = VO.init_vae_state_dict(cfg, seed=int(g["wseed"]))
    z = torch.randn(int(g["N"]), 4, int(g["H"]), int(g["W"]), generator=torch.Generator().manual_seed(int(g["zseed"]))) * 0.8
    taps = {"decoder.mid.attn_1": None}
    y = VO.vae_decode(sd, cfg, z, taps=taps)
    ref = torch.from_numpy(g["out"])
    assert y.shape == ref.shape == (int(g["N"]), 3, 8 * int(g["H"]), 8 * int(g["W"]))
    assert O.max_rel_err(y, ref) < 2e-5
    assert O.max_rel_err(taps["decoder.mid.attn_1"], torch.from_numpy(g["mid_attn"])) < 2e-5


def test_vae_production_census():
    from oracle import vae_oracle as VO

    shapes = VO.vae_param_shapes(VO.PRODUCTION_VAE)
    # decoder of the Stable-Diffusion KL-f8 autoencoder + post_quant_conv: 138 + 2 tensors, 49.49 M parameters
    assert len(shapes) == 140
    assert sum(int(np.prod(s)) for s in shapes.values()) == 49_490_199
    kinds = [k for k, *_ in VO.vae_decoder_topology(VO.PRODUCTION_VAE)]
    assert kinds.count("res") == 14 and kinds.count("up") == 3 and kinds.count("attn") == 1


@pytest.mark.parametrize("name", ["vae_enc_tiny_64", "vae_enc_tiny_64x128"])
def test_vae_encode_oracle_matches_reference_fixture(name):
    """AutoencoderKL.encode of the unmodified reference (oracle/make_golden.py:golden_vae_encode)."""
    from oracle import vae_oracle as VO

    g = np.load(os.path.join(GOLD, name + ".npz"))
    sd = VO.init_vae_state_dict(VO.TINY_VAE, seed=int(g["wseed"]))
    sd.update(VO.init_vae_encoder_state_dict(VO.TINY_VAE, seed=int(g["wseed"])))
    x = torch.tanh(torch.randn(int(g["N"]), 3, int(g["H"]), int(g["W"]),
                               generator=torch.Generator().manual_seed(int(g["xseed"]))))
    m = VO.vae_encode_moments(sd, VO.TINY_VAE, x)
    assert float((m - torch.from_numpy(g["moments"])).abs().max()) <= 3e-6
    assert torch.equal(torch.from_numpy(g["mode"]), torch.from_numpy(g["moments"])[:, :4])
    torch.manual_seed(int(g["xseed"]) + 1)
    s = VO.posterior_sample(m, torch.randn(m.shape[0], 4, *m.shape[2:]))
    assert float((s - torch.from_numpy(g["sample"])).abs().max()) <= 3e-6
