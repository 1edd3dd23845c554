"""VAE decode on the GPU, through the C ABI, against the committed fixtures the unmodified
reference AutoencoderKL produced and against the oracle on the same seeded inputs.

Tolerance: the decoder's output is an image in about [-1, 1] that the reference quantises to uint8
(cap4d/inference/utils.py:134-137, step 2/255 = 7.8e-3): max|a-b|/max|b| <= 2e-2 with bf16 operands and
PSNR >= 40 dB (peak = the reference image's range)."""
import os

import numpy as np
import pytest
import torch

from oracle import mmdm_oracle as O
from oracle import vae_oracle as VO

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 2e-2


@pytest.fixture(scope="module")
def tiny_vaes(cuda_device):
    from cap4d_b200 import B200VAEDecoder

    cache = {}

    def get(seed):
        if seed not in cache:
            sd = VO.init_vae_state_dict(VO.TINY_VAE, seed=seed)
            cache[seed] = (B200VAEDecoder(VO.TINY_VAE, sd, device=cuda_device), sd)
        return cache[seed]

    return get


@pytest.mark.parametrize("name", ["vae_tiny_h8", "vae_tiny_h16x8"])
def test_vae_matches_reference_fixture(cuda_device, tiny_vaes, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    vae, _ = tiny_vaes(int(g["wseed"]))
    z = torch.randn(int(g["N"]), 4, int(g["H"]), int(g["W"]), generator=torch.Generator().manual_seed(int(g["zseed"]))) * 0.8
    y = vae.decode_first_stage(z.to(cuda_device)).cpu()
    ref = torch.from_numpy(g["out"])
    assert y.shape == ref.shape and y.dtype == torch.float32
    err, p = O.max_rel_err(y, ref), O.psnr(y, ref)
    print(f"{name}: max-rel {err:.3e} PSNR {p:.1f} dB")
    assert err < TOL and p >= 40.0


@pytest.mark.parametrize("N,H,W,batch", [(3, 16, 16, 2), (1, 32, 32, 1), (5, 8, 8, 4), (1, 64, 32, 1)])
def test_vae_matches_oracle(cuda_device, tiny_vaes, N, H, W, batch):
    # 32x32 / 64x32 latents: 256- and 512-pixel-wide feature maps (image rows wider than a 128-pixel tile)
    vae, sd = tiny_vaes(0)
    z = torch.randn(N, 4, H, W, generator=torch.Generator().manual_seed(N * 100 + H)) * 0.8
    ref = VO.vae_decode(sd, VO.TINY_VAE, z)
    y = vae.decode_first_stage(z.to(cuda_device), batch=batch).cpu()
    err, p = O.max_rel_err(y, ref), O.psnr(y, ref)
    print(f"N{N} {H}x{W}: max-rel {err:.3e} PSNR {p:.1f} dB")
    assert y.shape == (N, 3, 8 * H, 8 * W) and err < TOL and p >= 40.0
    # uint8 images as written by the reference (utils.py:134-137): off by at most one level almost everywhere
    a, b = VO.to_uint8_bgr(y).int(), VO.to_uint8_bgr(ref).int()
    assert float(((a - b).abs() <= 2).float().mean()) > 0.999


def test_vae_call_conventions(cuda_device, tiny_vaes):
    vae, sd = tiny_vaes(0)
    z = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(3))
    y4 = vae.decode_first_stage(z.to(cuda_device))
    y5 = vae.decode_first_stage(z[None].to(cuda_device))       # the reference passes [1, n, 4, h, w] (utils.py:133)
    assert y5.shape == (1, 2, 3, 64, 64) and torch.equal(y5[0], y4)
    assert torch.equal(vae.decode_first_stage(z.to(cuda_device)), y4)  # fixed plan, fixed reduction order
    assert vae.decode_first_stage(z).device.type == "cpu"       # comes back on the caller's device
    with pytest.raises(ValueError):
        vae.decode_first_stage(torch.zeros(1, 3, 8, 8))


def test_vae_uint8_output_is_the_reference_conversion(cuda_device, tiny_vaes):
    """decode_to_uint8_bgr == the reference's host-side conversion (utils.py:134-137) applied to this decoder's
    own fp32 output, bit for bit; and within one or two levels of the oracle's image."""
    vae, sd = tiny_vaes(0)
    z = torch.randn(3, 4, 16, 16, generator=torch.Generator().manual_seed(9)) * 1.5   # some pixels clip
    u8 = vae.decode_to_uint8_bgr(z, batch=2)
    f32 = vae.decode_first_stage(z.to(cuda_device), batch=2).cpu()
    assert u8.shape == (3, 128, 128, 3) and u8.dtype == torch.uint8
    assert torch.equal(u8, VO.to_uint8_bgr(f32))
    ref = VO.to_uint8_bgr(VO.vae_decode(sd, VO.TINY_VAE, z))
    assert float(((u8.int() - ref.int()).abs() <= 2).float().mean()) > 0.999


def test_vae_production_config(cuda_device):
    """first_stage_config of cap4d_mmdm_final.yaml (ch 128, 49.5 M decoder parameters): one 64x64 latent -> 512x512.
    Checker: the oracle evaluated on the GPU in fp32 (TF32 off)."""
    from cap4d_b200 import B200VAEDecoder

    cfg = VO.PRODUCTION_VAE
    sd = VO.init_vae_state_dict(cfg, seed=0)
    vae = B200VAEDecoder(cfg, sd, device=cuda_device)
    z = torch.randn(2, 4, 64, 64, generator=torch.Generator().manual_seed(11)) * 0.8
    y = vae.decode_first_stage(z.to(cuda_device), batch=2)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ref = VO.vae_decode({k: v.to(cuda_device) for k, v in sd.items()}, cfg, z.to(cuda_device))
    err, p = O.max_rel_err(y.cpu(), ref.cpu()), O.psnr(y.cpu(), ref.cpu())
    print(f"production VAE decode 2 x 512x512: max-rel {err:.3e} PSNR {p:.1f} dB, {vae.num_launches()} launches")
    assert y.shape == (2, 3, 512, 512) and err < TOL and p >= 40.0


def test_convert_and_save_latent_images_end_to_end(cuda_device, tiny_vaes, tmp_path):
    """latents -> images/%05d.png through the reference-named writer; the files hold exactly the
    uint8 arrays of decode_to_uint8_bgr (batched decode, threaded encode) in latent order."""
    from cap4d_b200 import output as OUT

    vae, _ = tiny_vaes(0)
    z = torch.randn(5, 4, 8, 8, generator=torch.Generator().manual_seed(21)) * 0.8
    want = vae.decode_to_uint8_bgr(z, batch=2).numpy()
    assert OUT.convert_and_save_latent_images(z, vae, cuda_device, tmp_path, batch=2, writers=2) == 5
    assert sorted(os.listdir(tmp_path / "images")) == [f"{i:05d}.png" for i in range(5)]
    assert np.array_equal(OUT.read_output_images(tmp_path)[..., ::-1], want)


# ---- encoder half (AutoencoderKL.encode, autoencoder.py:82-85) -------------------------------------------------
ENC_TOL = 2e-2  # max|a-b| / max|b| over the posterior parameters (mean | logvar), bf16 operands


@pytest.fixture(scope="module")
def tiny_full_vaes(cuda_device):
    from cap4d_b200 import B200VAEDecoder

    cache = {}

    def get(seed):
        if seed not in cache:
            sd = VO.init_vae_state_dict(VO.TINY_VAE, seed=seed)
            sd.update(VO.init_vae_encoder_state_dict(VO.TINY_VAE, seed=seed))
            cache[seed] = (B200VAEDecoder(VO.TINY_VAE, sd, device=cuda_device), sd)
        return cache[seed]

    return get


@pytest.mark.parametrize("name", ["vae_enc_tiny_64", "vae_enc_tiny_64x128"])
def test_vae_encode_matches_reference_fixture(cuda_device, tiny_full_vaes, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    vae, _ = tiny_full_vaes(int(g["wseed"]))
    assert vae.has_encoder
    x = torch.tanh(torch.randn(int(g["N"]), 3, int(g["H"]), int(g["W"]),
                               generator=torch.Generator().manual_seed(int(g["xseed"]))))
    post = vae.encode(x.to(cuda_device))
    want = torch.from_numpy(g["moments"])
    assert post.parameters.shape == want.shape
    err = O.max_rel_err(post.parameters.cpu(), want)
    assert err <= ENC_TOL, f"{name}: encode parity {err:.3e}"
    assert O.max_rel_err(post.mode().cpu(), torch.from_numpy(g["mode"])) <= ENC_TOL
    torch.manual_seed(int(g["xseed"]) + 1)  # the reference draws the noise from the CPU generator
    assert O.max_rel_err(post.sample().cpu(), torch.from_numpy(g["sample"])) <= ENC_TOL
    # the decode path of the same handle is unaffected by the encoder weights
    z = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(1)) * 0.8
    _, sd = tiny_full_vaes(int(g["wseed"]))
    assert O.max_rel_err(vae.decode_first_stage(z.to(cuda_device)).cpu(), VO.vae_decode(sd, VO.TINY_VAE, z)) <= TOL


@pytest.mark.parametrize("N,H,W", [(3, 64, 64), (1, 128, 64), (1, 512, 512)])
def test_vae_encode_matches_oracle(cuda_device, tiny_full_vaes, N, H, W):
    """512 x 512 exercises the stride-2 conv on rows wider than one 128-pixel tile (the production 512 -> 256 step)."""
    vae, sd = tiny_full_vaes(7)
    x = torch.tanh(torch.randn(N, 3, H, W, generator=torch.Generator().manual_seed(H + W)))
    got = vae.encode_moments(x.to(cuda_device), batch=2).cpu()
    want = VO.vae_encode_moments(sd, VO.TINY_VAE, x)
    err = O.max_rel_err(got, want)
    assert err <= ENC_TOL, f"encode parity {err:.3e}"
    # encode_first_stage: scale_factor * sample, [B, T, ...] accepted like MMLDM.get_input (mmdm.py:60-63)
    torch.manual_seed(3)
    z = vae.encode_first_stage(x[None].to(cuda_device))
    torch.manual_seed(3)
    want_z = VO.SCALE_FACTOR * VO.posterior_sample(want, torch.randn(N, 4, H // 8, W // 8))
    assert z.shape == (1, N, 4, H // 8, W // 8) and O.max_rel_err(z[0].cpu(), want_z) <= ENC_TOL


def test_vae_without_encoder_weights_refuses_to_encode(cuda_device, tiny_vaes):
    vae, _ = tiny_vaes(0)
    assert not vae.has_encoder
    with pytest.raises(RuntimeError, match="no encoder weights"):
        vae.encode_moments(torch.zeros(1, 3, 64, 64, device=cuda_device))
