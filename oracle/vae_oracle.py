"""CPU restatement of the reference's VAE decode.  TEST INFRASTRUCTURE ONLY: imported by
tests/, never by the product (cap4d_b200/).

Follows, function by function:
  AutoencoderKL.decode                 controlnet/ldm/models/autoencoder.py:87-91  (post_quant_conv, decoder)
  DDPM-side scaling                    controlnet/ldm/models/diffusion/ddpm.py:822-830 (z / scale_factor)
  Decoder.__init__ / forward           controlnet/ldm/modules/diffusionmodules/model.py:546-640
  ResnetBlock.forward (temb = None)    model.py:125-146
  AttnBlock.forward                    model.py:176-203   (single head over all channels, scale c^-0.5)
  Upsample.forward                     model.py:71-76     (nearest 2x, conv3x3)
  Normalize                            model.py:46-47     (GroupNorm 32 groups, eps 1e-6)
  nonlinearity                         model.py:41-43     (x * sigmoid(x))

Pinned against the unmodified reference module by oracle/make_golden.py -> tests/golden/vae_*.npz.
"""
from collections import OrderedDict
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

# first_stage_config of configs/mmdm/cap4d_mmdm_final.yaml:117-137
PRODUCTION_VAE = dict(ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4, out_ch=3, embed_dim=4)
# same topology, half the channels (the conv kernel's K blocks need input channels in multiples of 64)
TINY_VAE = dict(ch=64, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4, out_ch=3, embed_dim=4)
SCALE_FACTOR = 0.18215  # cap4d_mmdm_final.yaml: scale_factor


def vae_decoder_topology(cfg: dict):
    """[(kind, prefix, cin, cout)] in execution order; kinds: conv_in, res, attn, up, out."""
    ch, mult, nrb = cfg["ch"], tuple(cfg["ch_mult"]), cfg["num_res_blocks"]
    nres = len(mult)
    block_in = ch * mult[nres - 1]
    layers = [("conv_in", "decoder.conv_in", cfg["z_channels"], block_in),
              ("res", "decoder.mid.block_1", block_in, block_in),
              ("attn", "decoder.mid.attn_1", block_in, block_in),
              ("res", "decoder.mid.block_2", block_in, block_in)]
    for lvl in reversed(range(nres)):
        block_out = ch * mult[lvl]
        for i in range(nrb + 1):
            layers.append(("res", f"decoder.up.{lvl}.block.{i}", block_in, block_out))
            block_in = block_out
        if lvl != 0:
            layers.append(("up", f"decoder.up.{lvl}.upsample.conv", block_in, block_in))
    layers.append(("out", "decoder", block_in, cfg["out_ch"]))
    return layers


def vae_param_shapes(cfg: dict) -> "OrderedDict[str, Tuple[int, ...]]":
    shapes: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    e = cfg["embed_dim"]
    shapes["post_quant_conv.weight"] = (cfg["z_channels"], e, 1, 1)
    shapes["post_quant_conv.bias"] = (cfg["z_channels"],)
    for kind, p, cin, cout in vae_decoder_topology(cfg):
        if kind in ("conv_in", "up"):
            shapes[p + ".weight"] = (cout, cin, 3, 3)
            shapes[p + ".bias"] = (cout,)
        elif kind == "res":
            shapes[p + ".norm1.weight"] = (cin,)
            shapes[p + ".norm1.bias"] = (cin,)
            shapes[p + ".conv1.weight"] = (cout, cin, 3, 3)
            shapes[p + ".conv1.bias"] = (cout,)
            shapes[p + ".norm2.weight"] = (cout,)
            shapes[p + ".norm2.bias"] = (cout,)
            shapes[p + ".conv2.weight"] = (cout, cout, 3, 3)
            shapes[p + ".conv2.bias"] = (cout,)
            if cin != cout:
                shapes[p + ".nin_shortcut.weight"] = (cout, cin, 1, 1)
                shapes[p + ".nin_shortcut.bias"] = (cout,)
        elif kind == "attn":
            shapes[p + ".norm.weight"] = (cin,)
            shapes[p + ".norm.bias"] = (cin,)
            for n in ("q", "k", "v", "proj_out"):
                shapes[f"{p}.{n}.weight"] = (cin, cin, 1, 1)
                shapes[f"{p}.{n}.bias"] = (cin,)
        elif kind == "out":
            shapes["decoder.norm_out.weight"] = (cin,)
            shapes["decoder.norm_out.bias"] = (cin,)
            shapes["decoder.conv_out.weight"] = (cout, cin, 3, 3)
            shapes["decoder.conv_out.bias"] = (cout,)
    return shapes


def init_vae_state_dict(cfg: dict, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded synthetic weights (there is no checkpoint offline): fan-in scaled normals, norm gains near 1."""
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in vae_param_shapes(cfg).items():
        if "norm" in name and name.endswith(".weight"):
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            sd[name] = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = shape[1] * shape[2] * shape[3]
            sd[name] = torch.randn(shape, generator=g) / fan_in ** 0.5
    return sd


def _gn_swish(x, sd, p, swish=True):
    h = F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)
    return h * torch.sigmoid(h) if swish else h


def _res(sd, p, x):
    h = F.conv2d(_gn_swish(x, sd, p + ".norm1"), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = F.conv2d(_gn_swish(h, sd, p + ".norm2"), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if p + ".nin_shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + ".nin_shortcut.weight"], sd[p + ".nin_shortcut.bias"])
    return x + h


def _attn(sd, p, x):
    h = _gn_swish(x, sd, p + ".norm", swish=False)
    q = F.conv2d(h, sd[p + ".q.weight"], sd[p + ".q.bias"])
    k = F.conv2d(h, sd[p + ".k.weight"], sd[p + ".k.bias"])
    v = F.conv2d(h, sd[p + ".v.weight"], sd[p + ".v.bias"])
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.softmax(torch.bmm(q, k) * (int(c) ** (-0.5)), dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + F.conv2d(h, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])


@torch.no_grad()
def vae_decode(sd: Dict[str, torch.Tensor], cfg: dict, z: torch.Tensor, scale_factor: float = SCALE_FACTOR,
               taps: dict = None) -> torch.Tensor:
    """decode_first_stage: z [N, 4, h, w] (sampler latents) -> images [N, 3, 8h, 8w] in ~[-1, 1]."""
    h = (1.0 / scale_factor) * z
    h = F.conv2d(h, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    for kind, p, cin, cout in vae_decoder_topology(cfg):
        if kind == "conv_in":
            h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
        elif kind == "res":
            h = _res(sd, p, h)
        elif kind == "attn":
            h = _attn(sd, p, h)
        elif kind == "up":
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
        elif kind == "out":
            h = _gn_swish(h, sd, "decoder.norm_out")
            h = F.conv2d(h, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)
        if taps is not None and p in taps:
            taps[p] = h.clone()
    return h


# ---- encoder half: AutoencoderKL.encode (autoencoder.py:82-85) -----------------------------------------------
#   Encoder.__init__ / forward           model.py:452-545   (attn_resolutions = []: mid-block attention only)
#   Downsample.forward                   model.py:80-84     (F.pad (0,1,0,1) then conv3x3 stride 2, padding 0)
#   quant_conv                           autoencoder.py:33,84
#   DiagonalGaussianDistribution         controlnet/ldm/modules/distributions/distributions.py:24-45
# Pinned against the unmodified reference module by oracle/make_golden.py -> tests/golden/vae_enc_*.npz.
def vae_encoder_topology(cfg: dict):
    """[(kind, prefix, cin, cout)] in execution order; kinds: conv_in, res, down, attn, out."""
    ch, mult, nrb = cfg["ch"], tuple(cfg["ch_mult"]), cfg["num_res_blocks"]
    layers = [("conv_in", "encoder.conv_in", cfg["out_ch"], ch)]
    block_in = ch
    for lvl in range(len(mult)):
        block_out = ch * mult[lvl]
        for i in range(nrb):
            layers.append(("res", f"encoder.down.{lvl}.block.{i}", block_in, block_out))
            block_in = block_out
        if lvl != len(mult) - 1:
            layers.append(("down", f"encoder.down.{lvl}.downsample.conv", block_in, block_in))
    layers += [("res", "encoder.mid.block_1", block_in, block_in), ("attn", "encoder.mid.attn_1", block_in, block_in),
               ("res", "encoder.mid.block_2", block_in, block_in), ("out", "encoder", block_in, 2 * cfg["z_channels"])]
    return layers


def vae_encoder_param_shapes(cfg: dict) -> "OrderedDict[str, Tuple[int, ...]]":
    shapes: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    for kind, p, cin, cout in vae_encoder_topology(cfg):
        if kind in ("conv_in", "down"):
            shapes[p + ".weight"] = (cout, cin, 3, 3)
            shapes[p + ".bias"] = (cout,)
        elif kind == "res":
            shapes[p + ".norm1.weight"] = (cin,)
            shapes[p + ".norm1.bias"] = (cin,)
            shapes[p + ".conv1.weight"] = (cout, cin, 3, 3)
            shapes[p + ".conv1.bias"] = (cout,)
            shapes[p + ".norm2.weight"] = (cout,)
            shapes[p + ".norm2.bias"] = (cout,)
            shapes[p + ".conv2.weight"] = (cout, cout, 3, 3)
            shapes[p + ".conv2.bias"] = (cout,)
            if cin != cout:
                shapes[p + ".nin_shortcut.weight"] = (cout, cin, 1, 1)
                shapes[p + ".nin_shortcut.bias"] = (cout,)
        elif kind == "attn":
            shapes[p + ".norm.weight"] = (cin,)
            shapes[p + ".norm.bias"] = (cin,)
            for n in ("q", "k", "v", "proj_out"):
                shapes[f"{p}.{n}.weight"] = (cin, cin, 1, 1)
                shapes[f"{p}.{n}.bias"] = (cin,)
        elif kind == "out":
            shapes["encoder.norm_out.weight"] = (cin,)
            shapes["encoder.norm_out.bias"] = (cin,)
            shapes["encoder.conv_out.weight"] = (cout, cin, 3, 3)
            shapes["encoder.conv_out.bias"] = (cout,)
    shapes["quant_conv.weight"] = (2 * cfg["embed_dim"], 2 * cfg["z_channels"], 1, 1)
    shapes["quant_conv.bias"] = (2 * cfg["embed_dim"],)
    return shapes


def init_vae_encoder_state_dict(cfg: dict, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded synthetic encoder weights, same recipe as init_vae_state_dict."""
    g = torch.Generator().manual_seed(seed + 1000)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in vae_encoder_param_shapes(cfg).items():
        if "norm" in name and name.endswith(".weight"):
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            sd[name] = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = shape[1] * shape[2] * shape[3]
            sd[name] = torch.randn(shape, generator=g) / fan_in ** 0.5
    return sd


@torch.no_grad()
def vae_encode_moments(sd: Dict[str, torch.Tensor], cfg: dict, x: torch.Tensor) -> torch.Tensor:
    """AutoencoderKL.encode(x).parameters: images [N, 3, H, W] in [-1, 1] -> moments [N, 2e, H/f, W/f] (mean | logvar)."""
    h = x
    for kind, p, cin, cout in vae_encoder_topology(cfg):
        if kind == "conv_in":
            h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
        elif kind == "res":
            h = _res(sd, p, h)
        elif kind == "attn":
            h = _attn(sd, p, h)
        elif kind == "down":
            h = F.conv2d(F.pad(h, (0, 1, 0, 1), mode="constant", value=0), sd[p + ".weight"], sd[p + ".bias"], stride=2)
        elif kind == "out":
            h = _gn_swish(h, sd, "encoder.norm_out")
            h = F.conv2d(h, sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"], padding=1)
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])


def posterior_sample(moments: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    """DiagonalGaussianDistribution.sample (distributions.py:24-37): mean + exp(0.5 * clamp(logvar, -30, 20)) * noise."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    return mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * noise


def to_uint8_bgr(x: torch.Tensor) -> torch.Tensor:
    """convert_and_save_latent_images (cap4d/inference/utils.py:131-137): [-1,1] CHW -> uint8 HWC, channels
    reversed for cv2.imwrite.  astype(uint8) truncates."""
    img = ((x + 1.0) / 2.0).clip(0.0, 1.0).permute(0, 2, 3, 1) * 255.0
    return img[..., [2, 1, 0]].to(torch.uint8)
