"""CPU oracle for the MMDM multi-view denoising hot path.  TEST INFRASTRUCTURE ONLY.

A plain, functional PyTorch (fp32) restatement of what the reference computes on this path.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / reference arm may import it;
the product (`cap4d_b200/`) never does.

Parity pinning: the reference ships no tests, golden vectors or checkpoints for this path
(SURVEY.md section 4), so the oracle is pinned against the reference's own code executed in the
authoring container: `oracle/make_golden.py` imports /root/reference (with two import stubs), loads
the SAME seeded weights and inputs into the reference modules and stores their outputs under
`tests/golden/`; `tests/test_oracle_golden.py` checks this file against those fixtures.

Each function cites the reference lines it restates (paths relative to the reference root).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# configs/mmdm/cap4d_mmdm_final.yaml:95-115
PRODUCTION_CONFIG = dict(
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

# small configuration with the same topology (every block kind, spatial + "3d" attention,
# concat-skip group norms whose groups straddle the seam), cheap enough for CPU tests
TINY_CONFIG = dict(
    in_channels=4,
    out_channels=4,
    model_channels=64,
    condition_channels=50,
    num_res_blocks=2,
    channel_mult=(1, 2, 4, 4),
    attention_resolutions=(4, 2, 1),
    num_head_channels=64,
    time_steps=4,
)


# ---------------------------------------------------------------------------------------------
# topology: controlnet/ldm/modules/diffusionmodules/openaimodel.py:544-774 with the MMDM overrides
# of cap4d/mmdm/net/mmdm_unet.py:35-65
# ---------------------------------------------------------------------------------------------
def unet_topology(cfg: dict):
    """Returns (input_blocks, middle, output_blocks); each block is a list of layer tuples:
    ("conv_in", cin, cout) | ("res", prefix, cin, cout) | ("tf", prefix, ch, is3d) |
    ("down", prefix, ch) | ("up", prefix, ch)."""
    mc = cfg["model_channels"]
    mults = list(cfg["channel_mult"])
    nrb = cfg["num_res_blocks"]
    attn = set(cfg["attention_resolutions"])
    input_blocks = [[("conv_in", cfg["in_channels"], mc)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, mult in enumerate(mults):
        for _ in range(nrb):
            i = len(input_blocks)
            blk = [("res", f"input_blocks.{i}.0.", ch, mult * mc)]
            ch = mult * mc
            if ds in attn:
                blk.append(("tf", f"input_blocks.{i}.1.", ch, mult >= 2))  # mmdm_unet.py:49-55
            input_blocks.append(blk)
            chans.append(ch)
        if level != len(mults) - 1:
            i = len(input_blocks)
            input_blocks.append([("down", f"input_blocks.{i}.0.op.", ch)])
            chans.append(ch)
            ds *= 2
    middle = [
        ("res", "middle_block.0.", ch, ch),
        ("tf", "middle_block.1.", ch, mults[-1] >= 2),
        ("res", "middle_block.2.", ch, ch),
    ]
    output_blocks = []
    for level, mult in list(enumerate(mults))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            o = len(output_blocks)
            blk = [("res", f"output_blocks.{o}.0.", ch + ich, mc * mult)]
            ch = mc * mult
            sub = 1
            if ds in attn:
                blk.append(("tf", f"output_blocks.{o}.1.", ch, mult >= 2))
                sub = 2
            if level and i == nrb:
                blk.append(("up", f"output_blocks.{o}.{sub}.conv.", ch))
                ds //= 2
            output_blocks.append(blk)
    return input_blocks, middle, output_blocks


def unet_param_shapes(cfg: dict) -> "OrderedDict[str, Tuple[int, ...]]":
    """state_dict keys and shapes of MMDMUnetModel for `cfg` (names as in the reference modules)."""
    mc = cfg["model_channels"]
    emb = 4 * mc
    shapes: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    shapes["time_embed.0.weight"] = (emb, mc)
    shapes["time_embed.0.bias"] = (emb,)
    shapes["time_embed.2.weight"] = (emb, emb)
    shapes["time_embed.2.bias"] = (emb,)

    def res(p, cin, cout):
        shapes[p + "in_layers.0.weight"] = (cin,)
        shapes[p + "in_layers.0.bias"] = (cin,)
        shapes[p + "in_layers.2.weight"] = (cout, cin, 3, 3)
        shapes[p + "in_layers.2.bias"] = (cout,)
        shapes[p + "emb_layers.1.weight"] = (cout, emb)
        shapes[p + "emb_layers.1.bias"] = (cout,)
        shapes[p + "out_layers.0.weight"] = (cout,)
        shapes[p + "out_layers.0.bias"] = (cout,)
        shapes[p + "out_layers.3.weight"] = (cout, cout, 3, 3)
        shapes[p + "out_layers.3.bias"] = (cout,)
        if cin != cout:
            shapes[p + "skip_connection.weight"] = (cout, cin, 1, 1)
            shapes[p + "skip_connection.bias"] = (cout,)

    def tf(p, c):
        shapes[p + "norm.weight"] = (c,)
        shapes[p + "norm.bias"] = (c,)
        shapes[p + "proj_in.weight"] = (c, c)
        shapes[p + "proj_in.bias"] = (c,)
        t = p + "transformer_blocks.0."
        shapes[t + "attn1.to_q.weight"] = (c, c)
        shapes[t + "attn1.to_k.weight"] = (c, c)
        shapes[t + "attn1.to_v.weight"] = (c, c)
        shapes[t + "attn1.to_out.0.weight"] = (c, c)
        shapes[t + "attn1.to_out.0.bias"] = (c,)
        shapes[t + "norm1.weight"] = (c,)
        shapes[t + "norm1.bias"] = (c,)
        shapes[t + "norm3.weight"] = (c,)
        shapes[t + "norm3.bias"] = (c,)
        shapes[t + "ff.net.0.proj.weight"] = (8 * c, c)
        shapes[t + "ff.net.0.proj.bias"] = (8 * c,)
        shapes[t + "ff.net.2.weight"] = (c, 4 * c)
        shapes[t + "ff.net.2.bias"] = (c,)
        shapes[p + "proj_out.weight"] = (c, c)
        shapes[p + "proj_out.bias"] = (c,)

    def conv(p, c):
        shapes[p + "weight"] = (c, c, 3, 3)
        shapes[p + "bias"] = (c,)

    ib, mid, ob = unet_topology(cfg)
    for blk in ib + [mid] + ob:
        for layer in blk:
            if layer[0] == "conv_in":
                shapes["input_blocks.0.0.weight"] = (layer[2], layer[1], 3, 3)
                shapes["input_blocks.0.0.bias"] = (layer[2],)
            elif layer[0] == "res":
                res(layer[1], layer[2], layer[3])
            elif layer[0] == "tf":
                tf(layer[1], layer[2])
            else:
                conv(layer[1], layer[2])
    shapes["out.0.weight"] = (mc,)
    shapes["out.0.bias"] = (mc,)
    shapes["out.2.weight"] = (cfg["out_channels"], mc, 3, 3)
    shapes["out.2.bias"] = (cfg["out_channels"],)
    shapes["cond_linear.weight"] = (mc, cfg["condition_channels"])
    shapes["cond_linear.bias"] = (mc,)
    return shapes


# parameters the reference constructor zero-initialises (SURVEY.md note Z): openaimodel.py:230-232,
# :773; attention.py:176, :371; mmdm_unet.py:33
def _is_zero_module(name: str) -> bool:
    return (
        ".out_layers.3." in name
        or ".proj_out." in name
        or ".attn1.to_out.0." in name
        or name.startswith("out.2.")
        or name.startswith("cond_linear.")
    )


def init_state_dict(cfg: dict, seed: int = 0, zero_std: float = 0.02) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic random-init weights (fp32, CPU): PyTorch-default-like scale for ordinary layers,
    N(0, zero_std) for the modules the reference zero-initialises (a freshly constructed reference
    model would output exactly 0 on generated views), perturbed affine parameters for the norms."""
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in unet_param_shapes(cfg).items():
        is_norm = (
            ".in_layers.0." in name
            or ".out_layers.0." in name
            or name.startswith("out.0.")
            or ".norm." in name
            or ".norm1." in name
            or ".norm3." in name
        )
        if is_norm:
            base = 1.0 if name.endswith("weight") else 0.0
            t = base + 0.1 * torch.randn(shape, generator=g)
        elif _is_zero_module(name):
            t = zero_std * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = int(np.prod(shape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[name] = t.float()
    return sd


def make_inputs(cfg: dict, B: int, V: int, H: int, W: int, R: int, seed: int = 0, timestep: int = 501):
    """Seeded synthetic inputs: latents ~ N(0,1), ref_mask = 1 on the R leading views,
    pos_enc ~ N(0,1) on the conditional half and zero on the unconditional half (b = 0) with
    z_input = 0 there, as CAP4DConditioning(unconditional=True) produces (cap4dcond.py:78-88)."""
    g = torch.Generator().manual_seed(seed)
    C, cc = cfg["in_channels"], cfg["condition_channels"]
    x = torch.randn(B, V, C, H, W, generator=g)
    z = torch.randn(B, V, C, H, W, generator=g)
    pos = torch.randn(B, V, H, W, cc, generator=g)
    mask = torch.zeros(B, V, 1, H, W)
    mask[:, :R] = 1.0
    if B >= 2:
        half = B // 2
        pos[:half] = 0.0
        z[:half] = 0.0
    t = torch.full((B, V), timestep, dtype=torch.long)
    return x, t, dict(z_input=z, ref_mask=mask, pos_enc=pos)


def make_sampler_conditioning(cfg, n_ref, n_gen, H, W, seed):
    """Seeded cond/uncond dicts shaped like get_condition_from_dataloader's output
    (cap4d/inference/utils.py:64-100): uncond = zero pos_enc / zero z_input, same ref_mask."""
    g = torch.Generator().manual_seed(seed)
    C, cc = cfg["in_channels"], cfg["condition_channels"]

    def mk(n, is_ref):
        cond = dict(
            z_input=torch.randn(n, C, H, W, generator=g) if is_ref else torch.zeros(n, C, H, W),
            ref_mask=torch.full((n, 1, H, W), 1.0 if is_ref else 0.0),
            pos_enc=torch.randn(n, H, W, cc, generator=g),
        )
        unc = dict(z_input=cond["z_input"] * 0.0, ref_mask=cond["ref_mask"].clone(), pos_enc=cond["pos_enc"] * 0.0)
        return cond, unc

    ref_cond, ref_unc = mk(n_ref, True)
    gen_cond, gen_unc = mk(n_gen, False)
    return ref_cond, ref_unc, gen_cond, gen_unc



# ---------------------------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------------------------
def timestep_embedding(t: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    """controlnet/ldm/modules/diffusionmodules/util.py:154-174 (repeat_only=False)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _gn(x, sd, p, eps):
    # GroupNorm32: fp32 group norm with 32 groups (util.py:217-219, :202-208)
    return F.group_norm(_c(x), 32, sd[p + "weight"], sd[p + "bias"], eps)


def res_block(sd, p, x, emb):
    """ResBlock._forward without up/down (openaimodel.py:256-276)."""
    h = F.conv2d(F.silu(_gn(x, sd, p + "in_layers.0.", 1e-5)), sd[p + "in_layers.2.weight"], sd[p + "in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"])
    h = h + e[:, :, None, None]
    h = F.conv2d(F.silu(_gn(h, sd, p + "out_layers.0.", 1e-5)), sd[p + "out_layers.3.weight"], sd[p + "out_layers.3.bias"], padding=1)
    if (p + "skip_connection.weight") in sd:
        x = F.conv2d(x, sd[p + "skip_connection.weight"], sd[p + "skip_connection.bias"])
    return x + h


# Arithmetic type of unet_forward's internal casts (the reference's `.float()` / `.type(self.dtype)` calls).  float32 is
# the reference; float64 - with a float64 state_dict and inputs - gives the ground truth that tells the rounding noise of
# two correct fp32 implementations apart from a real error (tests/test_gpu_fp32_mode.py).
COMPUTE_DTYPE = torch.float32


def _c(x: torch.Tensor) -> torch.Tensor:
    return x.to(COMPUTE_DTYPE)


# "legacy": the reference's materialised softmax (what the parity tests check against).  "sdpa": the same contraction
# through torch's fused scaled_dot_product_attention - only for bench.py's gpu_eager_baseline leg (the best stock
# PyTorch can do on the same GPU); never used as a checker.
ATTENTION_IMPL = "legacy"


def attention_core(q, k, v, heads: int, views: int, is3d: bool):
    """AttentionModule.forward 'NORMAL ATTENTION' branch + legacy_attention
    (cap4d/mmdm/net/attention.py:229-251, :112-132): softmax(q k^T / sqrt(d)) v, fp32."""
    bt, n, c = q.shape
    d = c // heads
    if is3d:
        b = bt // views

        def split(y):  # '(b t) n (h d) -> (b h) (n t) d'
            return y.reshape(b, views, n, heads, d).permute(0, 3, 2, 1, 4).reshape(b * heads, n * views, d)

    else:

        def split(y):  # 'b n (h d) -> (b h) n d'
            return y.reshape(bt, n, heads, d).permute(0, 2, 1, 3).reshape(bt * heads, n, d)

    q, k, v = split(q), split(k), split(v)
    if ATTENTION_IMPL == "sdpa":
        out = F.scaled_dot_product_attention(q[None], k[None], v[None], scale=d ** -0.5)[0]
    else:
        sim = torch.einsum("bid,bjd->bij", _c(q), _c(k)) * (d ** -0.5)
        sim = sim.softmax(dim=-1)
        out = torch.einsum("bij,bjd->bid", sim, v)
    if is3d:
        out = out.reshape(b, heads, n, views, d).permute(0, 3, 2, 1, 4).reshape(bt, n, c)
    else:
        out = out.reshape(bt, heads, n, d).permute(0, 2, 1, 3).reshape(bt, n, c)
    return out


def transformer(sd, p, x, views: int, is3d: bool, head_dim: int = 64):
    """SpatioTemporalTransformer.forward (attention.py:375-387) with one BasicTransformerBlock
    (:311-326; use_context=False, no temporal attention) and GEGLU feed-forward (:68-95)."""
    n, c, h, w = x.shape
    heads = c // head_dim
    t = p + "transformer_blocks.0."
    y = _gn(x, sd, p + "norm.", 1e-6)
    y = y.permute(0, 2, 3, 1).reshape(n, h * w, c)
    y = F.linear(y, sd[p + "proj_in.weight"], sd[p + "proj_in.bias"])
    z = F.layer_norm(y, (c,), sd[t + "norm1.weight"], sd[t + "norm1.bias"], 1e-5)
    q = F.linear(z, sd[t + "attn1.to_q.weight"])
    k = F.linear(z, sd[t + "attn1.to_k.weight"])
    v = F.linear(z, sd[t + "attn1.to_v.weight"])
    a = attention_core(q, k, v, heads, views, is3d)
    y = F.linear(a, sd[t + "attn1.to_out.0.weight"], sd[t + "attn1.to_out.0.bias"]) + y
    z = F.layer_norm(y, (c,), sd[t + "norm3.weight"], sd[t + "norm3.bias"], 1e-5)
    u = F.linear(z, sd[t + "ff.net.0.proj.weight"], sd[t + "ff.net.0.proj.bias"])
    a_, gate = u.chunk(2, dim=-1)
    u = a_ * F.gelu(gate)
    y = F.linear(u, sd[t + "ff.net.2.weight"], sd[t + "ff.net.2.bias"]) + y
    y = F.linear(y, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    y = y.reshape(n, h, w, c).permute(0, 3, 1, 2)
    return y + x


def _run_block(sd, blk, h, emb, views):
    for layer in blk:
        kind = layer[0]
        if kind == "conv_in":
            h = F.conv2d(h, sd["input_blocks.0.0.weight"], sd["input_blocks.0.0.bias"], padding=1)
        elif kind == "res":
            h = res_block(sd, layer[1], h, emb)
        elif kind == "tf":
            h = transformer(sd, layer[1], h, views, layer[3])
        elif kind == "down":  # Downsample: conv3x3 stride 2 (openaimodel.py:150-153)
            h = F.conv2d(h, sd[layer[1] + "weight"], sd[layer[1] + "bias"], stride=2, padding=1)
        elif kind == "up":  # Upsample: nearest 2x then conv3x3 (openaimodel.py:111-119)
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[layer[1] + "weight"], sd[layer[1] + "bias"], padding=1)
    return h


@torch.no_grad()
def unet_forward(sd: Dict[str, torch.Tensor], cfg: dict, x: torch.Tensor, timesteps: torch.Tensor,
                 control: Dict[str, torch.Tensor], taps: Optional[dict] = None) -> torch.Tensor:
    """MMDMUnetModel.forward (cap4d/mmdm/net/mmdm_unet.py:67-126).  x: [B,V,C,H,W]."""
    z = control["z_input"]
    mask = control["ref_mask"]
    x_input = x - z
    inv = torch.logical_not(mask)
    x = z * mask + x * inv
    B, V = x.shape[:2]
    h = x.reshape(B * V, *x.shape[2:])
    t = timesteps.reshape(B * V)
    pos = _c(control["pos_enc"].reshape(B * V, *control["pos_enc"].shape[2:]))
    pos_emb = F.linear(pos, sd["cond_linear.weight"], sd["cond_linear.bias"]).permute(0, 3, 1, 2)
    temb = _c(timestep_embedding(t, cfg["model_channels"]))
    emb = F.linear(F.silu(F.linear(temb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])),
                   sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    ib, mid, ob = unet_topology(cfg)
    hs = []
    h = _c(h)
    for i, blk in enumerate(ib):
        h = _run_block(sd, blk, h, emb, V)
        if i == 0:
            h = h + pos_emb
        hs.append(h)
        if taps is not None:
            taps[f"input_blocks.{i}"] = h
    h = _run_block(sd, mid, h, emb, V)
    if taps is not None:
        taps["middle_block"] = h
    for i, blk in enumerate(ob):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_block(sd, blk, h, emb, V)
        if taps is not None:
            taps[f"output_blocks.{i}"] = h
    h = F.conv2d(F.silu(_gn(h, sd, "out.0.", 1e-5)), sd["out.2.weight"], sd["out.2.bias"], padding=1)
    h = h.reshape(B, V, *h.shape[1:])
    return x_input * mask + h * inv


# ---------------------------------------------------------------------------------------------
# noise schedule: cap4d/mmdm/mmdm.py:276-324, cap4d/mmdm/utils.py:4-37, util.py:21-25
# ---------------------------------------------------------------------------------------------
def mmdm_schedule(timesteps: int = 1000, linear_start: float = 0.00085, linear_end: float = 0.0120,
                  n_frames: int = 8, image_size: int = 64, zero_snr_shift: bool = True, shift_schedule: bool = True,
                  sqrt_shift: bool = True, minus_one_shift: bool = True):
    """Returns float64 numpy (betas, alphas_cumprod, alphas_cumprod_prev) exactly as MMLDM.register_schedule
    derives them before its float32 cast (defaults = configs/mmdm/cap4d_mmdm_final.yaml:73-93)."""
    betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2).numpy()
    if zero_snr_shift:  # enforce_zero_terminal_snr, utils.py:18-37
        ab_sqrt = np.sqrt((1 - betas).cumprod(0))
        a0, aT = ab_sqrt[0].copy(), ab_sqrt[-1].copy()
        ab_sqrt = (ab_sqrt - aT) * (a0 / (a0 - aT))
        ab = ab_sqrt ** 2
        alphas = np.concatenate([ab[0:1], ab[1:] / ab[:-1]])
        betas = 1 - alphas
    betas[betas > 0.99] = 0.99
    alphas_cumprod = np.cumprod(1.0 - betas, axis=0)
    if shift_schedule:  # mmdm.py:293-308 + utils.py:4-14
        n_gen = n_frames - 1 if minus_one_shift else n_frames
        ratio = (64 ** 2) / (image_size ** 2 * n_gen)
        if sqrt_shift:
            ratio = np.sqrt(ratio)
        snr = alphas_cumprod / (1.0 - alphas_cumprod)
        log_snr = np.log(snr) + np.log(ratio)
        a_shift = np.exp(log_snr) / (1 + np.exp(log_snr))
        betas = 1 - np.concatenate([[1], a_shift[1:] / a_shift[:-1]])
        alphas_cumprod = a_shift
    alphas_cumprod_prev = np.append(1.0, alphas_cumprod[:-1])
    return betas, alphas_cumprod, alphas_cumprod_prev


def ddim_schedule(alphas_cumprod_f32: np.ndarray, S: int, num_ddpm: int = 1000, eta: float = 0.0):
    """make_ddim_timesteps('uniform') + make_ddim_sampling_parameters (util.py:46-74) as called from
    StochasticIOSampler.make_schedule (cap4d/mmdm/sampler.py:32-61) on the model's fp32 buffers."""
    c = num_ddpm // S
    ts = np.asarray(list(range(0, num_ddpm, c))) + 1
    ac = torch.as_tensor(alphas_cumprod_f32, dtype=torch.float32)
    alphas = ac[ts]
    alphas_prev = np.asarray([ac[0]] + ac[ts[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return ts, alphas, alphas_prev, sigmas


def ddim_coefficients(alphas, alphas_prev, sigmas, index: int):
    """sampler.py:215-229: float64 intermediate, float32 factors."""
    alpha_t = alphas.float()[index].double()
    sqrt_one_minus = np.sqrt(1.0 - alphas)[index].double()
    sigma_t = sigmas[index]
    alpha_prev = torch.tensor(alphas_prev).float()[index].double()
    e_f = -alpha_prev.sqrt() * sqrt_one_minus / alpha_t.sqrt() + (1.0 - alpha_prev - sigma_t ** 2).sqrt()
    x_f = alpha_prev.sqrt() / alpha_t.sqrt()
    return x_f.float(), e_f.float()


@torch.no_grad()
def stochastic_io_sample(eps_fn, alphas_cumprod_f32, S, ref_cond, ref_uncond, gen_cond, gen_uncond, latent_shape,
                         V=8, R_max=4, cfg_scale=1.0, eta=0.0, num_ddpm=1000):
    """StochasticIOSampler.sample (cap4d/mmdm/sampler.py:64-233), single device.  Consumes the global
    torch CPU RNG (x_T) and numpy RNG (permutations) in the reference's order.
    eps_fn(x[2,V,C,h,w], t[2,V], control) -> eps[2,V,C,h,w]."""
    ts, alphas, alphas_prev, sigmas = ddim_schedule(alphas_cumprod_f32, S, num_ddpm, eta)
    n_gen = gen_cond["z_input"].shape[0]
    n_ref = ref_cond["z_input"].shape[0]
    R = min(n_ref, R_max)
    G = V - R
    assert n_gen % G == 0
    n_its = n_gen // G
    x_all = torch.randn((n_gen, *latent_shape))
    total = ts.shape[0]
    for i, step in enumerate(np.flip(ts)):
        index = total - i - 1
        t_in = torch.full((2, V), int(step), dtype=torch.long)
        e_all = torch.zeros_like(x_all)
        if R == 1:
            ref_batches = np.zeros((n_its, R), dtype=np.int64)
        else:
            ref_batches = np.stack([np.random.permutation(np.arange(n_ref))[:R] for _ in range(n_its)], axis=0)
        gen_batches = np.reshape(np.random.permutation(np.arange(n_gen)), (n_its, -1))
        for b in range(n_its):
            control = {}
            for key in ref_cond:
                cond = torch.cat([ref_cond[key][ref_batches[b]], gen_cond[key][gen_batches[b]]], dim=0)[None]
                unc = torch.cat([ref_uncond[key][ref_batches[b]], gen_uncond[key][gen_batches[b]]], dim=0)[None]
                control[key] = torch.cat([unc, cond], dim=0)
            x_in = torch.cat([ref_cond["z_input"][ref_batches[b]], x_all[gen_batches[b]]], dim=0)[None]
            x_in = torch.cat([x_in] * 2, dim=0)
            eps_u, eps_c = eps_fn(x_in, t_in, control).chunk(2)
            e = eps_u + cfg_scale * (eps_c - eps_u)
            e_all[gen_batches[b]] += e[0, R:]
        x_f, e_f = ddim_coefficients(alphas, alphas_prev, sigmas, index)
        x_all = x_all * x_f + e_all * e_f
    return x_all


# ---------------------------------------------------------------------------------------------
# error metrics used by the parity tests
# ---------------------------------------------------------------------------------------------
def max_rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b| (pointwise relative error is meaningless where eps crosses 0)."""
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def psnr(a: torch.Tensor, b: torch.Tensor) -> float:
    peak = float(b.max() - b.min())
    mse = float(((a.double() - b.double()) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * math.log10(peak * peak / mse)
