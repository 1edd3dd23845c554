"""B200 VAE decoder: the reference's `decode_first_stage` (controlnet/ldm/models/diffusion/ddpm.py:822-830 ->
AutoencoderKL.decode, controlnet/ldm/models/autoencoder.py:87-91) behind the same call, on libcap4d_b200.so.

    vae = B200VAEDecoder.from_reference(model.first_stage_model, scale_factor=model.scale_factor)
    images = vae.decode_first_stage(latents)          # [N, 4, h, w] -> [N, 3, 8h, 8w], about [-1, 1]

`cap4d/inference/utils.py:131-137` decodes the generated views one at a time; this decodes `batch` views per
launch plan.  There is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Mapping, Optional

import torch

from . import _lib

# first_stage_config of configs/mmdm/cap4d_mmdm_final.yaml:117-137
VAE_CONFIG = dict(ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4, embed_dim=4, out_ch=3)
SCALE_FACTOR = 0.18215


def _make_config(cfg: Mapping) -> _lib.VaeConfig:
    c = _lib.VaeConfig()
    c.ch = int(cfg["ch"])
    mult = list(cfg["ch_mult"])
    if len(mult) > _lib.MAX_LEVELS:
        raise ValueError("too many levels")
    c.n_levels = len(mult)
    for i, m in enumerate(mult):
        c.ch_mult[i] = int(m)
    c.num_res_blocks = int(cfg["num_res_blocks"])
    c.z_channels = int(cfg["z_channels"])
    c.embed_dim = int(cfg.get("embed_dim", cfg["z_channels"]))
    c.out_ch = int(cfg["out_ch"])
    return c


class B200VAEDecoder(torch.nn.Module):
    def __init__(self, config: Mapping, state_dict: Mapping[str, torch.Tensor], scale_factor: float = SCALE_FACTOR,
                 device: Optional[torch.device] = None):
        super().__init__()
        if not torch.cuda.is_available():
            raise RuntimeError("cap4d_b200: a CUDA device (B200, sm_100a) is required; there is no CPU path")
        self.config = dict(config)
        self.scale_factor = float(scale_factor)
        self._device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._lib = _lib.load()
        self._handle = ctypes.c_void_p()
        self._ws: Dict = {}
        cfg = _make_config(config)
        with torch.cuda.device(self._device):
            _lib.check(self._lib.cap4d_b200_vae_create(ctypes.byref(cfg), ctypes.byref(self._handle)), "vae_create")
            for name, t in state_dict.items():
                if not name.startswith(("decoder.", "post_quant_conv.", "encoder.", "quant_conv.")):
                    continue  # loss / EMA entries of a full AutoencoderKL state_dict
                t32 = t.detach().to(dtype=torch.float32).contiguous()
                shape = (ctypes.c_int64 * max(1, t32.dim()))(*t32.shape)
                _lib.check(self._lib.cap4d_b200_vae_load_weight(self._handle, name.encode(),
                                                               ctypes.c_void_p(t32.data_ptr()), shape, t32.dim()),
                           f"vae_load_weight({name})")
            _lib.check(self._lib.cap4d_b200_vae_finalize(self._handle), "vae_finalize")

    @staticmethod
    def param_shapes(config: Mapping) -> Dict[str, tuple]:
        """state_dict keys -> shapes the decoder needs (host-only; no GPU needed)."""
        lib = _lib.load()
        h = ctypes.c_void_p()
        cfg = _make_config(config)
        _lib.check(lib.cap4d_b200_vae_create(ctypes.byref(cfg), ctypes.byref(h)), "vae_create")
        try:
            n = ctypes.c_int()
            _lib.check(lib.cap4d_b200_vae_num_params(h, ctypes.byref(n)), "vae_num_params")
            out = {}
            name = ctypes.create_string_buffer(256)
            shape = (ctypes.c_int64 * 4)()
            nd = ctypes.c_int()
            for i in range(n.value):
                _lib.check(lib.cap4d_b200_vae_param_info(h, i, name, 256, shape, ctypes.byref(nd)), "vae_param_info")
                out[name.value.decode()] = tuple(int(shape[k]) for k in range(nd.value))
            return out
        finally:
            lib.cap4d_b200_vae_destroy(h)

    @staticmethod
    def encoder_param_shapes(config: Mapping) -> Dict[str, tuple]:
        """state_dict keys -> shapes of the optional encoder half (Encoder.__init__, model.py:466-516, + quant_conv)."""
        ch, mult, nrb = int(config["ch"]), list(config["ch_mult"]), int(config["num_res_blocks"])
        zc, e = int(config["z_channels"]), int(config.get("embed_dim", config["z_channels"]))
        out: Dict[str, tuple] = {"encoder.conv_in.weight": (ch, int(config["out_ch"]), 3, 3), "encoder.conv_in.bias": (ch,)}

        def res(p, cin, cout):
            out.update({p + "norm1.weight": (cin,), p + "norm1.bias": (cin,), p + "conv1.weight": (cout, cin, 3, 3),
                        p + "conv1.bias": (cout,), p + "norm2.weight": (cout,), p + "norm2.bias": (cout,),
                        p + "conv2.weight": (cout, cout, 3, 3), p + "conv2.bias": (cout,)})
            if cin != cout:
                out.update({p + "nin_shortcut.weight": (cout, cin, 1, 1), p + "nin_shortcut.bias": (cout,)})

        b = ch
        for lvl, m in enumerate(mult):
            for i in range(nrb):
                res(f"encoder.down.{lvl}.block.{i}.", b, ch * m)
                b = ch * m
            if lvl != len(mult) - 1:
                out[f"encoder.down.{lvl}.downsample.conv.weight"] = (b, b, 3, 3)
                out[f"encoder.down.{lvl}.downsample.conv.bias"] = (b,)
        res("encoder.mid.block_1.", b, b)
        out.update({"encoder.mid.attn_1.norm.weight": (b,), "encoder.mid.attn_1.norm.bias": (b,)})
        for n in ("q", "k", "v", "proj_out"):
            out[f"encoder.mid.attn_1.{n}.weight"] = (b, b, 1, 1)
            out[f"encoder.mid.attn_1.{n}.bias"] = (b,)
        res("encoder.mid.block_2.", b, b)
        out.update({"encoder.norm_out.weight": (b,), "encoder.norm_out.bias": (b,),
                    "encoder.conv_out.weight": (2 * zc, b, 3, 3), "encoder.conv_out.bias": (2 * zc,),
                    "quant_conv.weight": (2 * e, 2 * zc, 1, 1), "quant_conv.bias": (2 * e,)})
        return out

    @classmethod
    def random_init(cls, config: Mapping = VAE_CONFIG, seed: int = 0, device=None,
                    with_encoder: bool = False) -> "B200VAEDecoder":
        """Synthetic weights of the right architecture, generated on the GPU (benchmarks; no checkpoint offline)."""
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        g = torch.Generator(device=dev).manual_seed(seed)
        sd = {}
        shapes = dict(cls.param_shapes(config))
        if with_encoder:
            shapes.update(cls.encoder_param_shapes(config))
        for name, shape in shapes.items():
            if "norm" in name and name.endswith("weight"):
                t = 1.0 + 0.1 * torch.randn(shape, generator=g, device=dev)
            elif name.endswith("bias"):
                t = 0.05 * torch.randn(shape, generator=g, device=dev)
            else:
                fan_in = shape[1] * shape[2] * shape[3]
                t = torch.randn(shape, generator=g, device=dev) / fan_in ** 0.5
            sd[name] = t
        return cls(config, sd, device=dev)

    @classmethod
    def from_reference(cls, first_stage_model, scale_factor: float = SCALE_FACTOR, device=None) -> "B200VAEDecoder":
        return cls(config_from_reference(first_stage_model), first_stage_model.state_dict(), scale_factor=scale_factor,
                   device=device)

    def __deepcopy__(self, memo):
        raise RuntimeError("B200VAEDecoder is bound to one GPU: build one instance per device instead of deepcopy")

    @property
    def device(self):
        return self._device

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                self._lib.cap4d_b200_vae_destroy(self._handle)
        except Exception:
            pass

    def _workspace(self, N, H, W) -> torch.Tensor:
        key = (N, H, W)
        ws = self._ws.get(key)
        if ws is None:
            self._ws.clear()  # the library keeps one plan
            n = ctypes.c_size_t()
            _lib.check(self._lib.cap4d_b200_vae_workspace_bytes(self._handle, N, H, W, ctypes.byref(n)), "vae_workspace")
            ws = torch.empty(n.value + 2048, dtype=torch.uint8, device=self._device)
            self._ws[key] = ws
        return ws

    @torch.no_grad()
    def decode_to_uint8_bgr(self, z: torch.Tensor, batch: int = 8) -> torch.Tensor:
        """z [N, 4, h, w] -> uint8 [N, 8h, 8w, 3] on the host: the arrays `convert_and_save_latent_images`
        (cap4d/inference/utils.py:131-137) passes to cv2.imwrite, converted on the device (a quarter of the bytes
        of the fp32 images cross PCIe)."""
        if z.dim() != 4 or z.shape[1] != self.config["z_channels"]:
            raise ValueError("z must be [N, z_channels, h, w]")
        zs = z.to(device=self._device, dtype=torch.float32).contiguous()
        N, _, H, W = zs.shape
        out = torch.empty((N, 8 * H, 8 * W, 3), dtype=torch.uint8, device=self._device)
        with torch.cuda.device(self._device):
            stream = torch.cuda.current_stream(self._device).cuda_stream
            for i in range(0, N, batch):
                n = min(batch, N - i)
                ws = self._workspace(n, H, W)
                _lib.check(
                    self._lib.cap4d_b200_vae_decode_u8(self._handle, zs[i:i + n].data_ptr(), out[i:i + n].data_ptr(), n, H,
                                                       W, self.scale_factor, ws.data_ptr(), ws.numel(),
                                                       ctypes.c_void_p(stream)),
                    "vae_decode_u8")
        return out.cpu()

    @torch.no_grad()
    def decode_first_stage(self, z: torch.Tensor, batch: int = 4) -> torch.Tensor:
        """z: [N, 4, h, w] (or the reference's [N, V, 4, h, w]) sampler latents -> images like the reference's
        decode_first_stage; `batch` latents per launch plan."""
        lead = None
        if z.dim() == 5:
            lead = z.shape[:2]
            z = z.reshape(-1, *z.shape[2:])
        if z.dim() != 4 or z.shape[1] != self.config["z_channels"]:
            raise ValueError("z must be [N, z_channels, h, w]")
        zs = z.to(device=self._device, dtype=torch.float32).contiguous()
        N, _, H, W = zs.shape
        out = torch.empty((N, self.config["out_ch"], 8 * H, 8 * W), dtype=torch.float32, device=self._device)
        with torch.cuda.device(self._device):
            stream = torch.cuda.current_stream(self._device).cuda_stream
            i = 0
            while i < N:
                n = min(batch, N - i)
                ws = self._workspace(n, H, W)
                _lib.check(
                    self._lib.cap4d_b200_vae_decode(self._handle, zs[i:i + n].data_ptr(), out[i:i + n].data_ptr(), n, H, W,
                                                    self.scale_factor, ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream)),
                    "vae_decode")
                i += n
        out = out.to(z.device) if z.device != out.device else out
        return out.reshape(*lead, *out.shape[1:]) if lead is not None else out

    forward = decode_first_stage

    # ---- encoder half: AutoencoderKL.encode (controlnet/ldm/models/autoencoder.py:82-85) -------------------------
    @property
    def has_encoder(self) -> bool:
        """True when the state_dict carried the "encoder.*" / "quant_conv.*" entries."""
        yes = ctypes.c_int()
        _lib.check(self._lib.cap4d_b200_vae_has_encoder(self._handle, ctypes.byref(yes)), "vae_has_encoder")
        return bool(yes.value)

    @torch.no_grad()
    def encode_moments(self, x: torch.Tensor, batch: int = 2) -> torch.Tensor:
        """images [N, 3, H, W] in [-1, 1] -> posterior parameters [N, 2*embed_dim, H/f, W/f] (mean | logvar)."""
        if x.dim() != 4 or x.shape[1] != self.config["out_ch"]:
            raise ValueError("x must be [N, out_ch, H, W]")
        xs = x.to(device=self._device, dtype=torch.float32).contiguous()
        N, _, H, W = xs.shape
        f = 2 ** (len(self.config["ch_mult"]) - 1)
        out = torch.empty((N, 2 * self.config.get("embed_dim", self.config["z_channels"]), H // f, W // f),
                          dtype=torch.float32, device=self._device)
        with torch.cuda.device(self._device):
            stream = torch.cuda.current_stream(self._device).cuda_stream
            for i in range(0, N, batch):
                n = min(batch, N - i)
                key = ("enc", n, H, W)
                ws = self._ws.get(key)
                if ws is None:
                    nb = ctypes.c_size_t()
                    _lib.check(self._lib.cap4d_b200_vae_encode_workspace_bytes(self._handle, n, H, W, ctypes.byref(nb)),
                               "vae_encode_workspace_bytes")
                    self._ws = {k: v for k, v in self._ws.items() if k[0] != "enc"}  # the library keeps one plan
                    ws = self._ws[key] = torch.empty(nb.value + 2048, dtype=torch.uint8, device=self._device)
                _lib.check(
                    self._lib.cap4d_b200_vae_encode(self._handle, xs[i:i + n].data_ptr(), out[i:i + n].data_ptr(), n, H, W,
                                                    ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream)),
                    "vae_encode")
        return out

    def encode(self, x: torch.Tensor) -> "DiagonalGaussian":
        """AutoencoderKL.encode: returns the posterior (`.sample()`, `.mode()`, `.parameters`) like the reference."""
        return DiagonalGaussian(self.encode_moments(x))

    def encode_first_stage(self, x: torch.Tensor) -> torch.Tensor:
        """MMLDM.get_input's latent (cap4d/mmdm/mmdm.py:60-63 -> ddpm.py:611-619): scale_factor * posterior.sample().
        Accepts the reference's [B, T, 3, H, W] as well."""
        lead = None
        if x.dim() == 5:
            lead = x.shape[:2]
            x = x.reshape(-1, *x.shape[2:])
        z = self.scale_factor * self.encode(x).sample()
        return z.reshape(*lead, *z.shape[1:]) if lead is not None else z

    def num_launches(self) -> int:
        n = ctypes.c_int()
        _lib.check(self._lib.cap4d_b200_vae_num_launches(self._handle, ctypes.byref(n)), "vae_num_launches")
        return n.value


class DiagonalGaussian:
    """DiagonalGaussianDistribution (controlnet/ldm/modules/distributions/distributions.py:24-45) over device
    moments: same attributes and the same RNG use (`torch.randn(shape)` on the CPU generator, then moved)."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self) -> torch.Tensor:
        return self.mean + self.std * torch.randn(self.mean.shape).to(device=self.parameters.device)

    def mode(self) -> torch.Tensor:
        return self.mean


def config_from_reference(first_stage_model) -> Dict:
    """Decoder configuration of a reference AutoencoderKL (controlnet/ldm/models/autoencoder.py:14-45)."""
    dec = first_stage_model.decoder
    return dict(ch=dec.ch, ch_mult=tuple(_ch_mult(dec)), num_res_blocks=dec.num_res_blocks,
                z_channels=dec.conv_in.in_channels, embed_dim=first_stage_model.post_quant_conv.in_channels,
                out_ch=dec.conv_out.out_channels)


def install_vae(mmldm, device=None) -> B200VAEDecoder:
    """Route `mmldm.decode_first_stage` (cap4d/mmdm/mmdm.py:99-103 -> ddpm.py:822-830) through the B200 decoder:
    replaces `mmldm.first_stage_model.decode` in place.  ddpm.py has already divided by scale_factor there."""
    vae = B200VAEDecoder.from_reference(mmldm.first_stage_model, scale_factor=1.0, device=device)
    mmldm.first_stage_model.decode = vae.decode_first_stage
    if vae.has_encoder:
        # encode_first_stage (ddpm.py:832-834) -> first_stage_model.encode; get_first_stage_encoding (ddpm.py:656-663)
        # checks isinstance against the reference's own posterior class, so hand the moments to that class
        import sys

        post_cls = getattr(sys.modules.get(type(mmldm.first_stage_model).__module__), "DiagonalGaussianDistribution",
                           DiagonalGaussian)
        mmldm.first_stage_model.encode = lambda x: post_cls(vae.encode_moments(x))
    return vae


def _ch_mult(decoder):
    """ch_mult of a reference Decoder (it only stores ch and the built blocks): out channels of each level's
    last block / ch, lowest resolution last (model.py:583-600)."""
    return [decoder.up[lvl].block[-1].out_channels // decoder.ch for lvl in range(decoder.num_resolutions)]
