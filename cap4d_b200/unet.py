"""Host-side mirror of the reference U-Net interface, backed by the sm_100a C-ABI library.

`B200MMDMUnet` has the call signature of `MMDMUnetModel.forward`
(reference cap4d/mmdm/net/mmdm_unet.py:67-126) so it can be assigned to
`mmldm.model.diffusion_model` (controlnet/ldm/models/diffusion/ddpm.py:1318) and be driven by the
unmodified `MMLDM.apply_model` (cap4d/mmdm/mmdm.py:113-124) and `StochasticIOSampler`.
PyTorch is used for device memory and streams only; all arithmetic runs in libcap4d_b200.so.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Mapping, Optional

import torch

from . import _lib

_CFG_KEYS = ("in_channels", "out_channels", "model_channels", "condition_channels", "num_res_blocks",
             "channel_mult", "attention_resolutions", "num_head_channels", "time_steps")


def _make_config(cfg: Mapping) -> _lib.UnetConfig:
    missing = [k for k in _CFG_KEYS if k not in cfg]
    if missing:
        raise ValueError(f"unet config is missing {missing}")
    c = _lib.UnetConfig()
    c.in_channels = int(cfg["in_channels"])
    c.out_channels = int(cfg["out_channels"])
    c.model_channels = int(cfg["model_channels"])
    c.condition_channels = int(cfg["condition_channels"])
    c.num_res_blocks = int(cfg["num_res_blocks"])
    mult = list(cfg["channel_mult"])
    attn = list(cfg["attention_resolutions"])
    if len(mult) > _lib.MAX_LEVELS or len(attn) > _lib.MAX_LEVELS:
        raise ValueError("too many levels")
    c.n_levels = len(mult)
    for i, m in enumerate(mult):
        c.channel_mult[i] = int(m)
    c.n_attention_resolutions = len(attn)
    for i, a in enumerate(attn):
        c.attention_resolutions[i] = int(a)
    c.num_head_channels = int(cfg["num_head_channels"])
    c.time_steps = int(cfg["time_steps"])
    return c


def config_from_reference(ref_unet) -> Dict:
    """Read the hyper-parameters back from a constructed reference MMDMUnetModel."""
    return dict(
        in_channels=ref_unet.in_channels,
        out_channels=ref_unet.out_channels,
        model_channels=ref_unet.model_channels,
        condition_channels=ref_unet.cond_linear.in_features,
        num_res_blocks=ref_unet.num_res_blocks[0] if isinstance(ref_unet.num_res_blocks, (list, tuple)) else ref_unet.num_res_blocks,
        channel_mult=tuple(ref_unet.channel_mult),
        attention_resolutions=tuple(ref_unet.attention_resolutions),
        num_head_channels=ref_unet.num_head_channels,
        time_steps=ref_unet.time_steps,
    )


class B200MMDMUnet(torch.nn.Module):
    """Drop-in for MMDMUnetModel on one B200.  Weights are uploaded and repacked once."""

    def __init__(self, config: Mapping, state_dict: Mapping[str, torch.Tensor], device: Optional[torch.device] = None):
        super().__init__()
        if not torch.cuda.is_available():
            raise RuntimeError("cap4d_b200: a CUDA device (B200, sm_100a) is required; there is no CPU path")
        self.config = dict(config)
        self._device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._lib = _lib.load()
        self._handle = ctypes.c_void_p()
        self._ws: Optional[torch.Tensor] = None
        self._ws_key = None
        self.dtype = torch.float32
        self.time_steps = int(config["time_steps"])
        cfg = _make_config(config)
        with torch.cuda.device(self._device):
            _lib.check(self._lib.cap4d_b200_unet_create(ctypes.byref(cfg), ctypes.byref(self._handle)), "unet_create")
            for name, t in state_dict.items():
                t32 = t.detach().to(dtype=torch.float32).contiguous()
                shape = (ctypes.c_int64 * max(1, t32.dim()))(*t32.shape)
                _lib.check(
                    self._lib.cap4d_b200_unet_load_weight(self._handle, name.encode(), ctypes.c_void_p(t32.data_ptr()),
                                                          shape, t32.dim()),
                    f"load_weight({name})",
                )
            _lib.check(self._lib.cap4d_b200_unet_finalize(self._handle), "unet_finalize")

    @classmethod
    def from_reference(cls, ref_unet, device=None) -> "B200MMDMUnet":
        return cls(config_from_reference(ref_unet), ref_unet.state_dict(), device=device)

    # the reference moves/copies whole models around (generate_images.py:62-71); this module is bound
    # to the device it was built on
    def __deepcopy__(self, memo):
        raise RuntimeError("B200MMDMUnet is bound to one GPU: build one instance per device instead of deepcopy")

    @property
    def device(self):
        return self._device

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None and self._handle.value:
                self._lib.cap4d_b200_unet_destroy(self._handle)
                self._handle = ctypes.c_void_p()
        except Exception:
            pass

    def _workspace(self, B, V, H, W) -> torch.Tensor:
        key = (B, V, H, W)
        if self._ws is None or self._ws_key != key:
            n = ctypes.c_size_t()
            _lib.check(self._lib.cap4d_b200_unet_workspace_bytes(self._handle, B, V, H, W, ctypes.byref(n)),
                       "workspace_bytes")
            self._ws = None
            self._ws = torch.empty(n.value + 2048, dtype=torch.uint8, device=self._device)
            self._ws_key = key
        return self._ws

    def _prep(self, x, timesteps, control):
        if x.dim() != 5:
            raise ValueError("x must be [B, V, C, H, W]")
        dev = self._device

        def f32(t):
            return t.to(device=dev, dtype=torch.float32).contiguous()

        xs = f32(x)
        z = f32(control["z_input"])
        m = f32(control["ref_mask"])
        p = f32(control["pos_enc"])
        t = timesteps.to(device=dev, dtype=torch.int64).contiguous()
        B, V, C, H, W = xs.shape
        if z.shape != xs.shape or m.shape != (B, V, 1, H, W) or p.shape[:4] != (B, V, H, W) or t.shape != (B, V):
            raise ValueError("control tensors do not match x")
        if p.shape[4] != self.config["condition_channels"] or C != self.config["in_channels"]:
            raise ValueError("channel count mismatch")
        return xs, t, z, m, p

    @torch.no_grad()
    def forward(self, x, timesteps=None, context=None, control=None, **kwargs):
        assert context is None  # mmdm_unet.py:85
        xs, t, z, m, p = self._prep(x, timesteps, control)
        B, V, C, H, W = xs.shape
        out = torch.empty((B, V, self.config["out_channels"], H, W), dtype=torch.float32, device=self._device)
        with torch.cuda.device(self._device):
            ws = self._workspace(B, V, H, W)
            stream = torch.cuda.current_stream(self._device).cuda_stream
            _lib.check(
                self._lib.cap4d_b200_unet_forward(self._handle, xs.data_ptr(), t.data_ptr(), z.data_ptr(), m.data_ptr(),
                                                  p.data_ptr(), out.data_ptr(), B, V, H, W, ws.data_ptr(), ws.numel(),
                                                  ctypes.c_void_p(stream)),
                "unet_forward",
            )
        return out.to(dtype=x.dtype) if x.dtype != torch.float32 else out

    @torch.no_grad()
    def forward_timed(self, x, timesteps, control):
        """One forward with CUDA events around every launch; returns (out, {class: ms})."""
        xs, t, z, m, p = self._prep(x, timesteps, control)
        B, V, C, H, W = xs.shape
        out = torch.empty((B, V, self.config["out_channels"], H, W), dtype=torch.float32, device=self._device)
        ms = (ctypes.c_float * _lib.N_CLASSES)()
        with torch.cuda.device(self._device):
            ws = self._workspace(B, V, H, W)
            stream = torch.cuda.current_stream(self._device).cuda_stream
            _lib.check(
                self._lib.cap4d_b200_unet_forward_timed(self._handle, xs.data_ptr(), t.data_ptr(), z.data_ptr(),
                                                        m.data_ptr(), p.data_ptr(), out.data_ptr(), B, V, H, W,
                                                        ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream), ms),
                "unet_forward_timed",
            )
        return out, {n: float(ms[i]) for i, n in enumerate(_lib.CLASS_NAMES)}

    def class_stats(self):
        """Algorithmic FLOPs / bytes / launches per kernel class of the current plan."""
        fl = (ctypes.c_double * _lib.N_CLASSES)()
        by = (ctypes.c_double * _lib.N_CLASSES)()
        ln = (ctypes.c_int * _lib.N_CLASSES)()
        _lib.check(self._lib.cap4d_b200_unet_class_stats(self._handle, fl, by, ln), "class_stats")
        return {n: dict(flops=fl[i], bytes=by[i], launches=ln[i]) for i, n in enumerate(_lib.CLASS_NAMES)}

    def num_launches(self) -> int:
        n = ctypes.c_int()
        _lib.check(self._lib.cap4d_b200_unet_num_launches(self._handle, ctypes.byref(n)), "num_launches")
        return n.value


def install(mmldm, device=None) -> B200MMDMUnet:
    """Replace `mmldm.model.diffusion_model` (ddpm.py:1318) by the B200 implementation, in place."""
    ref_unet = mmldm.model.diffusion_model
    new = B200MMDMUnet.from_reference(ref_unet, device=device)
    mmldm.model.diffusion_model = new
    return new
