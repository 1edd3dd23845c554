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

    def __init__(self, config: Mapping, state_dict: Mapping[str, torch.Tensor], device: Optional[torch.device] = None,
                 keep_state: bool = False, lazy: bool = False, precision: str = "bf16"):
        """precision: "bf16" (default: 16-bit tensor-core operands, fp32 accumulation, <= 1e-2 of the fp32 reference)
        or "fp32" (the reference's own accuracy, <= 1e-4: exact bf16 x 3 operand splits on the same tensor-core
        kernels, exact pointwise maths; an order of magnitude slower).
        keep_state: keep references (no copies) to the fp32 `state_dict` tensors, which is what lets the module
        be deep-copied and moved to another GPU like the reference's nn.Module (generate_images.py:62-71:
        `copy.deepcopy(model).to(f"cuda:{i}")`).  lazy: upload / repack the weights at the first forward or
        `.to(device)` instead of now (implies keep_state)."""
        super().__init__()
        if not torch.cuda.is_available():
            raise RuntimeError("cap4d_b200: a CUDA device (B200, sm_100a) is required; there is no CPU path")
        self.config = dict(config)
        self._device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self._device.type != "cuda":
            raise RuntimeError("cap4d_b200: a CUDA device (B200, sm_100a) is required; there is no CPU path")
        if self._device.index is None:
            self._device = torch.device(f"cuda:{torch.cuda.current_device()}")
        self._lib = _lib.load()
        self._handle = ctypes.c_void_p()
        self._ws: Dict = {}  # (B, V, H, W, R) -> workspace tensor (a launch plan is bound to its workspace)
        self.dtype = torch.float32
        self.time_steps = int(config["time_steps"])
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.precision = precision
        _make_config(config)  # validate now, even when lazy
        self._state = dict(state_dict) if (keep_state or lazy) else None
        if not lazy:
            self._build(state_dict)

    def _build(self, state_dict: Mapping[str, torch.Tensor]) -> None:
        """Create the library handle on self._device: upload every tensor, repack (finalize)."""
        self._release()
        cfg = _make_config(self.config)
        handle = ctypes.c_void_p()
        with torch.cuda.device(self._device):
            _lib.check(self._lib.cap4d_b200_unet_create(ctypes.byref(cfg), ctypes.byref(handle)), "unet_create")
            try:
                _lib.check(self._lib.cap4d_b200_unet_set_precision(handle, int(self.precision == "fp32")), "set_precision")
                for name, t in state_dict.items():
                    t32 = t.detach().to(dtype=torch.float32).contiguous()
                    shape = (ctypes.c_int64 * max(1, t32.dim()))(*t32.shape)
                    _lib.check(
                        self._lib.cap4d_b200_unet_load_weight(handle, name.encode(), ctypes.c_void_p(t32.data_ptr()),
                                                              shape, t32.dim()),
                        f"load_weight({name})",
                    )
                _lib.check(self._lib.cap4d_b200_unet_finalize(handle), "unet_finalize")
            except Exception:
                self._lib.cap4d_b200_unet_destroy(handle)
                raise
        self._handle = handle

    def _release(self) -> None:
        if getattr(self, "_handle", None) is not None and self._handle.value:
            with torch.cuda.device(self._device):
                torch.cuda.synchronize(self._device)
                self._lib.cap4d_b200_unet_destroy(self._handle)
        self._handle = ctypes.c_void_p()
        self._ws = {}

    def _ensure_built(self) -> None:
        if not self._handle.value:
            if self._state is None:
                raise RuntimeError("B200MMDMUnet has no weights to build from")
            self._build(self._state)

    @staticmethod
    def param_shapes(config: Mapping) -> Dict[str, tuple]:
        """state_dict keys -> shapes of MMDMUnetModel for `config` (host-only; no GPU needed)."""
        lib = _lib.load()
        h = ctypes.c_void_p()
        cfg = _make_config(config)
        _lib.check(lib.cap4d_b200_unet_create(ctypes.byref(cfg), ctypes.byref(h)), "unet_create")
        try:
            n = ctypes.c_int()
            _lib.check(lib.cap4d_b200_unet_num_params(h, ctypes.byref(n)), "num_params")
            out = {}
            name = ctypes.create_string_buffer(256)
            shape = (ctypes.c_int64 * 4)()
            nd = ctypes.c_int()
            for i in range(n.value):
                _lib.check(lib.cap4d_b200_unet_param_info(h, i, name, 256, shape, ctypes.byref(nd)), "param_info")
                out[name.value.decode()] = tuple(int(shape[k]) for k in range(nd.value))
            return out
        finally:
            lib.cap4d_b200_unet_destroy(h)

    @classmethod
    def random_init(cls, config: Mapping, seed: int = 0, device=None, zero_std: float = 0.02) -> "B200MMDMUnet":
        """Synthetic weights of the right architecture, generated on the GPU (benchmarks; there is no
        network for checkpoints).  Layers the reference zero-initialises get N(0, zero_std) so the
        network output is not identically zero (SURVEY.md note Z)."""
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        g = torch.Generator(device=dev).manual_seed(seed)
        sd = {}
        for name, shape in cls.param_shapes(config).items():
            is_norm = name.startswith("out.0.") or any(
                k in name for k in (".in_layers.0.", ".out_layers.0.", ".norm.", ".norm1.", ".norm3."))
            is_zero = name.startswith(("out.2.", "cond_linear.")) or any(
                k in name for k in (".out_layers.3.", ".proj_out.", ".attn1.to_out.0."))
            if is_norm:
                t = (1.0 if name.endswith("weight") else 0.0) + 0.1 * torch.randn(shape, generator=g, device=dev)
            elif is_zero:
                t = zero_std * torch.randn(shape, generator=g, device=dev)
            elif name.endswith("bias"):
                t = 0.05 * torch.randn(shape, generator=g, device=dev)
            else:
                fan_in = 1
                for d in shape[1:]:
                    fan_in *= d
                t = (torch.rand(shape, generator=g, device=dev) * 2 - 1) / (fan_in ** 0.5)
            sd[name] = t
        return cls(config, sd, device=dev)

    @classmethod
    def from_reference(cls, ref_unet, device=None, lazy: bool = False, precision: str = "bf16") -> "B200MMDMUnet":
        """Built from a constructed reference MMDMUnetModel; keeps references to its fp32 tensors so that the
        result can be deep-copied / moved between GPUs like the module it replaces."""
        sd = {k: v.detach() for k, v in ref_unet.state_dict().items()}
        return cls(config_from_reference(ref_unet), sd, device=device, keep_state=True, lazy=lazy, precision=precision)

    # The reference copies and moves whole models (generate_images.py:62-71: copy.deepcopy(model).to("cuda:i")).
    # The device state of this module lives behind the library handle, so a copy shares the fp32 source tensors
    # (by reference) and builds its own handle where it lands: at `.to(device)` or at its first forward.
    def __deepcopy__(self, memo):
        if self._state is None:
            raise RuntimeError("this B200MMDMUnet was built without keep_state=True (e.g. random_init): it does not "
                               "hold its fp32 weights and cannot be copied; build one instance per device instead")
        new = B200MMDMUnet.__new__(B200MMDMUnet)
        torch.nn.Module.__init__(new)
        new.config = dict(self.config)
        new._device = self._device
        new._lib = self._lib
        new._handle = ctypes.c_void_p()
        new._ws = {}
        new.dtype = self.dtype
        new.time_steps = self.time_steps
        new.precision = self.precision
        new._state = self._state
        new.record_every = self.record_every
        memo[id(self)] = new
        return new

    def _apply(self, fn, recurse=True):
        """nn.Module.to / .cuda / .float land here.  Only the device matters (the arithmetic types are fixed);
        moving to another GPU rebuilds the handle there from the kept fp32 weights."""
        probe = fn(torch.empty(0, dtype=torch.float32, device=self._device))
        target = probe.device
        if target.type != "cuda":
            raise RuntimeError("cap4d_b200: there is no CPU path; B200MMDMUnet cannot be moved off the GPU")
        if target.index is None:
            target = torch.device(f"cuda:{torch.cuda.current_device()}")
        if target != self._device:
            if self._state is None:
                raise RuntimeError("this B200MMDMUnet was built without keep_state=True and cannot change device")
            self._release()
            self._device = target
        if self._state is not None:
            self._ensure_built()
        return self

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

    def check_ref_views(self) -> int:
        """Number of forwards since the last call whose `n_ref_views` promise was broken (a view declared a
        reference view had ref_mask != 1; its outputs were poisoned with NaN).  Synchronises the stream."""
        self._ensure_built()
        n = ctypes.c_int()
        with torch.cuda.device(self._device):
            torch.cuda.synchronize(self._device)  # forwards may run on non-blocking streams the library's read cannot see
            _lib.check(self._lib.cap4d_b200_unet_ref_view_violations(self._handle, ctypes.byref(n)), "ref_view_violations")
        return n.value

    def _workspace(self, B, V, H, W, R=0) -> torch.Tensor:
        self._ensure_built()
        _lib.check(self._lib.cap4d_b200_unet_set_ref_views(self._handle, int(R)), "set_ref_views")
        key = (B, V, H, W, R)
        ws = self._ws.get(key)
        if ws is None:
            if len(self._ws) >= 4:  # the library caches 4 plans as well
                self._ws.pop(next(iter(self._ws)))
            n = ctypes.c_size_t()
            _lib.check(self._lib.cap4d_b200_unet_workspace_bytes(self._handle, B, V, H, W, ctypes.byref(n)),
                       "workspace_bytes")
            ws = torch.empty(n.value + 2048, dtype=torch.uint8, device=self._device)
            self._ws[key] = ws
        return ws

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

    # every `record_every`-th forward records CUDA events around its launches (no synchronisation);
    # collect_timings() returns the per-class sums.  0 = off.
    record_every = 0
    _calls = 0

    @torch.no_grad()
    def forward(self, x, timesteps=None, context=None, control=None, **kwargs):
        """Reference signature (mmdm_unet.py:67).  Extra keyword `n_ref_views=R` (not in the reference; other
        unknown keywords such as only_mid_control are swallowed like there): the caller promises that the
        first R views of every group are reference views (ref_mask == 1), which lets the executor skip them
        once no later layer mixes views.  The returned tensor is the same either way."""
        assert context is None  # mmdm_unet.py:85
        xs, t, z, m, p = self._prep(x, timesteps, control)
        B, V, C, H, W = xs.shape
        R = int(kwargs.get("n_ref_views") or 0)
        if not 0 <= R < V:
            raise ValueError("n_ref_views must be in [0, V)")
        out = torch.empty((B, V, self.config["out_channels"], H, W), dtype=torch.float32, device=self._device)
        self._calls += 1
        with torch.cuda.device(self._device):
            ws = self._workspace(B, V, H, W, R)
            stream = torch.cuda.current_stream(self._device).cuda_stream
            if self.record_every and self._calls % self.record_every == 0:
                _lib.check(
                    self._lib.cap4d_b200_unet_forward_timed(self._handle, xs.data_ptr(), t.data_ptr(), z.data_ptr(),
                                                            m.data_ptr(), p.data_ptr(), out.data_ptr(), B, V, H, W,
                                                            ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream), None),
                    "unet_forward_timed",
                )
            else:
                _lib.check(
                    self._lib.cap4d_b200_unet_forward(self._handle, xs.data_ptr(), t.data_ptr(), z.data_ptr(),
                                                      m.data_ptr(), p.data_ptr(), out.data_ptr(), B, V, H, W,
                                                      ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream)),
                    "unet_forward",
                )
        return out.to(dtype=x.dtype) if x.dtype != torch.float32 else out

    @torch.no_grad()
    def forward_into(self, x, timesteps, z_input, ref_mask, pos_enc, out, n_ref_views: int = 0) -> None:
        """The same forward on caller-owned, contiguous fp32 device tensors (x, z_input, out [B, V, C, H, W];
        ref_mask [B, V, 1, H, W]; pos_enc [B, V, H, W, Cc]; timesteps int64 [B, V]): nothing is allocated or
        copied, so the call can be captured in a CUDA graph (B200StochasticIOSampler does)."""
        B, V, C, H, W = x.shape
        with torch.cuda.device(self._device):
            ws = self._workspace(B, V, H, W, int(n_ref_views))
            stream = torch.cuda.current_stream(self._device).cuda_stream
            self._calls += 1
            if self.record_every and self._calls % self.record_every == 0 and not torch.cuda.is_current_stream_capturing():
                _lib.check(
                    self._lib.cap4d_b200_unet_forward_timed(self._handle, x.data_ptr(), timesteps.data_ptr(),
                                                            z_input.data_ptr(), ref_mask.data_ptr(), pos_enc.data_ptr(),
                                                            out.data_ptr(), B, V, H, W, ws.data_ptr(), ws.numel(),
                                                            ctypes.c_void_p(stream), None),
                    "unet_forward_timed",
                )
            else:
                _lib.check(
                    self._lib.cap4d_b200_unet_forward(self._handle, x.data_ptr(), timesteps.data_ptr(),
                                                      z_input.data_ptr(), ref_mask.data_ptr(), pos_enc.data_ptr(),
                                                      out.data_ptr(), B, V, H, W, ws.data_ptr(), ws.numel(),
                                                      ctypes.c_void_p(stream)),
                    "unet_forward",
                )

    def enable_taps(self, on: bool = True) -> None:
        """Debug: keep the activation after every block of the next forwards (see taps())."""
        self._ensure_built()
        _lib.check(self._lib.cap4d_b200_unet_enable_taps(self._handle, int(bool(on))), "enable_taps")
        self._ws = {}  # the workspace size depends on it

    def taps(self, H: int, W: int) -> Dict[str, torch.Tensor]:
        """After a forward with taps enabled: {"input_blocks.3": [n_img, C, h, w] fp32, ...} (copies), in the
        reference's module naming and NCHW layout.  H, W: the latent size of that forward."""
        n = ctypes.c_int()
        _lib.check(self._lib.cap4d_b200_unet_num_taps(self._handle, ctypes.byref(n)), "num_taps")
        out = {}
        name = ctypes.create_string_buffer(128)
        ptr, rows, ch, nimg = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int(), ctypes.c_int()
        with torch.cuda.device(self._device):
            torch.cuda.synchronize(self._device)
            for i in range(n.value):
                _lib.check(self._lib.cap4d_b200_unet_tap_info(self._handle, i, name, 128, ctypes.byref(ptr),
                                                              ctypes.byref(rows), ctypes.byref(ch), ctypes.byref(nimg)),
                           "tap_info")
                ws = next(w for w in self._ws.values()
                          if w.data_ptr() <= ptr.value < w.data_ptr() + w.numel())
                off = ptr.value - ws.data_ptr()
                flat = ws[off: off + rows.value * ch.value * 4].view(torch.float32)
                hw = rows.value // nimg.value
                scale = int(round((H * W / hw) ** 0.5))
                t = flat.view(nimg.value, H // scale, W // scale, ch.value).permute(0, 3, 1, 2).contiguous()
                out[name.value.decode()] = t
        return out

    def collect_timings(self):
        """Per-class ms summed over the recorded forwards since the last call -> ({class: ms}, n_forwards)."""
        self._ensure_built()
        ms = (ctypes.c_float * _lib.N_CLASSES)()
        n = ctypes.c_int()
        _lib.check(self._lib.cap4d_b200_unet_collect_timings(self._handle, ms, ctypes.byref(n)), "collect_timings")
        return {c: float(ms[i]) for i, c in enumerate(_lib.CLASS_NAMES)}, n.value

    @torch.no_grad()
    def forward_timed(self, x, timesteps, control, n_ref_views=0):
        """One forward with CUDA events around every launch; returns (out, {class: ms})."""
        xs, t, z, m, p = self._prep(x, timesteps, control)
        B, V, C, H, W = xs.shape
        out = torch.empty((B, V, self.config["out_channels"], H, W), dtype=torch.float32, device=self._device)
        ms = (ctypes.c_float * _lib.N_CLASSES)()
        with torch.cuda.device(self._device):
            ws = self._workspace(B, V, H, W, int(n_ref_views))
            stream = torch.cuda.current_stream(self._device).cuda_stream
            _lib.check(
                self._lib.cap4d_b200_unet_forward_timed(self._handle, xs.data_ptr(), t.data_ptr(), z.data_ptr(),
                                                        m.data_ptr(), p.data_ptr(), out.data_ptr(), B, V, H, W,
                                                        ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream), ms),
                "unet_forward_timed",
            )
        return out, {n: float(ms[i]) for i, n in enumerate(_lib.CLASS_NAMES)}

    def select_plan(self, B: int, V: int, H: int, W: int, n_ref_views: int = 0) -> None:
        """Make the launch plan of this batch shape the current one (what class_stats / num_launches describe)."""
        with torch.cuda.device(self._device):
            ws = self._workspace(B, V, H, W, int(n_ref_views))
            _lib.check(self._lib.cap4d_b200_unet_plan(self._handle, B, V, H, W, ws.data_ptr(), ws.numel()), "unet_plan")

    def class_stats(self):
        """Algorithmic FLOPs / bytes / launches per kernel class of the current plan."""
        self._ensure_built()
        fl = (ctypes.c_double * _lib.N_CLASSES)()
        by = (ctypes.c_double * _lib.N_CLASSES)()
        ln = (ctypes.c_int * _lib.N_CLASSES)()
        _lib.check(self._lib.cap4d_b200_unet_class_stats(self._handle, fl, by, ln), "class_stats")
        ex = (ctypes.c_double * _lib.N_CLASSES)()
        _lib.check(self._lib.cap4d_b200_unet_class_exec_flops(self._handle, ex), "class_exec_flops")
        return {n: dict(flops=fl[i], exec_flops=ex[i], bytes=by[i], launches=ln[i])
                for i, n in enumerate(_lib.CLASS_NAMES)}

    def num_launches(self) -> int:
        self._ensure_built()
        n = ctypes.c_int()
        _lib.check(self._lib.cap4d_b200_unet_num_launches(self._handle, ctypes.byref(n)), "num_launches")
        return n.value


def install(mmldm, device=None, lazy: bool = False, precision: str = "bf16") -> B200MMDMUnet:
    """Replace `mmldm.model.diffusion_model` (ddpm.py:1318) by the B200 implementation, in place.

    lazy=True is the form for the reference's own driver (generate_images.py:59-71): call it right after
    `load_model(...)` while the model is still on the CPU; every `copy.deepcopy(model).to(f"cuda:{i}")` that
    follows then uploads and repacks the weights once, on its own GPU."""
    ref_unet = mmldm.model.diffusion_model
    new = B200MMDMUnet.from_reference(ref_unet, device=device, lazy=lazy, precision=precision)
    mmldm.model.diffusion_model = new
    return new
