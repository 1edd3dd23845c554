"""Stochastic-I/O DDIM sampler with a device-resident data plane.

`B200StochasticIOSampler.sample` keeps the signature and the random-number consumption of the
reference `StochasticIOSampler.sample` (cap4d/mmdm/sampler.py:64-233): x_T from the global torch
generator of the conditioning's device, per-step reference / generated-view permutations from the
global numpy generator in the reference's order, float64->float32 DDIM factors.  What changes is
where the data lives: conditioning and latents are uploaded once and stay in HBM, view groups are
batched `groups_per_call` at a time into one U-Net launch plan, the CFG combine and the DDIM update
are one fused kernel that scatters straight into the latent store, and with torch.distributed
initialised (one rank per GPU, NCCL) the groups of a step are dealt round-robin to the ranks like
the reference deals them to its device replicas (sampler.py:151-158), followed by one all-gather of
the freshly updated latents per step.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .schedule import MMDMSchedule, ddim_factors
from .unet import B200MMDMUnet


class B200MMLDM:
    """What the sampler needs from MMLDM (cap4d/mmdm/mmdm.py): schedule buffers, .device and
    apply_model(x_noisy, t, cond) (mmdm.py:113-124)."""

    def __init__(self, unet: B200MMDMUnet, schedule: Optional[MMDMSchedule] = None):
        self.unet = unet
        sch = schedule if schedule is not None else MMDMSchedule()
        self.num_timesteps = sch.num_timesteps
        self.betas = sch.betas
        self.alphas_cumprod = sch.alphas_cumprod
        self.alphas_cumprod_prev = sch.alphas_cumprod_prev
        self.only_mid_control = False

    @property
    def device(self):
        return self.unet.device

    def apply_model(self, x_noisy, t, cond, *args, **kwargs):
        assert isinstance(cond, dict)
        assert len(cond["c_concat"]) == 1
        return self.unet(x=x_noisy, timesteps=t, context=None, control=cond["c_concat"][0],
                         only_mid_control=self.only_mid_control)


def _find_unet(model) -> B200MMDMUnet:
    if isinstance(model, B200MMLDM):
        return model.unet
    inner = getattr(getattr(model, "model", None), "diffusion_model", None)
    if isinstance(inner, B200MMDMUnet):
        return inner
    raise TypeError("B200StochasticIOSampler needs a B200MMLDM or an MMLDM whose diffusion_model is a B200MMDMUnet "
                    "(see cap4d_b200.unet.install)")


class _SamplerState:
    """Device-resident state of one sample() call."""


class _CudaBackend:
    """The product data path: U-Net forward and fused CFG+DDIM update in libcap4d_b200.so."""

    def __init__(self, unet: B200MMDMUnet):
        self.unet = unet
        self.device = unet.device
        self._lib = _lib.load()

    def eps(self, x_in, t_in, control, n_ref_views=0):
        # the groups are built as cat([ref, gen], dim=1): the first R views are reference views, and only the
        # generated views' noise prediction is consumed (sampler.py:207-213)
        return self.unet(x_in, timesteps=t_in, context=None, control=control, n_ref_views=n_ref_views)

    def cfg_ddim_update(self, latents, eps, gen_idx, n, V, R, chw, cfg_scale, x_f, e_f):
        stream = torch.cuda.current_stream(self.device)
        _lib.check(
            self._lib.cap4d_b200_cfg_ddim_update(latents.data_ptr(), eps.data_ptr(), gen_idx.data_ptr(), n, V, R, chw,
                                                 float(cfg_scale), x_f, e_f, ctypes.c_void_p(stream.cuda_stream)),
            "cfg_ddim_update",
        )


class B200StochasticIOSampler:
    def __init__(self, model, groups_per_call: int = 1, backend=None, **kwargs):
        """`backend` exists for the host-logic tests (gloo, CPU); the default and only shipped backend is
        the CUDA library, which raises if the extension or a GPU is missing."""
        if isinstance(model, dict):  # the reference's {device_key: model} map: this process drives ONE GPU
            model = next(iter(model.values()))
        self.main_model = model
        self.backend = backend if backend is not None else _CudaBackend(_find_unet(model))
        self.ddpm_num_timesteps = model.num_timesteps
        self.groups_per_call = max(1, int(groups_per_call))
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.unet_calls = 0

    # -- distributed helpers -------------------------------------------------------------------
    @staticmethod
    def _dist():
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist, dist.get_rank(), dist.get_world_size()
        return None, 0, 1

    def _upload(self, d: Dict[str, torch.Tensor], dev) -> Dict[str, torch.Tensor]:
        out = {}
        for k, v in d.items():
            if v.device != dev:
                self.h2d_bytes += v.numel() * v.element_size()
            out[k] = v.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
        return out

    @torch.no_grad()
    def sample(self, S: int, ref_cond: Dict[str, torch.Tensor], ref_uncond: Dict[str, torch.Tensor],
               gen_cond: Dict[str, torch.Tensor], gen_uncond: Dict[str, torch.Tensor],
               latent_shape: Tuple[int, int, int], V: int = 8, R_max: int = 4, cfg_scale: float = 1.0,
               eta: float = 0.0, verbose: bool = False) -> torch.Tensor:
        """Same contract as StochasticIOSampler.sample (sampler.py:64-233)."""
        st = self.begin(S, ref_cond, ref_uncond, gen_cond, gen_uncond, latent_shape, V=V, R_max=R_max,
                        cfg_scale=cfg_scale, eta=eta)
        while st.i < st.n_steps:
            self.step(st)
        return self.end(st)

    # The three phases of sample(), exposed so that a caller (bench.py) can time the steady state:
    # begin = schedule + x_T + one-time upload, step = one DDIM step over all views, end = download.
    @torch.no_grad()
    def begin(self, S, ref_cond, ref_uncond, gen_cond, gen_uncond, latent_shape, V=8, R_max=4, cfg_scale=1.0,
              eta=0.0) -> "_SamplerState":
        st = _SamplerState()
        dev = self.backend.device
        st.mem_device = next(iter(gen_cond.values())).device
        st.n_gen = next(iter(gen_cond.values())).shape[0]
        st.n_all_ref = next(iter(ref_cond.values())).shape[0]
        st.V, st.R = V, min(st.n_all_ref, R_max)
        st.G = V - st.R
        assert st.n_gen % st.G == 0, \
            f"number of generated images ({st.n_gen}) has to be divisible by G ({st.G})"  # sampler.py:108
        st.n_its = st.n_gen // st.G
        st.steps, st.x_factors, st.e_factors = ddim_factors(self.main_model.alphas_cumprod, S, eta)
        st.n_steps = len(st.steps)
        st.cfg_scale = float(cfg_scale)
        # same generator, same call as the reference (sampler.py:112)
        all_x = torch.randn((st.n_gen, *latent_shape), device=st.mem_device)
        if all_x.device != dev:
            self.h2d_bytes += all_x.numel() * 4
        st.latents = all_x.to(dev).contiguous()
        st.rc, st.ru = self._upload(ref_cond, dev), self._upload(ref_uncond, dev)
        st.gc, st.gu = self._upload(gen_cond, dev), self._upload(gen_uncond, dev)
        st.chw = int(np.prod(latent_shape))
        st.i = 0
        return st

    @torch.no_grad()
    def step(self, st: "_SamplerState") -> None:
        dev = self.backend.device
        dist, rank, world = self._dist()
        n_its, R, V = st.n_its, st.R, st.V
        step = st.steps[st.i]
        # permutations: identical numpy consumption to sampler.py:131-139 (on every rank)
        if R == 1:
            ref_batches = np.zeros((n_its, R), dtype=np.int64)
        else:
            ref_batches = np.stack([np.random.permutation(np.arange(st.n_all_ref))[:R] for _ in range(n_its)], axis=0)
        gen_batches = np.reshape(np.random.permutation(np.arange(st.n_gen)), (n_its, -1))
        my_groups = np.arange(rank, n_its, world)  # round-robin like sampler.py:151-158
        x_f, e_f = float(st.x_factors[st.i]), float(st.e_factors[st.i])
        rc, ru, gc, gu, latents = st.rc, st.ru, st.gc, st.gu, st.latents

        for c0 in range(0, len(my_groups), self.groups_per_call):
            grp = my_groups[c0:c0 + self.groups_per_call]
            n = len(grp)
            ref_idx = torch.from_numpy(ref_batches[grp]).to(dev, non_blocking=True)                       # [n, R]
            gen_idx = torch.from_numpy(np.ascontiguousarray(gen_batches[grp])).to(dev, non_blocking=True)  # [n, G]
            control = {}
            for key in rc:
                cond = torch.cat([rc[key][ref_idx], gc[key][gen_idx]], dim=1)
                unc = torch.cat([ru[key][ref_idx], gu[key][gen_idx]], dim=1)
                control[key] = torch.cat([unc, cond], dim=0)                                              # [2n, V, ...]
            x_in = torch.cat([rc["z_input"][ref_idx], latents[gen_idx]], dim=1)
            x_in = torch.cat([x_in, x_in], dim=0)
            t_in = torch.full((2 * n, V), int(step), device=dev, dtype=torch.long)
            eps = self.backend.eps(x_in, t_in, control, n_ref_views=R)
            self.unet_calls += 1
            self.backend.cfg_ddim_update(latents, eps, gen_idx, n, V, R, st.chw, st.cfg_scale, x_f, e_f)

        if world > 1:
            self._exchange(dist, rank, world, latents, gen_batches, n_its, st.G, st.chw)
        st.i += 1

    @torch.no_grad()
    def end(self, st: "_SamplerState") -> torch.Tensor:
        out = st.latents.to(st.mem_device)
        if out.device != st.latents.device:
            self.d2h_bytes += out.numel() * 4
        return out

    def _exchange(self, dist, rank, world, latents, gen_batches, n_its, G, chw):
        """One all-gather per DDIM step: every rank contributes the views it just updated."""
        dev = latents.device
        per_rank = (n_its + world - 1) // world
        send = torch.zeros((per_rank * G, chw), device=dev, dtype=torch.float32)
        mine = np.arange(rank, n_its, world)
        if len(mine):
            idx = torch.from_numpy(np.ascontiguousarray(gen_batches[mine]).reshape(-1)).to(dev)
            send[: idx.numel()] = latents.view(-1, chw)[idx]
        recv = torch.empty((world, per_rank * G, chw), device=dev, dtype=torch.float32)
        dist.all_gather_into_tensor(recv.view(-1, chw), send)
        flat = latents.view(-1, chw)
        for r in range(world):
            if r == rank:
                continue
            theirs = np.arange(r, n_its, world)
            if len(theirs) == 0:
                continue
            idx = torch.from_numpy(np.ascontiguousarray(gen_batches[theirs]).reshape(-1)).to(dev)
            flat[idx] = recv[r, : idx.numel()]
