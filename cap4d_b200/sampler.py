"""Stochastic-I/O DDIM sampler with a device-resident data plane.

`B200StochasticIOSampler.sample` keeps the signature and the random-number consumption of the
reference `StochasticIOSampler.sample` (cap4d/mmdm/sampler.py:64-233): x_T from the global torch
generator of the conditioning's device, per-step reference / generated-view permutations from the
global numpy generator in the reference's order, float64->float32 DDIM factors.  What changes is
where the data lives.  Conditioning and latents are uploaded once and stay in HBM; the all-zero
unconditional conditioning (cap4dcond.py:78-88) is neither stored nor uploaded; per step the host
draws the permutations and sends the two index tables and the per-call parameters in ONE small
copy; view groups are batched `groups_per_call` at a time, and one call = gather kernel (builds the
U-Net batch from the stores through the index tables) -> U-Net -> fused CFG + DDIM update that
scatters straight into the latent store.  That sequence is captured once per batch shape in a CUDA
graph and replayed for every call of every step (everything that varies is read from device
memory).  With torch.distributed initialised (one rank per GPU, NCCL) the groups of a step are
dealt round-robin to the ranks like the reference deals them to its device replicas
(sampler.py:151-158), followed by pack kernel -> one all-gather -> unpack kernel per step.
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
    """State of one sample() call."""


_KEYS = ("z_input", "ref_mask", "pos_enc")


class _CudaBackend:
    """The product data path: everything between the host's permutation draw and the updated latent store runs
    in libcap4d_b200.so on device-resident data."""

    def __init__(self, unet: B200MMDMUnet, use_cuda_graph: bool = True):
        self.unet = unet
        self.device = unet.device
        self._lib = _lib.load()
        self.use_cuda_graph = bool(use_cuda_graph)
        self.graphs_captured = 0
        self.graph_replays = 0
        self.h2d_bytes = 0

    # ---- one-time upload -----------------------------------------------------------------------
    def _store(self, t: torch.Tensor, rows: int) -> torch.Tensor:
        if t.device != self.device:
            self.h2d_bytes += t.numel() * 4
        return t.to(device=self.device, dtype=torch.float32, non_blocking=True).reshape(rows, -1).contiguous()

    @staticmethod
    def _all_zero(t: torch.Tensor) -> bool:
        if t.device.type == "cpu" and t.is_contiguous() and t.dtype == torch.float32 and t.numel() % 2 == 0:
            if not t.numpy().reshape(-1).view(np.uint64).any():  # one pass over the bytes: all +0.0
                return True
        return not bool(torch.count_nonzero(t))  # -0.0 (the reference's `z_input * 0.`) is zero too

    def begin(self, st, ref_cond, ref_uncond, gen_cond, gen_uncond, all_x) -> None:
        dev = self.device
        for d in (ref_cond, ref_uncond, gen_cond, gen_uncond):
            if set(d.keys()) != set(_KEYS):
                raise ValueError(f"conditioning dicts must hold exactly {_KEYS}")
        zshape = gen_cond["z_input"].shape      # [n, 1, C, H, W]
        pshape = gen_cond["pos_enc"].shape      # [n, 1, H, W, Cc]
        st.C, st.H, st.W, st.Cc = int(zshape[-3]), int(zshape[-2]), int(zshape[-1]), int(pshape[-1])
        if all_x.device != dev:
            self.h2d_bytes += all_x.numel() * 4
        st.latents = all_x.to(dev).reshape(st.n_gen, -1).contiguous()
        keep = {}
        stores = _lib.SamplerStores()
        for prefix, cond, unc, rows in (("ref", ref_cond, ref_uncond, st.n_all_ref), ("gen", gen_cond, gen_uncond, st.n_gen)):
            for key, field in (("z_input", "z"), ("ref_mask", "mask"), ("pos_enc", "pos")):
                t = keep[f"{prefix}_{field}"] = self._store(cond[key], rows)
                setattr(stores, f"{prefix}_{field}", t.data_ptr())
                u = unc[key]
                # the unconditional branch of CAP4DConditioning is zeros for z_input / pos_enc and the same
                # ref_mask (cap4dcond.py:78-88): then there is nothing to keep or upload
                redundant = torch.equal(u, cond[key]) if key == "ref_mask" else self._all_zero(u)
                if redundant:
                    setattr(stores, f"{prefix}_{field}_u", None)
                else:
                    tu = keep[f"{prefix}_{field}_u"] = self._store(u, rows)
                    setattr(stores, f"{prefix}_{field}_u", tu.data_ptr())
        stores.latents = st.latents.data_ptr()
        st.stores, st.keep = stores, keep
        # per-step host -> device block: [ref_idx int64 n_its*R | gen_idx int64 n_its*G | calls], double-buffered
        n_calls = (len(range(st.rank, st.n_its, st.world)) + st.gpc - 1) // st.gpc
        st.n_calls = n_calls
        call_bytes = ctypes.sizeof(_lib.SamplerCall)
        st.off_gen = st.n_its * st.R * 8
        st.off_calls = st.off_gen + st.n_its * st.G * 8
        st.block_bytes = st.off_calls + max(1, n_calls) * call_bytes
        st.host_blocks = [torch.empty(st.block_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
        st.host_events = [None, None]
        st.dev_block = torch.empty(st.block_bytes, dtype=torch.uint8, device=dev)
        st.dev_call = torch.empty(call_bytes, dtype=torch.uint8, device=dev)  # the slot the kernels read
        st.call_bytes = call_bytes
        st.bufs = {}    # n -> (x_in, z_in, mask, pos, t_in, eps)
        st.graphs = {}  # n -> torch.cuda.CUDAGraph
        st.xchg = None
        if self.use_cuda_graph:
            # Capture the call graph of every batch shape this rank will use (full calls and a ragged tail) NOW, so
            # that step() is replays from its first call on.  A capture needs one eager run first (it builds the
            # launch plan and sets the kernels' attributes): gather + U-Net on zeroed index tables (row 0 of every
            # store) - the update kernel, the only one that writes the latent store, is not part of the warm-up.
            st.dev_block.zero_()
            st.dev_call.zero_()
            mine = len(range(st.rank, st.n_its, st.world))
            shapes = sorted({min(st.gpc, mine)} | ({mine % st.gpc} if mine > st.gpc else set()), reverse=True)
            for n in shapes:
                if n > 0:
                    self._capture(st, n)

    def _capture(self, st, n: int) -> None:
        with torch.cuda.device(self.device):
            self._launch_call(st, n, update=False)
            torch.cuda.current_stream(self.device).synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._launch_call(st, n)
            st.graphs[n] = g
            self.graphs_captured += 1

    # ---- per step ------------------------------------------------------------------------------
    def start_step(self, st, ref_batches: np.ndarray, gen_batches: np.ndarray, step: int, x_f: float, e_f: float,
                   my_groups: np.ndarray) -> None:
        slot = st.i & 1
        if st.host_events[slot] is not None:
            st.host_events[slot].synchronize()  # the copy issued two steps ago has read this pinned block
        hb = st.host_blocks[slot].numpy()
        if st.R > 0:
            hb[: st.off_gen].view(np.int64)[:] = ref_batches.reshape(-1)
        hb[st.off_gen: st.off_calls].view(np.int64)[:] = gen_batches.reshape(-1)
        calls = (_lib.SamplerCall * max(1, st.n_calls)).from_buffer(hb, st.off_calls)
        for ci, c0 in enumerate(range(0, len(my_groups), st.gpc)):
            grp = my_groups[c0:c0 + st.gpc]
            calls[ci].timestep = int(step)
            calls[ci].x_coef = x_f
            calls[ci].e_coef = e_f
            calls[ci].n_groups = len(grp)
            for k, gidx in enumerate(grp):
                calls[ci].groups[k] = int(gidx)
        del calls
        st.dev_block.copy_(st.host_blocks[slot], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        st.host_events[slot] = ev
        self.h2d_bytes += st.block_bytes

    def _buffers(self, st, n: int):
        b = st.bufs.get(n)
        if b is None:
            dev, V = self.device, st.V
            f32 = dict(dtype=torch.float32, device=dev)
            b = (torch.empty((2 * n, V, st.C, st.H, st.W), **f32), torch.empty((2 * n, V, st.C, st.H, st.W), **f32),
                 torch.empty((2 * n, V, 1, st.H, st.W), **f32), torch.empty((2 * n, V, st.H, st.W, st.Cc), **f32),
                 torch.empty((2 * n, V), dtype=torch.int64, device=dev),
                 torch.empty((2 * n, V, st.C, st.H, st.W), **f32))
            st.bufs[n] = b
        return b

    def _launch_call(self, st, n: int, update: bool = True) -> None:
        x_in, z_in, m_in, p_in, t_in, eps = self._buffers(st, n)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        ref_ptr = st.dev_block.data_ptr() if st.R > 0 else None
        gen_ptr = st.dev_block.data_ptr() + st.off_gen
        _lib.check(
            self._lib.cap4d_b200_sampler_gather(ctypes.byref(st.stores), ref_ptr, gen_ptr, st.dev_call.data_ptr(), n,
                                                st.V, st.R, st.C, st.H, st.W, st.Cc, x_in.data_ptr(), z_in.data_ptr(),
                                                m_in.data_ptr(), p_in.data_ptr(), t_in.data_ptr(), stream),
            "sampler_gather")
        self.unet.forward_into(x_in, t_in, z_in, m_in, p_in, eps, n_ref_views=st.R)
        if not update:
            return
        _lib.check(
            self._lib.cap4d_b200_sampler_update(st.latents.data_ptr(), eps.data_ptr(), gen_ptr, st.dev_call.data_ptr(), n,
                                                st.V, st.R, st.chw, st.cfg_scale, stream),
            "sampler_update")

    def run_call(self, st, call_index: int, n: int) -> None:
        """Groups `calls[call_index]` of the current step: gather -> U-Net -> CFG + DDIM scatter."""
        with torch.cuda.device(self.device):
            # the kernels (and the captured graph) read their parameters from one fixed device slot
            off = st.off_calls + call_index * st.call_bytes
            st.dev_call.copy_(st.dev_block[off: off + st.call_bytes], non_blocking=True)
            # every record_every-th call is launched from the host with CUDA events around its kernels
            # (B200MMDMUnet.record_every, bench.py's per-class times); a graph replay has no host side to time
            rec = self.unet.record_every
            if not self.use_cuda_graph or (rec and (self.unet._calls + 1) % rec == 0):
                self._launch_call(st, n)
                return
            g = st.graphs.get(n)
            if g is None:  # a shape begin() did not foresee: run it from the host
                self._launch_call(st, n)
                return
            g.replay()
            self.unet._calls += 1
            self.graph_replays += 1

    # ---- multi-GPU -----------------------------------------------------------------------------
    def exchange(self, st, dist) -> None:
        """One all-gather per DDIM step: every rank contributes the views it just updated."""
        per_rank = (st.n_its + st.world - 1) // st.world
        if st.xchg is None:
            st.xchg = (torch.zeros((per_rank * st.G, st.chw), device=self.device, dtype=torch.float32),
                       torch.empty((st.world, per_rank * st.G, st.chw), device=self.device, dtype=torch.float32))
        send, recv = st.xchg
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        gen_ptr = st.dev_block.data_ptr() + st.off_gen
        _lib.check(self._lib.cap4d_b200_sampler_pack(st.latents.data_ptr(), gen_ptr, st.n_its, st.G, st.chw, st.rank,
                                                     st.world, send.data_ptr(), stream), "sampler_pack")
        dist.all_gather_into_tensor(recv.view(-1, st.chw), send)
        _lib.check(self._lib.cap4d_b200_sampler_unpack(st.latents.data_ptr(), recv.data_ptr(), gen_ptr, st.n_its, st.G,
                                                       st.chw, st.rank, st.world, stream), "sampler_unpack")

    def end(self, st) -> torch.Tensor:
        return st.latents.view(st.n_gen, *st.latent_shape)


class B200StochasticIOSampler:
    def __init__(self, model, groups_per_call: int = 1, backend=None, use_cuda_graph: bool = True, **kwargs):
        """`backend` exists for the host-logic tests (gloo, CPU): an object with `.device`, `eps(x_in, t_in,
        control, n_ref_views)` and `cfg_ddim_update(...)` drives the same host logic through eager torch
        indexing.  The default and only shipped backend is the CUDA library, which raises if the extension or a
        GPU is missing."""
        if isinstance(model, dict):  # the reference's {device_key: model} map: this process drives ONE GPU
            model = next(iter(model.values()))
        self.main_model = model
        self.backend = backend if backend is not None else _CudaBackend(_find_unet(model), use_cuda_graph)
        self.ddpm_num_timesteps = model.num_timesteps
        self.groups_per_call = max(1, min(int(groups_per_call), _lib.MAX_GROUPS_PER_CALL))
        self._h2d_eager = 0
        self.d2h_bytes = 0
        self.unet_calls = 0

    @property
    def h2d_bytes(self) -> int:
        return self._h2d_eager + getattr(self.backend, "h2d_bytes", 0)

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
                self._h2d_eager += v.numel() * v.element_size()
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
        _, st.rank, st.world = self._dist()
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
        st.latent_shape = tuple(latent_shape)
        st.chw = int(np.prod(latent_shape))
        st.gpc = self.groups_per_call
        st.i = 0
        # same generator, same call as the reference (sampler.py:112)
        all_x = torch.randn((st.n_gen, *latent_shape), device=st.mem_device)
        st.device_plane = hasattr(self.backend, "run_call")
        if st.device_plane:
            self.backend.begin(st, ref_cond, ref_uncond, gen_cond, gen_uncond, all_x)
        else:
            if all_x.device != dev:
                self._h2d_eager += all_x.numel() * 4
            st.latents = all_x.to(dev).contiguous()
            st.rc, st.ru = self._upload(ref_cond, dev), self._upload(ref_uncond, dev)
            st.gc, st.gu = self._upload(gen_cond, dev), self._upload(gen_uncond, dev)
        return st

    @torch.no_grad()
    def step(self, st: "_SamplerState") -> None:
        dist, rank, world = self._dist()
        n_its, R = st.n_its, st.R
        step = st.steps[st.i]
        # permutations: identical numpy consumption to sampler.py:131-139 (on every rank)
        if R == 1:
            ref_batches = np.zeros((n_its, R), dtype=np.int64)
        else:
            ref_batches = np.stack([np.random.permutation(np.arange(st.n_all_ref))[:R] for _ in range(n_its)], axis=0)
        gen_batches = np.reshape(np.random.permutation(np.arange(st.n_gen)), (n_its, -1))
        my_groups = np.arange(rank, n_its, world)  # round-robin like sampler.py:151-158
        x_f, e_f = float(st.x_factors[st.i]), float(st.e_factors[st.i])
        if st.device_plane:
            self.backend.start_step(st, ref_batches, gen_batches, int(step), x_f, e_f, my_groups)
            for ci, c0 in enumerate(range(0, len(my_groups), st.gpc)):
                self.backend.run_call(st, ci, len(my_groups[c0:c0 + st.gpc]))
                self.unet_calls += 1
            if world > 1:
                self.backend.exchange(st, dist)
        else:
            self._step_eager(st, ref_batches, gen_batches, my_groups, int(step), x_f, e_f)
            if world > 1:
                self._exchange_eager(dist, rank, world, st.latents, gen_batches, n_its, st.G, st.chw)
        st.i += 1

    @torch.no_grad()
    def end(self, st: "_SamplerState") -> torch.Tensor:
        lat = self.backend.end(st) if st.device_plane else st.latents
        out = lat.to(st.mem_device)
        if out.device != lat.device:
            self.d2h_bytes += out.numel() * 4
        return out

    # ---- host-logic test route (backend doubles): the same grouping through eager torch indexing -------------
    def _step_eager(self, st, ref_batches, gen_batches, my_groups, step, x_f, e_f) -> None:
        dev = self.backend.device
        R, V = st.R, st.V
        rc, ru, gc, gu, latents = st.rc, st.ru, st.gc, st.gu, st.latents
        for c0 in range(0, len(my_groups), st.gpc):
            grp = my_groups[c0:c0 + st.gpc]
            n = len(grp)
            ref_idx = torch.from_numpy(ref_batches[grp]).to(dev)                                    # [n, R]
            gen_idx = torch.from_numpy(np.ascontiguousarray(gen_batches[grp])).to(dev)              # [n, G]
            control = {}
            for key in rc:
                cond = torch.cat([rc[key][ref_idx], gc[key][gen_idx]], dim=1)
                unc = torch.cat([ru[key][ref_idx], gu[key][gen_idx]], dim=1)
                control[key] = torch.cat([unc, cond], dim=0)                                        # [2n, V, ...]
            x_in = torch.cat([rc["z_input"][ref_idx], latents[gen_idx]], dim=1)
            x_in = torch.cat([x_in, x_in], dim=0)
            t_in = torch.full((2 * n, V), step, device=dev, dtype=torch.long)
            eps = self.backend.eps(x_in, t_in, control, n_ref_views=R)
            self.unet_calls += 1
            self.backend.cfg_ddim_update(latents, eps, gen_idx, n, V, R, st.chw, st.cfg_scale, x_f, e_f)

    def _exchange_eager(self, dist, rank, world, latents, gen_batches, n_its, G, chw):
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
