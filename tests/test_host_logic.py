"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header declares,
the product's own schedule code reproduces the reference fixtures bit for bit, and the sampler's
host logic (RNG consumption, grouping, multi-rank partition + exchange) reproduces the reference
sampler when the device kernels are replaced by a torch stand-in defined HERE (tests only)."""
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

from oracle import mmdm_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


# ---------------------------------------------------------------------------------------------
# C ABI
# ---------------------------------------------------------------------------------------------
def _header_symbols():
    text = open(os.path.join(ROOT, "include", "cap4d_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cap4d_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from cap4d_b200 import _lib

    syms = _header_symbols()
    assert len(syms) >= 15
    lib = _lib.load()  # builds nothing: the .so must already exist (see __graft_entry__.build)
    for s in syms:
        assert hasattr(lib, s), f"{s} is declared in include/cap4d_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in cap4d_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(syms)
    assert b"sm_100a" in lib.cap4d_b200_version()


def test_topology_and_argument_errors_without_gpu():
    """create() only derives the block topology (host code), so it works without a device; bad
    arguments come back as status codes + text, never as exceptions across the ABI."""
    import ctypes

    from cap4d_b200 import _lib
    from cap4d_b200.unet import _make_config

    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.cap4d_b200_unet_create(ctypes.byref(_make_config(O.PRODUCTION_CONFIG)), ctypes.byref(h)) == 0
    n = ctypes.c_size_t()
    assert lib.cap4d_b200_unet_workspace_bytes(h, 2, 8, 64, 64, ctypes.byref(n)) == 0
    assert 0.5e9 < n.value < 8e9  # skip tensors + scratch of one production-shape forward
    assert lib.cap4d_b200_unet_workspace_bytes(h, 2, 8, 60, 60, ctypes.byref(n)) != 0  # not divisible by 8
    assert "divisible" in _lib.last_error()
    assert lib.cap4d_b200_unet_destroy(h) == 0
    bad = dict(O.PRODUCTION_CONFIG, num_head_channels=32)
    assert lib.cap4d_b200_unet_create(ctypes.byref(_make_config(bad)), ctypes.byref(h)) != 0
    assert "64" in _lib.last_error()
    assert lib.cap4d_b200_unet_create(None, ctypes.byref(h)) != 0


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cap4d_b200 import B200MMDMUnet

    with pytest.raises(RuntimeError, match="no CPU path"):
        B200MMDMUnet(O.TINY_CONFIG, {})


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cap4d_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("TEST INFRASTRUCTURE", ""), f"{f} mentions the oracle"


# ---------------------------------------------------------------------------------------------
# schedule (product code, cap4d_b200/schedule.py) against the reference fixtures
# ---------------------------------------------------------------------------------------------
def test_schedule_bit_exact():
    from cap4d_b200 import MMDMSchedule, ddim_factors

    g = np.load(os.path.join(GOLD, "schedule.npz"))
    s = MMDMSchedule()
    assert s.num_timesteps == 1000
    assert np.array_equal(s.betas.numpy(), g["betas"])
    assert np.array_equal(s.alphas_cumprod.numpy(), g["alphas_cumprod"])
    assert np.array_equal(s.alphas_cumprod_prev.numpy(), g["alphas_cumprod_prev"])
    for S in (10, 100):
        steps, xf, ef = ddim_factors(s.alphas_cumprod, S)
        assert np.array_equal(steps, np.flip(g[f"ddim_timesteps_{S}"]))
        assert np.array_equal(xf, np.flip(g[f"x_factor_{S}"]))
        assert np.array_equal(ef, np.flip(g[f"e_factor_{S}"]))
    with pytest.raises(IndexError):
        ddim_factors(s.alphas_cumprod, 3)


def test_ddim_factors_with_eta_match_oracle():
    # eta only changes the e_t coefficient (there is no sigma * noise term in sampler.py:215-231)
    from cap4d_b200 import MMDMSchedule, ddim_factors

    s = MMDMSchedule()
    for S, eta in ((10, 0.5), (20, 1.0)):
        steps, xf, ef = ddim_factors(s.alphas_cumprod, S, eta)
        ts, a, ap, sg = O.ddim_schedule(s.alphas_cumprod.numpy(), S, eta=eta)
        assert np.array_equal(steps, np.flip(ts))
        for i in range(S):
            x_ref, e_ref = O.ddim_coefficients(a, ap, sg, S - 1 - i)
            assert float(x_ref) == float(xf[i]) and float(e_ref) == float(ef[i])


# ---------------------------------------------------------------------------------------------
# sampler host logic with a torch stand-in for the device kernels
# ---------------------------------------------------------------------------------------------
class _TorchBackend:
    """Test double for cap4d_b200.sampler._CudaBackend: oracle U-Net + eager CFG/DDIM arithmetic."""

    def __init__(self, sd, cfg):
        self.sd, self.cfg = sd, cfg
        self.device = torch.device("cpu")

    def eps(self, x_in, t_in, control, n_ref_views=0):
        assert float(control["ref_mask"][:, :n_ref_views].min()) == 1.0  # the promise the product path relies on
        return O.unet_forward(self.sd, self.cfg, x_in, t_in, control)

    def cfg_ddim_update(self, latents, eps, gen_idx, n, V, R, chw, cfg_scale, x_f, e_f):
        eu, ec = eps[:n], eps[n:]
        e = (eu + cfg_scale * (ec - eu))[:, R:].reshape(-1, *latents.shape[1:])
        idx = gen_idx.reshape(-1)
        latents[idx] = latents[idx] * torch.tensor(x_f, dtype=torch.float32) + e * torch.tensor(e_f, dtype=torch.float32)


class _Model:
    def __init__(self):
        from cap4d_b200 import MMDMSchedule

        s = MMDMSchedule()
        self.num_timesteps, self.alphas_cumprod = s.num_timesteps, s.alphas_cumprod
        self.betas, self.alphas_cumprod_prev = s.betas, s.alphas_cumprod_prev


def _run_sampler(name, gpc):
    from cap4d_b200 import B200StochasticIOSampler

    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = O.TINY_CONFIG
    sd = O.init_state_dict(cfg, seed=int(g["wseed"]))
    H, W, V = int(g["H"]), int(g["W"]), int(g["V"])
    rc, ru, gc, gu = O.make_sampler_conditioning(cfg, int(g["n_ref"]), int(g["n_gen"]), H, W, seed=int(g["cseed"]))
    torch.manual_seed(int(g["seed"]))
    np.random.seed(int(g["seed"]))
    sampler = B200StochasticIOSampler(_Model(), groups_per_call=gpc, backend=_TorchBackend(sd, cfg))
    z = sampler.sample(S=int(g["S"]), ref_cond=rc, ref_uncond=ru, gen_cond=gc, gen_uncond=gu,
                       latent_shape=(cfg["in_channels"], H, W), V=V, R_max=int(g["R_max"]),
                       cfg_scale=float(g["cfg_scale"]))
    return z, torch.from_numpy(g["out"])


@pytest.mark.parametrize("name,gpc", [("sampler_r1", 1), ("sampler_r1", 2), ("sampler_r2", 1), ("sampler_r2", 4)])
def test_sampler_host_logic_matches_reference(name, gpc):
    z, ref = _run_sampler(name, gpc)
    assert z.shape == ref.shape
    assert O.max_rel_err(z, ref) < 1e-4  # same RNG stream, same grouping, fp32 summation-order noise only
    assert O.psnr(z, ref) > 80.0


def test_sampler_rejects_indivisible_view_count():
    from cap4d_b200 import B200StochasticIOSampler

    cfg = O.TINY_CONFIG
    rc, ru, gc, gu = O.make_sampler_conditioning(cfg, 1, 5, 8, 8, seed=0)
    s = B200StochasticIOSampler(_Model(), backend=_TorchBackend({}, cfg))
    with pytest.raises(AssertionError, match="divisible"):  # sampler.py:108
        s.sample(S=4, ref_cond=rc, ref_uncond=ru, gen_cond=gc, gen_uncond=gu, latent_shape=(4, 8, 8), V=4)


# ---------------------------------------------------------------------------------------------
# ragged group shapes against the oracle sampler, with a cheap cross-view stand-in for the U-Net (a view's
# output depends on every view of its group, its conditioning and the timestep, so a wrong grouping, a
# wrong reference draw or a wrong scatter changes the result)
# ---------------------------------------------------------------------------------------------
def _toy_eps(x, t, control):
    ctx = (x + control["z_input"]).mean(dim=1, keepdim=True)                       # couples the views of a group
    pe = control["pos_enc"].mean(dim=-1)[:, :, None]                               # [B, V, 1, h, w]
    tt = (t.to(torch.float32) / 1000.0)[:, :, None, None, None]
    return torch.tanh(0.3 * x + 0.2 * ctx + 0.1 * pe + 0.05 * control["ref_mask"] + tt)


class _ToyBackend(_TorchBackend):
    def __init__(self):
        self.device = torch.device("cpu")
        self.calls = []

    def eps(self, x_in, t_in, control, n_ref_views=0):
        assert float(control["ref_mask"][:, :n_ref_views].min()) == 1.0
        assert float(control["ref_mask"][:, n_ref_views:].max()) == 0.0
        self.calls.append(x_in.shape[0] // 2)
        return _toy_eps(x_in, t_in, control)


RAGGED = [  # n_ref, n_gen, V, R_max, groups_per_call, eta
    (1, 6, 4, 4, 5, 0.0),    # 2 groups, groups_per_call larger than the group count
    (5, 4, 4, 2, 3, 0.0),    # more references than R_max: 2 of 5 drawn per group and step
    (3, 3, 4, 4, 2, 0.0),    # R = 3: a single generated view per group, 3 groups in calls of 2 + 1
    (2, 8, 6, 4, 3, 0.5),    # eta > 0 (only the eps coefficient changes, sampler.py:215-231)
    (10, 12, 8, 4, 2, 0.0),  # multi_ref.yaml's shape: 4 of 10 references, 3 groups
]


def _ragged_inputs(n_ref, n_gen, seed=5):
    cfg = dict(in_channels=4, condition_channels=6)
    return O.make_sampler_conditioning(cfg, n_ref, n_gen, 4, 4, seed=seed)


def _run_ragged(case, backend):
    from cap4d_b200 import B200StochasticIOSampler

    n_ref, n_gen, V, R_max, gpc, eta = case
    rc, ru, gc, gu = _ragged_inputs(n_ref, n_gen)
    torch.manual_seed(9)
    np.random.seed(9)
    sampler = B200StochasticIOSampler(_Model(), groups_per_call=gpc, backend=backend)
    return sampler.sample(S=5, ref_cond=rc, ref_uncond=ru, gen_cond=gc, gen_uncond=gu, latent_shape=(4, 4, 4), V=V,
                          R_max=R_max, cfg_scale=1.7, eta=eta)


@pytest.mark.parametrize("case", RAGGED)
def test_sampler_ragged_groups_match_oracle_sampler(case):
    n_ref, n_gen, V, R_max, gpc, eta = case
    rc, ru, gc, gu = _ragged_inputs(n_ref, n_gen)
    torch.manual_seed(9)
    np.random.seed(9)
    want = O.stochastic_io_sample(_toy_eps, _Model().alphas_cumprod.numpy(), 5, rc, ru, gc, gu, (4, 4, 4), V=V,
                                  R_max=R_max, cfg_scale=1.7, eta=eta)
    be = _ToyBackend()
    z = _run_ragged(case, be)
    assert z.shape == want.shape == (n_gen, 4, 4, 4)
    assert O.max_rel_err(z, want) < 1e-5
    n_its = n_gen // (V - min(n_ref, R_max))
    per_step = [min(gpc, n_its - c0) for c0 in range(0, n_its, gpc)]
    assert be.calls == per_step * 5  # every group exactly once per step, batched groups_per_call at a time


def _ragged_rank_main(rank, world, port, case, out_dir):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        be = _ToyBackend()
        z = _run_ragged(case, be)
        torch.save((z, sum(be.calls)), os.path.join(out_dir, f"z{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,case", [(3, (1, 12, 4, 4, 1, 0.0)),    # 4 groups on 3 ranks: shares of 2, 1, 1
                                        (3, (10, 8, 8, 4, 2, 0.0))])   # 2 groups on 3 ranks: one rank idles
def test_sampler_ragged_rank_shares_gloo(tmp_path, world, case):
    """Ranks with unequal (or no) shares of a step's groups: the padded all-gather must still leave every rank
    with the single-process latents."""
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_ragged_rank_main, args=(world, port, case, str(tmp_path)), nprocs=world, join=True)
    want = _run_ragged(case, _ToyBackend())
    n_its = case[1] // (case[2] - min(case[0], case[3]))
    groups = 0
    for r in range(world):
        z, g = torch.load(tmp_path / f"z{r}.pt")
        assert torch.equal(z, want), f"rank {r}"   # same arithmetic per group: bit-identical to one process
        groups += g
    assert groups == n_its * 5


# ---------------------------------------------------------------------------------------------
# N > 1: two gloo ranks on CPU must reproduce the single-process result
# ---------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, name, out_dir):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z, _ = _run_sampler(name, 1)
        torch.save(z, os.path.join(out_dir, f"z{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["sampler_r1", "sampler_r2"])
def test_sampler_two_ranks_gloo(tmp_path, name):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_rank_main, args=(2, port, name, str(tmp_path)), nprocs=2, join=True)
    z0 = torch.load(tmp_path / "z0.pt")
    z1 = torch.load(tmp_path / "z1.pt")
    _, ref = _run_sampler(name, 1)
    assert torch.equal(z0, z1)  # every rank holds the full, identical latent store after each exchange
    assert O.max_rel_err(z0, ref) < 1e-4


def test_vae_config_from_reference_module():
    """The decoder configuration is read off a reference AutoencoderKL (only where the reference is present)."""
    from oracle import ref_import as RI
    from oracle import vae_oracle as VO

    if not RI.reference_available():
        pytest.skip("reference sources not present (GPU box)")
    from cap4d_b200.vae import B200VAEDecoder, config_from_reference

    vae = RI.build_reference_vae(VO.TINY_VAE)
    cfg = config_from_reference(vae)
    assert cfg == dict(VO.TINY_VAE, ch_mult=tuple(VO.TINY_VAE["ch_mult"]))
    want = {k: tuple(v.shape) for k, v in vae.state_dict().items() if k.startswith(("decoder.", "post_quant_conv."))}
    assert B200VAEDecoder.param_shapes(cfg) == want


def test_vae_encoder_param_shapes_match_oracle():
    from cap4d_b200 import B200VAEDecoder
    from oracle import vae_oracle as VO

    for cfg in (VO.TINY_VAE, VO.PRODUCTION_VAE):
        assert B200VAEDecoder.encoder_param_shapes(cfg) == dict(VO.vae_encoder_param_shapes(cfg))


def test_config_from_reference_matches_the_real_module():
    """config_from_reference / param_shapes against a constructed reference MMDMUnetModel (the unmodified module,
    from /root/reference or its verbatim copy under oracle/_ref): same hyper-parameters back, same state_dict keys
    and shapes - what install() relies on."""
    from oracle import mmdm_oracle as O
    from oracle import ref_import as RI

    if not RI.reference_available():
        pytest.skip("reference modules not available (run oracle/build_ref.py)")
    from cap4d_b200 import B200MMDMUnet
    from cap4d_b200.unet import config_from_reference

    ref = RI.build_reference_unet(O.TINY_CONFIG)
    cfg = config_from_reference(ref)
    assert cfg == {k: (tuple(v) if isinstance(v, (list, tuple)) else v) for k, v in O.TINY_CONFIG.items()}
    want = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    got = B200MMDMUnet.param_shapes(cfg)
    assert got == want and len(got) > 100
