"""The sampler's device-resident data plane (cap4d_b200/csrc/sampler_plane.cu) against the reference's own
host-side data movement (cap4d/mmdm/sampler.py:141-213), restated with eager torch indexing: bit-exact, because it
is data movement plus the separately rounded CFG / DDIM arithmetic."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import mmdm_oracle as O

pytestmark = pytest.mark.gpu


def _stores(lib_mod, dev, rc, ru, gc, gu, latents, drop_uncond):
    keep, st = {}, lib_mod.SamplerStores()
    for prefix, cond, unc in (("ref", rc, ru), ("gen", gc, gu)):
        for key, field in (("z_input", "z"), ("ref_mask", "mask"), ("pos_enc", "pos")):
            n = cond[key].shape[0]
            t = keep[(prefix, field)] = cond[key].to(dev).reshape(n, -1).contiguous()
            setattr(st, f"{prefix}_{field}", t.data_ptr())
            if drop_uncond:
                setattr(st, f"{prefix}_{field}_u", None)
            else:
                u = keep[(prefix, field, "u")] = unc[key].to(dev).reshape(n, -1).contiguous()
                setattr(st, f"{prefix}_{field}_u", u.data_ptr())
    st.latents = latents.data_ptr()
    return st, keep


def _call(lib_mod, dev, timestep, x_f, e_f, groups):
    c = lib_mod.SamplerCall()
    c.timestep, c.x_coef, c.e_coef, c.n_groups = timestep, x_f, e_f, len(groups)
    for i, g in enumerate(groups):
        c.groups[i] = int(g)
    host = torch.frombuffer(bytearray(bytes(c)), dtype=torch.uint8)
    return host.to(dev)


@pytest.mark.parametrize("n_ref,R,V,zero_uncond", [(1, 1, 4, True), (5, 3, 6, True), (4, 2, 4, False), (3, 3, 8, False)])
def test_gather_and_update_match_reference_data_movement(cuda_device, n_ref, R, V, zero_uncond):
    from cap4d_b200 import _lib

    lib = _lib.load()
    dev = cuda_device
    G = V - R
    n_its, C, H, W, Cc = 5, 4, 8, 8, 6
    n_gen = n_its * G
    cfg = dict(in_channels=C, condition_channels=Cc)
    rc, ru, gc, gu = O.make_sampler_conditioning(cfg, n_ref, n_gen, H, W, seed=3)
    g = torch.Generator().manual_seed(7)
    if not zero_uncond:  # exercise the stored unconditional branch with data the reference would never produce
        for d in (ru, gu):
            d["z_input"] = torch.randn(d["z_input"].shape, generator=g)
            d["pos_enc"] = torch.randn(d["pos_enc"].shape, generator=g)
            d["ref_mask"] = (torch.rand(d["ref_mask"].shape, generator=g) > 0.5).float()
    latents = torch.randn(n_gen, C * H * W, generator=g).to(dev)
    lat0 = latents.clone()
    rng = np.random.RandomState(1)
    ref_b = np.stack([rng.permutation(n_ref)[:R] for _ in range(n_its)]).astype(np.int64)
    gen_b = rng.permutation(n_gen).reshape(n_its, G).astype(np.int64)
    ref_d, gen_d = torch.from_numpy(ref_b).to(dev), torch.from_numpy(gen_b).to(dev)
    groups = [3, 0, 4]
    n = len(groups)
    stores, keep = _stores(_lib, dev, rc, ru, gc, gu, latents, drop_uncond=zero_uncond)
    call = _call(_lib, dev, 417, 0.9375, -0.3125, groups)
    f32 = dict(dtype=torch.float32, device=dev)
    x_in, z_in = torch.full((2 * n, V, C, H, W), 9.0, **f32), torch.full((2 * n, V, C, H, W), 9.0, **f32)
    m_in, p_in = torch.full((2 * n, V, 1, H, W), 9.0, **f32), torch.full((2 * n, V, H, W, Cc), 9.0, **f32)
    t_in = torch.zeros((2 * n, V), dtype=torch.int64, device=dev)
    _lib.check(lib.cap4d_b200_sampler_gather(ctypes.byref(stores), ref_d.data_ptr(), gen_d.data_ptr(), call.data_ptr(), n,
                                             V, R, C, H, W, Cc, x_in.data_ptr(), z_in.data_ptr(), m_in.data_ptr(),
                                             p_in.data_ptr(), t_in.data_ptr(), None), "gather")
    # the reference's construction (sampler.py:171-195) on the host
    ri, gi = torch.from_numpy(ref_b[groups]), torch.from_numpy(gen_b[groups])
    want = {}
    for key in rc:
        cond = torch.cat([rc[key][ri], gc[key][gi]], dim=1)
        unc = torch.cat([ru[key][ri], gu[key][gi]], dim=1)
        want[key] = torch.cat([unc, cond], dim=0)
    x_t = lat0.cpu().view(n_gen, 1, C, H, W)[gi].squeeze(2)
    xw = torch.cat([rc["z_input"][ri].squeeze(2) if rc["z_input"].dim() == 5 else rc["z_input"][ri], x_t], dim=1)
    xw = torch.cat([xw, xw], dim=0)
    assert torch.equal(x_in.cpu(), xw.reshape(x_in.shape))
    assert torch.equal(z_in.cpu(), want["z_input"].reshape(z_in.shape))
    assert torch.equal(m_in.cpu(), want["ref_mask"].reshape(m_in.shape))
    assert torch.equal(p_in.cpu(), want["pos_enc"].reshape(p_in.shape))
    assert torch.equal(t_in.cpu(), torch.full((2 * n, V), 417, dtype=torch.int64))
    # update: sampler.py:205-231 with separately rounded ops
    eps = torch.randn(2 * n, V, C * H * W, generator=g).to(dev)
    _lib.check(lib.cap4d_b200_sampler_update(latents.data_ptr(), eps.data_ptr(), gen_d.data_ptr(), call.data_ptr(), n, V,
                                             R, C * H * W, 1.75, None), "update")
    e = eps.cpu()
    mo = (e[:n] + 1.75 * (e[n:] - e[:n]))[:, R:]
    exp = lat0.cpu().clone()
    idx = gi.reshape(-1)
    exp[idx] = exp[idx] * torch.tensor(0.9375) + mo.reshape(-1, C * H * W) * torch.tensor(-0.3125)
    assert torch.equal(latents.cpu(), exp)


@pytest.mark.parametrize("world,n_its", [(2, 6), (3, 4), (4, 2), (3, 7)])
def test_pack_unpack_emulated_ranks(cuda_device, world, n_its):
    """Every rank's pack -> (all-gather, emulated) -> unpack on ONE GPU: each rank must end up with the union of
    all ranks' updates, whatever the (ragged) shares."""
    from cap4d_b200 import _lib

    lib = _lib.load()
    dev = cuda_device
    G, chw = 3, 64
    n_gen = n_its * G
    rng = np.random.RandomState(world * 10 + n_its)
    gen_b = torch.from_numpy(rng.permutation(n_gen).reshape(n_its, G).astype(np.int64)).to(dev)
    base = torch.randn(n_gen, chw, device=dev)
    per_rank = (n_its + world - 1) // world
    stores, final = [], base.clone()
    for r in range(world):  # rank r updated the views of its groups
        lat = base.clone()
        for gidx in range(r, n_its, world):
            rows = gen_b[gidx]
            lat[rows] = lat[rows] + 100.0 * (r + 1)
            final[rows] = lat[rows]
        stores.append(lat)
    recv = torch.zeros(world, per_rank * G, chw, device=dev)
    for r in range(world):
        send = torch.zeros(per_rank * G, chw, device=dev)
        _lib.check(lib.cap4d_b200_sampler_pack(stores[r].data_ptr(), gen_b.data_ptr(), n_its, G, chw, r, world,
                                               send.data_ptr(), None), "pack")
        recv[r] = send
    for r in range(world):
        _lib.check(lib.cap4d_b200_sampler_unpack(stores[r].data_ptr(), recv.data_ptr(), gen_b.data_ptr(), n_its, G, chw,
                                                 r, world, None), "unpack")
        assert torch.equal(stores[r], final), f"rank {r}"


def test_sampler_graph_replay_is_bit_identical_and_uploads_no_zeros(cuda_device):
    from cap4d_b200 import B200MMDMUnet, B200MMLDM, B200StochasticIOSampler

    cfg = O.TINY_CONFIG
    model = B200MMLDM(B200MMDMUnet(cfg, O.init_state_dict(cfg, seed=0), device=cuda_device))
    rc, ru, gc, gu = O.make_sampler_conditioning(cfg, 1, 15, 8, 8, seed=11)
    kw = dict(S=4, ref_cond=rc, ref_uncond=ru, gen_cond=gc, gen_uncond=gu, latent_shape=(4, 8, 8), V=4, R_max=4,
              cfg_scale=2.0)
    outs = {}
    for graphs in (True, False):
        torch.manual_seed(3)
        np.random.seed(3)
        s = B200StochasticIOSampler(model, groups_per_call=2, use_cuda_graph=graphs)
        outs[graphs] = s.sample(**kw)
        if graphs:
            # 5 groups in calls of 2 + 2 + 1: two batch shapes, both captured in begin(); every call is a replay
            assert s.backend.graphs_captured == 2 and s.backend.graph_replays == 3 * 4
            cond_bytes = sum(t.numel() * 4 for d in (rc, gc) for t in d.values())
            x_bytes = 15 * 4 * 8 * 8 * 4
            step_bytes = 4 * ((5 * 1 + 5 * 3) * 8 + 3 * ctypes.sizeof(__import__("cap4d_b200")._lib.SamplerCall))
            assert s.h2d_bytes == cond_bytes + x_bytes + step_bytes  # nothing of the all-zero unconditional dicts
    assert torch.equal(outs[True], outs[False])
