"""Drop-in proof: the UNMODIFIED reference caller on top of the B200 U-Net.

The reference's own `StochasticIOSampler.sample` (cap4d/mmdm/sampler.py:63-233) and `MMLDM.apply_model`
(cap4d/mmdm/mmdm.py:113-124) are imported from oracle/_ref - the verbatim copy of the reference modules made by
oracle/build_ref.py (see there; /root/reference does not exist on the GPU box) - and run over
`cap4d_b200.unet.install(mmldm)`; the result is held to the fixtures the reference produced on its own U-Net
(tests/golden/sampler_r*.npz, oracle/make_golden.py).  Also: the model-moving idiom of the reference's driver,
`copy.deepcopy(model).to("cuda:i")` (cap4d/inference/generate_images.py:62-71)."""
import copy
import os

import numpy as np
import pytest
import torch

from oracle import mmdm_oracle as O
from oracle import ref_import as RI

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not RI.reference_available(), reason="oracle/_ref missing: run oracle/build_ref.py")]
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _reference_mmldm(wseed):
    model = RI.build_reference_mmldm(O.TINY_CONFIG)
    model.model.diffusion_model.load_state_dict(O.init_state_dict(O.TINY_CONFIG, seed=wseed))
    return model


@pytest.mark.parametrize("name", ["sampler_r1", "sampler_r2"])
def test_reference_sampler_over_installed_unet(cuda_device, name):
    from cap4d_b200 import B200MMDMUnet
    from cap4d_b200.unet import install

    _, Sampler, _ = RI.import_reference()
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = O.TINY_CONFIG
    model = _reference_mmldm(int(g["wseed"]))
    new = install(model, device=cuda_device)
    assert isinstance(model.model.diffusion_model, B200MMDMUnet) and new is model.model.diffusion_model
    # the reference's sampler reads `.device` of every model in its map (sampler.py:166); Lightning tracks it
    # through .to(); the stub of oracle/ref_import.py does the same
    model = model.to(cuda_device)
    assert torch.device(model.device) == cuda_device
    H, W, V = int(g["H"]), int(g["W"]), int(g["V"])
    rc, ru, gc, gu = O.make_sampler_conditioning(cfg, int(g["n_ref"]), int(g["n_gen"]), H, W, seed=int(g["cseed"]))
    torch.manual_seed(int(g["seed"]))
    np.random.seed(int(g["seed"]))
    z = Sampler({str(cuda_device): model}).sample(
        S=int(g["S"]), ref_cond=rc, ref_uncond=ru, gen_cond=gc, gen_uncond=gu,
        latent_shape=(cfg["in_channels"], H, W), V=V, R_max=int(g["R_max"]), cfg_scale=float(g["cfg_scale"]))
    ref = torch.from_numpy(g["out"])
    assert z.shape == ref.shape and z.device.type == "cpu"
    p = O.psnr(z, ref)
    print(f"reference sampler over install(): {name} PSNR {p:.1f} dB max-rel {O.max_rel_err(z, ref):.3e}")
    assert p >= 40.0


def test_reference_apply_model_over_installed_unet(cuda_device):
    """MMLDM.apply_model -> DiffusionWrapper -> diffusion_model(x=..., timesteps=..., context=..., control=...,
    only_mid_control=...) (mmdm.py:113-124, ddpm.py:1330-1345) must reach the B200 forward unchanged."""
    from cap4d_b200.unet import install

    g = np.load(os.path.join(GOLD, "unet_tiny_v4_h16.npz"))
    model = _reference_mmldm(int(g["wseed"]))
    install(model, device=cuda_device)
    x, t, ctrl = O.make_inputs(O.TINY_CONFIG, B=int(g["B"]), V=int(g["V"]), H=int(g["H"]), W=int(g["W"]),
                               R=int(g["R"]), seed=int(g["iseed"]), timestep=int(g["timestep"]))
    ctrl = {k: v.to(cuda_device) for k, v in ctrl.items()}
    with torch.no_grad():
        y = model.apply_model(x.to(cuda_device), t.to(cuda_device), {"c_concat": [ctrl]}).cpu()
    ref = torch.from_numpy(g["out"])
    R = int(g["R"])
    assert torch.equal(y[:, :R], ref[:, :R])
    assert O.max_rel_err(y[:, R:], ref[:, R:]) < 1e-2


def test_deepcopy_and_to_like_generate_images(cuda_device):
    """generate_images.py:59-71: load on the CPU, then `copy.deepcopy(model).to(f"cuda:{i}")` per GPU."""
    from cap4d_b200 import B200MMDMUnet
    from cap4d_b200.unet import install

    g = np.load(os.path.join(GOLD, "unet_tiny_v4_h16.npz"))
    model = _reference_mmldm(int(g["wseed"]))
    first = install(model, lazy=True)            # nothing is uploaded yet
    assert first._handle.value is None
    device_model_map = {}
    n_dev = torch.cuda.device_count()
    for cuda_id in range(n_dev):
        key = f"cuda:{cuda_id}"
        device_model_map[key] = copy.deepcopy(model).to(key)
    x, t, ctrl = O.make_inputs(O.TINY_CONFIG, B=int(g["B"]), V=int(g["V"]), H=int(g["H"]), W=int(g["W"]),
                               R=int(g["R"]), seed=int(g["iseed"]), timestep=int(g["timestep"]))
    ref = torch.from_numpy(g["out"])
    R = int(g["R"])
    outs = []
    for key, m in device_model_map.items():
        unet = m.model.diffusion_model
        assert isinstance(unet, B200MMDMUnet) and unet is not first and unet.device == torch.device(key)
        assert unet._handle.value, "the copy builds its own handle when it is moved to its GPU"
        dev = torch.device(key)
        with torch.no_grad():
            y = m.apply_model(x.to(dev), t.to(dev), {"c_concat": [{k: v.to(dev) for k, v in ctrl.items()}]}).cpu()
        assert torch.equal(y[:, :R], ref[:, :R]) and O.max_rel_err(y[:, R:], ref[:, R:]) < 1e-2
        outs.append(y)
    for y in outs[1:]:
        assert torch.equal(y, outs[0]), "every GPU's copy computes the same bits"
    # a second copy of a built module, and a module without kept weights
    again = copy.deepcopy(device_model_map["cuda:0"].model.diffusion_model)
    y2 = again(x.to(cuda_device), timesteps=t.to(cuda_device), context=None,
               control={k: v.to(cuda_device) for k, v in ctrl.items()}).cpu()
    assert torch.equal(y2, outs[0])
    bare = B200MMDMUnet(O.TINY_CONFIG, O.init_state_dict(O.TINY_CONFIG, seed=0), device=cuda_device)
    with pytest.raises(RuntimeError, match="keep_state"):
        copy.deepcopy(bare)
    with pytest.raises(RuntimeError, match="no CPU path"):
        device_model_map["cuda:0"].model.diffusion_model.to("cpu")


def test_broken_ref_view_promise_is_loud(cuda_device):
    """n_ref_views=R promises ref_mask == 1 on the first R views.  A broken promise must not return plausible
    numbers: the affected views come out NaN and are counted."""
    from cap4d_b200 import B200MMDMUnet

    cfg = O.TINY_CONFIG
    unet = B200MMDMUnet(cfg, O.init_state_dict(cfg, seed=0), device=cuda_device)
    x, t, ctrl = O.make_inputs(cfg, B=2, V=4, H=8, W=8, R=1, seed=3)
    ctrl = {k: v.to(cuda_device) for k, v in ctrl.items()}
    ok = unet(x.to(cuda_device), timesteps=t.to(cuda_device), context=None, control=ctrl, n_ref_views=1)
    assert torch.isfinite(ok).all() and unet.check_ref_views() == 0
    ctrl["ref_mask"] = torch.zeros_like(ctrl["ref_mask"])  # view 0 is not a reference view after all
    bad = unet(x.to(cuda_device), timesteps=t.to(cuda_device), context=None, control=ctrl, n_ref_views=1)
    assert torch.isnan(bad[:, 0]).all() and torch.isfinite(bad[:, 1:]).all()
    assert unet.check_ref_views() == 2 and unet.check_ref_views() == 0


def test_views_per_group_must_match_time_steps(cuda_device):
    from cap4d_b200 import B200MMDMUnet

    cfg = O.TINY_CONFIG
    unet = B200MMDMUnet(cfg, O.init_state_dict(cfg, seed=0), device=cuda_device)
    V = cfg["time_steps"] + 1
    x, t, ctrl = O.make_inputs(cfg, B=1, V=V, H=8, W=8, R=1, seed=3)
    with pytest.raises(RuntimeError, match="time_steps"):
        unet(x.to(cuda_device), timesteps=t.to(cuda_device), context=None,
             control={k: v.to(cuda_device) for k, v in ctrl.items()})
