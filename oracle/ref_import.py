"""Import the reference (/root/reference) in the authoring container.  TEST INFRASTRUCTURE ONLY.

The reference needs `omegaconf` and `pytorch_lightning`, neither of which is installed here; the
hot path uses them only for an isinstance check (openaimodel.py:479-483) and as a base class
(ddpm.py:12,21 / autoencoder.py:2).  Two stub modules are enough.  /root/reference
does not exist on the GPU box; there the verbatim copy under oracle/_ref (oracle/build_ref.py) is imported.
"""
import os
import sys
import types

import torch

_REF_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")  # made by oracle/build_ref.py


def _find_root() -> str:
    """The reference tree itself where it exists (authoring container), else the verbatim copy of its hot-path
    modules that oracle/build_ref.py placed under oracle/_ref (the GPU box)."""
    for cand in (os.environ.get("CAP4D_REFERENCE_ROOT"), "/root/reference", _REF_COPY):
        if cand and os.path.isdir(os.path.join(cand, "cap4d", "mmdm")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "cap4d", "mmdm"))


def install_stubs() -> None:
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")
        lc = types.ModuleType("omegaconf.listconfig")

        class ListConfig(list):
            pass

        oc.ListConfig = ListConfig
        lc.ListConfig = ListConfig
        oc.listconfig = lc
        sys.modules["omegaconf"] = oc
        sys.modules["omegaconf.listconfig"] = lc
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(torch.nn.Module):
            # Lightning's DeviceDtypeModuleMixin: `.device` is what the last .to() / .cuda() / .cpu() set
            _stub_device = None

            @property
            def device(self):
                if self._stub_device is not None:
                    return self._stub_device
                try:
                    return next(self.parameters()).device
                except StopIteration:
                    return torch.device("cpu")

            def to(self, *args, **kwargs):
                device = torch._C._nn._parse_to(*args, **kwargs)[0]
                if device is not None:
                    self._stub_device = torch.device(device)
                return super().to(*args, **kwargs)

            def cuda(self, device=None):
                self._stub_device = torch.device("cuda", torch.cuda.current_device() if device is None else
                                                 torch.device(device).index if not isinstance(device, int) else device)
                return super().cuda(device)

            def cpu(self):
                self._stub_device = torch.device("cpu")
                return super().cpu()

            def log(self, *a, **k):
                pass

            def log_dict(self, *a, **k):
                pass

        pl.LightningModule = LightningModule
        util = types.ModuleType("pytorch_lightning.utilities")
        rz = types.ModuleType("pytorch_lightning.utilities.rank_zero")
        rz.rank_zero_only = lambda fn: fn
        util.rank_zero = rz
        pl.utilities = util
        sys.modules["pytorch_lightning"] = pl
        sys.modules["pytorch_lightning.utilities"] = util
        sys.modules["pytorch_lightning.utilities.rank_zero"] = rz


def import_reference():
    """Returns (MMDMUnetModel, StochasticIOSampler, MMLDM) classes of the unmodified reference."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from cap4d.mmdm.net.mmdm_unet import MMDMUnetModel
    from cap4d.mmdm.sampler import StochasticIOSampler
    from cap4d.mmdm.mmdm import MMLDM

    return MMDMUnetModel, StochasticIOSampler, MMLDM


def build_reference_unet(cfg: dict):
    """MMDMUnetModel with the fixed MMDM settings of configs/mmdm/cap4d_mmdm_final.yaml:95-115."""
    MMDMUnetModel, _, _ = import_reference()
    return MMDMUnetModel(
        image_size=64,
        time_steps=cfg["time_steps"],
        temporal_mode="3d",
        in_channels=cfg["in_channels"],
        out_channels=cfg["out_channels"],
        model_channels=cfg["model_channels"],
        condition_channels=cfg["condition_channels"],
        attention_resolutions=list(cfg["attention_resolutions"]),
        num_res_blocks=cfg["num_res_blocks"],
        channel_mult=list(cfg["channel_mult"]),
        num_head_channels=cfg["num_head_channels"],
        use_spatial_transformer=True,
        use_linear_in_transformer=True,
        transformer_depth=1,
        context_dim=1024,
        use_checkpoint=False,
        legacy=False,
    ).eval()


def build_reference_mmldm(cfg: dict):
    """MMLDM with an identity first stage and no conditioning stage (stub 2)."""
    _, _, MMLDM = import_reference()
    unet_params = dict(
        image_size=64, time_steps=cfg["time_steps"], temporal_mode="3d", in_channels=cfg["in_channels"],
        out_channels=cfg["out_channels"], model_channels=cfg["model_channels"],
        condition_channels=cfg["condition_channels"], attention_resolutions=list(cfg["attention_resolutions"]),
        num_res_blocks=cfg["num_res_blocks"], channel_mult=list(cfg["channel_mult"]),
        num_head_channels=cfg["num_head_channels"], use_spatial_transformer=True, use_linear_in_transformer=True,
        transformer_depth=1, context_dim=1024, use_checkpoint=False, legacy=False,
    )
    model = MMLDM(
        control_key="hint", only_mid_control=False, n_frames=8,
        shift_schedule=True, zero_snr_shift=True, sqrt_shift=True, minus_one_shift=True,
        linear_start=0.00085, linear_end=0.0120, num_timesteps_cond=1, log_every_t=200, timesteps=1000,
        first_stage_key="jpg", cond_stage_key="txt", image_size=64, channels=4, cond_stage_trainable=False,
        conditioning_key="crossattn", scale_factor=0.18215, use_ema=False,
        unet_config={"target": "cap4d.mmdm.net.mmdm_unet.MMDMUnetModel", "params": unet_params},
        first_stage_config={"target": "controlnet.ldm.models.autoencoder.IdentityFirstStage"},
        cond_stage_config="__is_unconditional__",
    )
    return model.eval()


def build_reference_vae(cfg: dict):
    """AutoencoderKL of configs/mmdm/cap4d_mmdm_final.yaml:117-139 (controlnet/ldm/models/autoencoder.py)."""
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from controlnet.ldm.models.autoencoder import AutoencoderKL

    dd = dict(double_z=True, z_channels=cfg["z_channels"], resolution=512, in_channels=3, out_ch=cfg["out_ch"],
              ch=cfg["ch"], ch_mult=list(cfg["ch_mult"]), num_res_blocks=cfg["num_res_blocks"], attn_resolutions=[],
              dropout=0.0)
    return AutoencoderKL(ddconfig=dd, lossconfig={"target": "torch.nn.Identity"}, embed_dim=cfg["embed_dim"]).eval()
