"""B200 conditioning maps: the reference's `CAP4DConditioning` (cap4d/mmdm/conditioning/cap4dcond.py:42-139) with its
`PropRenderer` (cap4d/mmdm/conditioning/mesh2img.py:314-379) behind the same constructor and `forward(batch,
unconditional)` call, on libcap4d_b200.so - one kernel per call instead of pytorch3d's rasteriser + a dozen eager
ops on super-resolved tensors.  pytorch3d is not needed (nor is its `load_obj`: `load_template` reads the .obj).

    cond = B200CAP4DConditioning.from_reference(model.cond_stage_model)     # takes faces / props / face_mask
    model.cond_stage_model = cond                                           # or install_conditioning(model)
    c = cond(batch["hint"], unconditional=False)    # {"pos_enc": [B,T,S,S,50], "z_input": ..., "ref_mask": ...}

There is no CPU path: `forward(unconditional=False)` needs CUDA tensors and the shared library.
"""
from __future__ import annotations

import ctypes
from typing import Mapping, Optional, Sequence

import numpy as np
import torch

from . import _lib


def load_template(template_path: str, head_vert_path: str, n_mouth_verts: int = 200):
    """PropRenderer.__init__ (mesh2img.py:319-366, prop_type="verts") without pytorch3d: returns
    (faces int32 [F,3], props fp32 [Nv,3], face_mask bool [F]) as torch CPU tensors."""
    verts, faces = [], []
    with open(template_path) as fh:
        for line in fh:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                verts.append([float(x) for x in t[1:4]])
            elif t[0] == "f":
                idx = [int(x.split("/")[0]) for x in t[1:]]
                idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
                for k in range(1, len(idx) - 1):
                    faces.append([idx[0], idx[k], idx[k + 1]])
    verts_t = torch.tensor(verts, dtype=torch.float32)
    faces_t = torch.tensor(faces, dtype=torch.int64)
    vert_mask = torch.zeros(verts_t.shape[0]).bool()
    head_verts = torch.tensor(np.genfromtxt(head_vert_path)).long()
    vert_mask[head_verts] = 1
    vert_mask[-n_mouth_verts:] = 1
    face_mask = vert_mask[faces_t].max(dim=-1)[0]
    props = verts_t - verts_t.mean(dim=-2, keepdim=True)
    props = props / props.max()
    return faces_t.to(torch.int32), props, face_mask


class B200CAP4DConditioning(torch.nn.Module):
    """Same parameters as CAP4DConditioning.__init__ (cap4dcond.py:43-65) plus the renderer's three buffers."""

    def __init__(self, faces: torch.Tensor, props: torch.Tensor, face_mask: torch.Tensor, image_size: int = 64,
                 positional_channels: int = 42, positional_multiplier: float = 1., super_resolution: int = 2,
                 use_ray_directions: bool = True, use_expr_deformation: bool = True, use_crop_mask: bool = False,
                 std_expr_deformation: float = 0.0104) -> None:
        super().__init__()
        assert super_resolution >= 1 and super_resolution % 1 == 0
        assert positional_channels % 3 == 0
        assert (positional_channels // 3) % 2 == 0
        if super_resolution not in (1, 2, 4):
            raise ValueError("cap4d_b200: super_resolution must be 1, 2 or 4")
        self.image_size = image_size
        self.super_resolution = int(super_resolution)
        self.positional_channels = positional_channels
        self.positional_multiplier = positional_multiplier
        self.use_ray_directions = use_ray_directions
        self.use_expr_deformation = use_expr_deformation
        self.std_expr_deformation = std_expr_deformation
        self.use_crop_mask = use_crop_mask
        self.register_buffer("faces", faces.detach().to(torch.int32).contiguous())
        self.register_buffer("props", props.detach().to(torch.float32).contiguous())
        self.register_buffer("face_mask", face_mask.detach().to(torch.uint8).contiguous())
        self._ws = None  # per-face tile ranges of the kernel's pre-pass (grown on demand)
        if self.faces.dim() != 2 or self.faces.shape[1] != 3 or self.props.shape[1] != 3:
            raise ValueError("faces must be [F,3] and props [Nv,3]")
        if int(self.faces.min()) < 0 or int(self.faces.max()) >= self.props.shape[0]:
            raise ValueError("faces index outside the vertex array")

    @classmethod
    def from_reference(cls, cond_stage_model) -> "B200CAP4DConditioning":
        r = cond_stage_model.renderer
        return cls(r.faces, r.props, r.face_mask, image_size=cond_stage_model.image_size,
                   positional_channels=cond_stage_model.positional_channels,
                   positional_multiplier=cond_stage_model.positional_multiplier,
                   super_resolution=cond_stage_model.super_resolution,
                   use_ray_directions=cond_stage_model.use_ray_directions,
                   use_expr_deformation=cond_stage_model.use_expr_deformation,
                   use_crop_mask=cond_stage_model.use_crop_mask,
                   std_expr_deformation=cond_stage_model.std_expr_deformation)

    @property
    def total_channels(self) -> int:
        # cap4dcond.py:79-86
        return (self.positional_channels + 1 + (1 if self.use_crop_mask else 0) + (3 if self.use_ray_directions else 0)
                + (3 if self.use_expr_deformation else 0))

    def forward(self, batch: Mapping[str, torch.Tensor], unconditional: bool = True):
        verts = batch["verts_2d"]
        offsets = batch["offsets_3d"]
        ref_mask = batch["reference_mask"][:, :, None]
        B, T = verts.shape[:2]
        z_input = batch["z"] if "z" in batch else None
        S = self.image_size
        if unconditional:
            pos_enc = torch.zeros((B, T, S, S, self.total_channels), device=verts.device)
            if z_input is not None:
                z_input = z_input * 0.
            return {"pos_enc": pos_enc, "z_input": z_input, "ref_mask": ref_mask}
        pos_enc = self.render_pos_enc(
            verts.reshape(B * T, *verts.shape[2:]), offsets.reshape(B * T, *offsets.shape[2:]),
            batch["reference_mask"].reshape(B * T, S, S),
            batch["ray_map"].reshape(B * T, 3, S, S) if self.use_ray_directions else None,
            batch["out_crop_mask"].reshape(B * T, S, S) if self.use_crop_mask else None)
        return {"pos_enc": pos_enc.view(B, T, S, S, -1), "z_input": z_input, "ref_mask": ref_mask}

    def get_vis(self, enc):
        """CAP4DConditioning.get_vis (cap4dcond.py:141-171): channel slices used by `log_cond`."""
        vis = {}
        n_pos = self.positional_channels // 3
        pos_enc = enc[..., 0:self.positional_channels]
        for i in range(n_pos - 2, n_pos):
            vis[f"pose_map_{i}"] = pos_enc[..., [i, i + n_pos, i + n_pos * 2]]
        counter = self.positional_channels
        if self.use_expr_deformation:
            vis["expr_disp"] = enc[..., counter:counter + 3]
            counter += 3
        if self.use_ray_directions:
            vis["ray_map"] = enc[..., counter:counter + 3]
            counter += 3
        vis["ref_mask"] = enc[..., [counter] * 3]
        counter += 1
        if self.use_crop_mask:
            vis["crop_mask"] = enc[..., [counter] * 3]
            counter += 1
        return vis

    def render_pos_enc(self, verts_2d: torch.Tensor, offsets_3d: Optional[torch.Tensor], ref_mask: torch.Tensor,
                       ray_map: Optional[torch.Tensor] = None, crop_mask: Optional[torch.Tensor] = None,
                       return_pix_to_face: bool = False):
        """Flattened views: verts_2d/offsets_3d [n,Nv,3], ref_mask/crop_mask [n,S,S], ray_map [n,3,S,S] (CUDA).
        Returns pos_enc fp32 [n,S,S,C] (and Fragments.pix_to_face int32 [n,S*sr,S*sr] on request)."""
        if not verts_2d.is_cuda:
            raise RuntimeError("cap4d_b200: conditioning maps are rendered on the GPU; there is no CPU path")
        dev = verts_2d.device
        if self.faces.device != dev:
            self.to(dev)
        f32 = lambda t: None if t is None else t.detach().to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
        verts_2d, ref_mask, ray_map, crop_mask = f32(verts_2d), f32(ref_mask), f32(ray_map), f32(crop_mask)
        offsets_3d = f32(offsets_3d) if self.use_expr_deformation else None
        n, nv = verts_2d.shape[0], verts_2d.shape[1]
        S, sr = self.image_size, self.super_resolution
        if nv != self.props.shape[0] or verts_2d.shape[2] != 3:
            raise ValueError(f"verts_2d must be [n,{self.props.shape[0]},3], got {tuple(verts_2d.shape)}")
        if offsets_3d is not None and offsets_3d.shape != verts_2d.shape:
            raise ValueError("offsets_3d must have the shape of verts_2d")
        if tuple(ref_mask.shape) != (n, S, S):
            raise ValueError(f"reference_mask must be [n,{S},{S}]")
        C = self.positional_channels + (3 if offsets_3d is not None else 0) + (3 if ray_map is not None else 0) + 1 \
            + (1 if crop_mask is not None else 0)
        out = torch.empty((n, S, S, C), device=dev, dtype=torch.float32)
        p2f = torch.empty((n, S * sr, S * sr), device=dev, dtype=torch.int32) if return_pix_to_face else None
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p()  # noqa: E731
        if n == 0:
            return (out, p2f) if return_pix_to_face else out
        nbytes = ctypes.c_size_t()
        _lib.check(_lib.load().cap4d_b200_cond_workspace_bytes(n, self.faces.shape[0], ctypes.byref(nbytes)),
                   "cond_workspace_bytes")
        ws = self._ws
        if ws is None or ws.device != dev or ws.numel() < nbytes.value:
            ws = self._ws = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().cap4d_b200_cond_pos_enc(
                ptr(verts_2d), ptr(offsets_3d), ptr(self.faces), ptr(self.props), ptr(self.face_mask), ptr(ray_map),
                ptr(ref_mask), ptr(crop_mask), ptr(out), ptr(p2f), n, nv, self.faces.shape[0], S, sr,
                self.positional_channels, float(self.positional_multiplier), float(self.std_expr_deformation),
                ptr(ws), ws.numel(), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "cond_pos_enc")
        return (out, p2f) if return_pix_to_face else out


def camera_rows(crop_boxes: Sequence, intrinsics: Sequence, extrinsics: Sequence, ref_extr,
                target_resolution: int) -> np.ndarray:
    """Host part of load_camera_rays (cap4d/datasets/utils.py:161-186): fp64 [n,22] = the cropped intrinsics
    (utils.py:169-173), inv(extr[:3,:3]) (utils.py:183) and ref_extr[:3,:3] (inference_data.py:99) per view."""
    rows = []
    ref_r = np.asarray(ref_extr, np.float64)[:3, :3].reshape(-1)
    for crop_box, intr, extr in zip(crop_boxes, intrinsics, extrinsics):
        scale = target_resolution / (crop_box[2] - crop_box[0])
        rows.append(np.concatenate([
            np.asarray([intr[0, 0] * scale, intr[1, 1] * scale, (intr[0, 2] - crop_box[0]) * scale,
                        (intr[1, 2] - crop_box[1]) * scale], np.float64),
            np.linalg.inv(np.asarray(extr)[:3, :3]).reshape(-1), ref_r]))
    return np.asarray(rows, np.float64).reshape(-1, 22)


def ray_maps(cam_rows, target_resolution: int, device) -> torch.Tensor:
    """load_camera_rays + rotation into the reference camera frame for n views on the device: fp32 [n,3,S,S]."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("cap4d_b200: ray maps are computed on the GPU; there is no CPU path")
    cam = torch.as_tensor(np.asarray(cam_rows, np.float64)).to(dev).contiguous()
    n = cam.shape[0]
    out = torch.empty((n, 3, target_resolution, target_resolution), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().cap4d_b200_cond_ray_map(
            ctypes.c_void_p(cam.data_ptr()), ctypes.c_void_p(out.data_ptr()), n, target_resolution,
            ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "cond_ray_map")
    return out


def install_conditioning(mmldm) -> B200CAP4DConditioning:
    """Replace `mmldm.cond_stage_model` (a reference CAP4DConditioning) in place; `MMLDM.get_learned_conditioning` /
    `get_unconditional_conditioning` keep calling it with the same arguments."""
    cond = B200CAP4DConditioning.from_reference(mmldm.cond_stage_model)
    mmldm.cond_stage_model = cond
    return cond


@torch.no_grad()
def get_condition_from_dataloader(cond_stage_model, vae, dataloader, device, first_stage_key: str = "jpg",
                                  control_key: str = "hint", to_cpu: bool = False, visualize: bool = False):
    """get_condition_from_dataloader(model, dataloader, device) of cap4d/inference/utils.py:64-100 with the model's
    two stages passed explicitly: `vae.encode_first_stage` plays MMLDM.get_input's encode (cap4d/mmdm/mmdm.py:47-66,
    every frame, same RNG use) and `cond_stage_model(batch, unconditional=...)` its get_learned / get_unconditional
    conditioning.  Returns the reference's dict ("cond_frames", "uncond_frames", "cond_vis_frames", "flame_params");
    the frames stay on `device` unless to_cpu=True (the reference moves every frame to host memory and the sampler
    uploads them again - B200StochasticIOSampler takes device tensors as they are)."""
    from collections import defaultdict

    cond_frames, uncond_frames, cond_vis_frames = defaultdict(list), defaultdict(list), defaultdict(list)
    flame_params = []
    dev = torch.device(device)
    for batch in dataloader:
        hint = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch[control_key].items()}
        x = batch[first_stage_key].to(dev)
        if x.dim() == 3:
            x = x[..., None]
        x = x.permute(0, 1, 4, 2, 3).contiguous()                       # 'b t h w c -> b t c h w'
        hint["z"] = vae.encode_first_stage(x)                           # [b, t, 4, h/8, w/8]
        c_uncond = cond_stage_model(hint, unconditional=True)
        c_cond = cond_stage_model(hint, unconditional=False)
        for key in c_cond:
            cc = c_cond[key].reshape(-1, *c_cond[key].shape[2:])        # 'b t ... -> (b t) ...'
            cu = c_uncond[key].reshape(-1, *c_uncond[key].shape[2:])
            cond_frames[key].append(cc.cpu() if to_cpu else cc)
            uncond_frames[key].append(cu.cpu() if to_cpu else cu)
        if visualize:  # log_cond (utils.py:26-41)
            vis = cond_stage_model.get_vis(c_cond["pos_enc"])
            for key, v in vis.items():
                b_ = v.shape[0]
                v = v.reshape(-1, *v.shape[2:]).permute(0, 3, 1, 2)
                v = torch.nn.functional.interpolate(v, scale_factor=8., mode="nearest").clamp(-1., 1.)
                cond_vis_frames[key].append(v.permute(0, 2, 3, 1).cpu())                # '(b t) c h w -> (b t) h w c'
                del b_
        if "flame_params" in batch:
            fp = batch["flame_params"]
            for b in range(next(iter(fp.values())).shape[0]):
                flame_params.append({k: (fp[k][b].cpu().numpy() if torch.is_tensor(fp[k]) else np.asarray(fp[k][b]))
                                     for k in fp})
    return {"cond_frames": cond_frames, "uncond_frames": uncond_frames, "cond_vis_frames": cond_vis_frames,
            "flame_params": flame_params}


def concat_frames(frames) -> dict:
    """generate_images.py:104-108: one tensor per key."""
    return {k: torch.cat(v, dim=0) for k, v in frames.items()}
