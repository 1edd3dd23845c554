"""Generate tests/golden/cond_*.npz: the UNMODIFIED reference `CAP4DConditioning.forward`
(cap4d/mmdm/conditioning/cap4dcond.py) + `PropRenderer` (mesh2img.py) run on a seeded synthetic mesh.
TEST INFRASTRUCTURE ONLY; runs only in the authoring container (needs /root/reference).

pytorch3d (pinned 0.7.8 by the reference) is not installable here, so its five imports are served by stubs:
`rasterize_meshes` is answered by oracle/cond_oracle.py:rasterize (the restatement - that part of the oracle stays
UNPINNED), `interpolate_face_attributes`, `Meshes`, `Fragments`, `load_obj` by minimal equivalents.  Everything the
reference itself does around them (template normalisation, face masks, unpacking, positional encoding, masking, area
down-sampling, channel order) runs unmodified, which pins the rest of the oracle and the CUDA kernel.

    python oracle/make_golden_cond.py            # writes the fixtures
    python oracle/make_golden_cond.py --flame    # additionally checks oracle vs reference on the FLAME template
"""
import collections
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cond_oracle as CO  # noqa: E402
from oracle import ref_import as RI  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def install_pytorch3d_stubs():
    Fragments = collections.namedtuple("Fragments", ["pix_to_face", "zbuf", "bary_coords", "dists"])

    class Meshes:
        def __init__(self, verts, faces):
            self._v, self._f = verts, faces

        def verts_padded(self):
            return self._v

        def faces_padded(self):
            return self._f

    def rasterize_meshes(meshes, image_size, blur_radius=0.0, faces_per_pixel=1, bin_size=None,
                         max_faces_per_bin=None, perspective_correct=False, clip_barycentric_coords=False,
                         cull_backfaces=False, z_clip_value=None, cull_to_frustum=False):
        assert blur_radius == 0.0 and faces_per_pixel == 1 and not perspective_correct and clip_barycentric_coords
        assert not cull_backfaces and z_clip_value is None and not cull_to_frustum
        H, W = image_size
        assert H == W
        v, f = meshes.verts_padded().cpu().numpy(), meshes.faces_padded().cpu().numpy()
        F = f.shape[1]
        p2f, zb, bc = [], [], []
        for i in range(v.shape[0]):
            a, z, b = CO.rasterize(v[i], f[i], H)
            p2f.append(np.where(a >= 0, a + i * F, -1))  # packed face indices, as pytorch3d returns them
            zb.append(z)
            bc.append(np.where((a >= 0)[..., None], b, -1.0))
        t = lambda x, dt: torch.from_numpy(np.stack(x)[:, :, :, None].astype(dt))  # noqa: E731
        return t(p2f, np.int64), t(zb, np.float32), t(bc, np.float32), t(zb, np.float32)

    def interpolate_face_attributes(pix_to_face, barycentric_coords, face_attributes):
        # pytorch3d/ops/interp_face_attrs.py: interpolate_face_attributes_python
        F, FV, D = face_attributes.shape
        N, H, W, K, _ = barycentric_coords.shape
        mask = pix_to_face < 0
        p2f = pix_to_face.clone()
        p2f[mask] = 0
        idx = p2f.view(N * H * W * K, 1, 1).expand(N * H * W * K, 3, D)
        pixel_face_vals = face_attributes.gather(0, idx).view(N, H, W, K, 3, D)
        pixel_vals = (barycentric_coords[..., None] * pixel_face_vals).sum(dim=-2)
        pixel_vals[mask] = 0
        return pixel_vals

    def load_obj(path):
        verts, uvs, fv, ft = [], [], [], []
        with open(path) as fh:
            for line in fh:
                t = line.split()
                if not t:
                    continue
                if t[0] == "v":
                    verts.append([float(x) for x in t[1:4]])
                elif t[0] == "vt":
                    uvs.append([float(x) for x in t[1:3]])
                elif t[0] == "f":
                    c = [x.split("/") for x in t[1:]]
                    for k in range(1, len(c) - 1):  # fan triangulation
                        fv.append([int(c[0][0]) - 1, int(c[k][0]) - 1, int(c[k + 1][0]) - 1])
                        if len(c[0]) > 1 and c[0][1]:
                            ft.append([int(c[0][1]) - 1, int(c[k][1]) - 1, int(c[k + 1][1]) - 1])
        Faces = collections.namedtuple("Faces", ["verts_idx", "textures_idx"])
        Aux = collections.namedtuple("Aux", ["verts_uvs"])
        ft_t = torch.tensor(ft, dtype=torch.int64) if ft else torch.full((len(fv), 3), -1, dtype=torch.int64)
        return (torch.tensor(verts, dtype=torch.float32), Faces(torch.tensor(fv, dtype=torch.int64), ft_t),
                Aux(torch.tensor(uvs, dtype=torch.float32) if uvs else None))

    mods = {}
    for name in ["pytorch3d", "pytorch3d.ops", "pytorch3d.ops.interp_face_attrs", "pytorch3d.renderer",
                 "pytorch3d.renderer.mesh", "pytorch3d.renderer.mesh.rasterizer", "pytorch3d.structures",
                 "pytorch3d.structures.meshes", "pytorch3d.io"]:
        mods[name] = types.ModuleType(name)
    mods["pytorch3d.ops.interp_face_attrs"].interpolate_face_attributes = interpolate_face_attributes
    r = mods["pytorch3d.renderer"]
    r.BlendParams = r.PerspectiveCameras = r.hard_rgb_blend = object  # imported by mesh2img.py, unused on this path
    r.rasterize_meshes = rasterize_meshes
    mods["pytorch3d.renderer.mesh.rasterizer"].Fragments = Fragments
    mods["pytorch3d.structures.meshes"].Meshes = Meshes
    mods["pytorch3d.io"].load_obj = load_obj
    sys.modules.update(mods)
    return load_obj


def write_obj(path, verts, faces):
    with open(path, "w") as fh:
        for v in verts:
            fh.write("v %.9g %.9g %.9g\n" % tuple(v))
        for f in faces:
            fh.write("f %d %d %d\n" % tuple(int(i) + 1 for i in f))


def reference_conditioning(template_path, head_vert_path, n_mouth_verts, **params):
    RI.install_stubs()
    if RI.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, RI.REFERENCE_ROOT)
    import cap4d.mmdm.conditioning.cap4dcond as CC
    import cap4d.mmdm.conditioning.mesh2img as M2I

    class Renderer(M2I.PropRenderer):  # same class, explicit asset paths instead of the cwd-relative defaults
        def __init__(self):
            super().__init__(template_path=template_path, head_vert_path=head_vert_path, n_mouth_verts=n_mouth_verts)

    orig = CC.PropRenderer
    CC.PropRenderer = Renderer
    try:
        return CC.CAP4DConditioning(**params).eval()
    finally:
        CC.PropRenderer = orig


def make_batch(template, n_views, S, seed):
    verts, offs = CO.make_views(template, n_views, seed=seed)
    rng = np.random.default_rng(seed + 1)
    ray = rng.standard_normal((n_views, 3, S, S)).astype(np.float32)
    ray /= np.linalg.norm(ray, axis=1, keepdims=True)
    ref = np.zeros((n_views, S, S), np.float32)
    ref[0] = 1.0
    crop = (rng.uniform(size=(n_views, S, S)) > 0.1).astype(np.float32)
    return verts, offs, ray, ref, crop


def run_reference(cond, verts, offs, ray, ref, crop):
    batch = {"verts_2d": torch.from_numpy(verts)[None], "offsets_3d": torch.from_numpy(offs)[None],
             "reference_mask": torch.from_numpy(ref)[None], "ray_map": torch.from_numpy(ray)[None],
             "out_crop_mask": torch.from_numpy(crop)[None], "z": torch.zeros(1, verts.shape[0], 4, 4, 4)}
    with torch.no_grad():
        out = cond(batch, unconditional=False)
        unc = cond(batch, unconditional=True)
    assert float(unc["pos_enc"].abs().max()) == 0.0 and unc["pos_enc"].shape == out["pos_enc"].shape
    return out["pos_enc"][0].numpy(), out["ref_mask"][0].numpy()


def golden(name, n_lat, n_lon, n_views, S, sr, seed, use_crop=True, n_mouth=40):
    tv, faces, _ = CO.make_mesh(n_lat, n_lon, seed=seed)
    head = np.nonzero(tv[:, 1] > -0.6)[0]  # plays head_vertices.txt
    with tempfile.TemporaryDirectory() as d:
        obj, hv = os.path.join(d, "t.obj"), os.path.join(d, "head.txt")
        write_obj(obj, tv, faces)
        np.savetxt(hv, head, fmt="%d")
        cond = reference_conditioning(obj, hv, n_mouth, image_size=S, positional_channels=42, positional_multiplier=1.,
                                      super_resolution=sr, use_ray_directions=True, use_expr_deformation=True,
                                      use_crop_mask=use_crop)
    props = cond.renderer.props.numpy()
    fmask = cond.renderer.face_mask.numpy()
    rfaces = cond.renderer.faces.numpy().astype(np.int32)
    assert np.array_equal(rfaces, faces)
    verts, offs, ray, ref, crop = make_batch(tv, n_views, S, seed + 10)
    pos_enc, ref_mask = run_reference(cond, verts, offs, ray, ref, crop)
    mine = CO.cond_pos_enc(verts, offs, faces, props, fmask, ray, ref, crop if use_crop else None, S, sr, 42, 1.0,
                           cond.std_expr_deformation)
    err = float(np.abs(mine - pos_enc).max())
    print(f"{name}: pos_enc {pos_enc.shape}, coverage {float((np.abs(pos_enc[..., :42]).sum(-1) > 0).mean()):.2f}, "
          f"oracle vs reference max abs err {err:.3e}")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), verts_2d=verts, offsets_3d=offs, faces=faces, props=props,
                        face_mask=fmask, ray_map=ray, ref_mask=ref, crop_mask=crop, image_size=S,
                        super_resolution=sr, std_expr_deformation=cond.std_expr_deformation, use_crop=use_crop,
                        pos_enc=pos_enc, ref_mask_out=ref_mask)


def check_flame(n_views=2):
    """Oracle vs the reference code on the real FLAME template (a reference asset: read in place, not committed)."""
    base = os.path.join(RI.REFERENCE_ROOT, "data", "assets", "flame")
    cond = reference_conditioning(os.path.join(base, "cap4d_flame_template.obj"),
                                  os.path.join(base, "head_vertices.txt"), 200, image_size=64, positional_channels=42,
                                  positional_multiplier=1., super_resolution=2, use_ray_directions=True,
                                  use_expr_deformation=True, use_crop_mask=True)
    faces = cond.renderer.faces.numpy().astype(np.int32)
    props = cond.renderer.props.numpy()
    fmask = cond.renderer.face_mask.numpy()
    tv = props * 1.0
    verts, offs, ray, ref, crop = make_batch(tv, n_views, 64, 5)
    pos_enc, _ = run_reference(cond, verts, offs, ray, ref, crop)
    mine = CO.cond_pos_enc(verts, offs, faces, props, fmask, ray, ref, crop, 64, 2, 42, 1.0, cond.std_expr_deformation)
    print(f"FLAME template: {props.shape[0]} verts, {faces.shape[0]} faces ({int(fmask.sum())} unmasked); "
          f"oracle vs reference max abs err {float(np.abs(mine - pos_enc).max()):.3e}")


def golden_rays(name="cond_rays", n_views=4, S=64, seed=7):
    """load_camera_rays / verts_to_pytorch3d of the unmodified reference (cap4d/datasets/utils.py:79-89,161-186) and
    the rotation of inference_data.py:89-100 on seeded cameras."""
    if "decord" not in sys.modules:  # imported at the top of cap4d/datasets/utils.py, unused on this path
        dec = types.ModuleType("decord")
        dec.VideoReader = object
        sys.modules["decord"] = dec
    if RI.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, RI.REFERENCE_ROOT)
    import einops
    import cap4d.datasets.utils as U

    rng = np.random.default_rng(seed)

    def rot():
        q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        return q * np.sign(np.linalg.det(q))

    ref_extr = np.eye(4)
    ref_extr[:3, :3] = rot()
    boxes, intrs, extrs, rays, v_in, v_out = [], [], [], [], [], []
    for _ in range(n_views):
        x0, y0 = rng.integers(-40, 200, size=2)
        side = int(rng.integers(300, 700))
        crop_box = (int(x0), int(y0), int(x0) + side, int(y0) + side)
        intr = np.eye(3)
        intr[0, 0], intr[1, 1] = rng.uniform(1500, 3000, size=2)
        intr[0, 2], intr[1, 2] = rng.uniform(200, 400, size=2)
        extr = np.eye(4)
        extr[:3, :3] = rot()
        extr[:3, 3] = rng.standard_normal(3)
        ray = U.load_camera_rays(crop_box, intr, extr, S)
        h = ray.shape[1]
        ray = einops.rearrange(ray, 'v h w -> v (h w)')
        ray = ref_extr[:3, :3] @ ray
        ray = einops.rearrange(ray, 'v (h w) -> v h w', h=h)
        v = rng.uniform(0, 800, size=(50, 3)).astype(np.float32)
        v_in.append(v.copy())
        v_out.append(U.verts_to_pytorch3d(v.copy(), np.array(crop_box)))
        boxes.append(crop_box), intrs.append(intr), extrs.append(extr), rays.append(ray)
    rays = np.stack(rays)
    mine = np.stack([CO.ray_map(b, i, e, ref_extr, S) for b, i, e in zip(boxes, intrs, extrs)])
    mv = np.stack([CO.verts_to_pytorch3d(v, np.array(b)) for v, b in zip(v_in, boxes)])
    print(f"{name}: rays {rays.shape} {rays.dtype}, oracle vs reference max abs err {float(np.abs(mine - rays).max()):.3e}, "
          f"verts_to_pytorch3d err {float(np.abs(mv - np.stack(v_out)).max()):.3e}")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), crop_boxes=np.asarray(boxes), intr=np.stack(intrs),
                        extr=np.stack(extrs), ref_extr=ref_extr, S=S, rays=rays, verts_in=np.stack(v_in),
                        verts_out=np.stack(v_out))


if __name__ == "__main__":
    install_pytorch3d_stubs()
    golden_rays()
    os.makedirs(OUT, exist_ok=True)
    golden("cond_sr2_s32", n_lat=24, n_lon=32, n_views=3, S=32, sr=2, seed=0)
    golden("cond_sr1_s24_nocrop", n_lat=12, n_lon=16, n_views=2, S=24, sr=1, seed=3, use_crop=False)
    if "--flame" in sys.argv:
        check_flame()
