"""CPU restatement (numpy fp32) of the conditioning-map step: TEST INFRASTRUCTURE, never the product path.

Only tests/, __graft_entry__.smoke() and bench scripts' cpu-baseline legs may import this module.

What it follows
  * CAP4DConditioning.forward(unconditional=False|True)   cap4d/mmdm/conditioning/cap4dcond.py:66-139
  * PositionalEncoding.forward                             cap4dcond.py:24-39
  * PropRenderer.__init__/render                           cap4d/mmdm/conditioning/mesh2img.py:314-379
  * VertexShader._get_fragments (cameras=None)             mesh2img.py:164-192
  * load_camera_rays + rotation into the reference frame   cap4d/datasets/utils.py:161-186,
                                                           cap4d/inference/data/inference_data.py:89-100
  * verts_to_pytorch3d                                     cap4d/datasets/utils.py:79-89

Third-party algorithm: the rasterisation itself lives in pytorch3d (pinned `pytorch3d==0.7.8`,
environment.yml:104), which is NOT vendored in the reference and NOT installable here.  `rasterize()` restates its
published algorithm (pytorch3d/csrc/rasterize_meshes/rasterize_meshes.cu: CheckPixelInsideFace,
rasterization_utils.cuh: PixToNonSquareNdc, utils/geometry_utils.cuh: EdgeFunctionForward,
BarycentricCoordsForward, BarycentricClipForward) for the one configuration the reference calls it in:
blur_radius 0, faces_per_pixel 1, perspective_correct False, clip_barycentric_coords True, cull_backfaces False,
z_clip_value None, cull_to_frustum False.  **Parity of `rasterize()` is unpinned** (no pytorch3d to run, no
golden vectors in the reference); everything downstream of the fragments IS pinned: tests/golden/cond_*.npz hold the
output of the unmodified reference `CAP4DConditioning.forward` run with its renderer's rasteriser call served by
`rasterize()` (oracle/make_golden_cond.py).
"""
from __future__ import annotations

import numpy as np

K_EPSILON = np.float32(1e-8)
F32 = np.float32


def pix_to_ndc(i: np.ndarray, S: int) -> np.ndarray:
    """PixToNonSquareNdc(i, S, S): NDC coordinate of the centre of pixel i (fp32 arithmetic)."""
    return F32(-1.0) + (F32(2.0) * i.astype(F32) + F32(1.0)) / F32(S)


def _edge(px, py, ax, ay, bx, by):
    """EdgeFunctionForward(p, v0, v1); every product and difference rounded to fp32."""
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax)


def rasterize(verts: np.ndarray, faces: np.ndarray, size: int):
    """rasterize_meshes(blur_radius=0, faces_per_pixel=1, ...) for ONE mesh, square image `size`.

    verts fp32 [Nv,3] (x, y NDC with +X left, +Y up; z depth), faces int [F,3].
    Returns pix_to_face int32 [size,size] (-1 = none), zbuf fp32, bary fp32 [size,size,3] (clipped)."""
    verts = np.ascontiguousarray(verts, dtype=F32)
    S = int(size)
    p2f = np.full((S, S), -1, np.int32)
    zbuf = np.full((S, S), np.inf, F32)
    bary = np.zeros((S, S, 3), F32)
    # row yi samples NDC(S-1-yi), column xi samples NDC(S-1-xi)
    ndc = pix_to_ndc(np.arange(S - 1, -1, -1), S)  # ndc[i] = coordinate of row/column i (decreasing)
    for f in range(faces.shape[0]):
        i0, i1, i2 = faces[f]
        x0, y0, z0 = verts[i0]
        x1, y1, z1 = verts[i1]
        x2, y2, z2 = verts[i2]
        xmin, xmax = min(x0, x1, x2), max(x0, x1, x2)
        ymin, ymax = min(y0, y1, y2), max(y0, y1, y2)
        zmin = min(z0, z1, z2)
        if zmin < K_EPSILON:  # CheckPointOutsideBoundingBox: z_invalid
            continue
        face_area = _edge(x0, y0, x1, y1, x2, y2)
        if -K_EPSILON <= face_area <= K_EPSILON:
            continue
        # window of rows/columns whose centre is inside the bounding box (exact fp32 test, as the reference)
        cols = np.nonzero(~((ndc > xmax) | (ndc < xmin)))[0]
        rows = np.nonzero(~((ndc > ymax) | (ndc < ymin)))[0]
        if cols.size == 0 or rows.size == 0:
            continue
        px = ndc[cols][None, :]
        py = ndc[rows][:, None]
        area = _edge(x2, y2, x0, y0, x1, y1) + K_EPSILON
        w0 = _edge(px, py, x1, y1, x2, y2) / area
        w1 = _edge(px, py, x2, y2, x0, y0) / area
        w2 = _edge(px, py, x0, y0, x1, y1) / area
        inside = (w0 > 0) & (w1 > 0) & (w2 > 0)
        if not inside.any():
            continue
        c0, c1, c2 = np.maximum(w0, F32(0)), np.maximum(w1, F32(0)), np.maximum(w2, F32(0))
        wsum = np.maximum(c0 + c1 + c2, F32(1e-5))
        c0, c1, c2 = c0 / wsum, c1 / wsum, c2 / wsum
        pz = c0 * z0 + c1 * z1 + c2 * z2
        win = (slice(rows[0], rows[-1] + 1), slice(cols[0], cols[-1] + 1))
        cur = zbuf[win]
        # faces are visited in index order, so a strict `<` keeps the smaller face index on depth ties
        take = inside & (pz >= 0) & (pz < cur)
        if take.any():
            zbuf[win] = np.where(take, pz, cur)
            p2f[win] = np.where(take, np.int32(f), p2f[win])
            b = bary[win]
            b[take] = np.stack([c0, c1, c2], -1)[take]
    zbuf[p2f < 0] = -1.0
    return p2f, zbuf, bary


def interpolate_face_attributes(p2f: np.ndarray, bary: np.ndarray, face_attrs: np.ndarray) -> np.ndarray:
    """pytorch3d.ops.interpolate_face_attributes for K = 1: (bary[..., None] * attrs[pix_to_face]).sum(-2); 0 where
    pix_to_face < 0.  face_attrs fp32 [F,3,D]."""
    a = face_attrs[np.clip(p2f, 0, None)]  # [S,S,3,D]
    out = bary[..., 0, None] * a[..., 0, :] + bary[..., 1, None] * a[..., 1, :] + bary[..., 2, None] * a[..., 2, :]
    out = out.astype(F32)
    out[p2f < 0] = 0
    return out


def positional_encoding(x: np.ndarray, n_freq: int) -> np.ndarray:
    """PositionalEncoding.forward (cap4dcond.py:24-39): [..., 3] -> [..., 3 * 2 * n_freq], (c f) order with
    f = [sin(2^0 x) .. sin(2^(n-1) x), cos(2^0 x) .. cos(2^(n-1) x)]."""
    freqs = (F32(2.0) ** np.arange(n_freq, dtype=F32)).astype(F32)
    arg = x[..., None].astype(F32) * freqs
    emb = np.concatenate([np.sin(arg), np.cos(arg)], -1).astype(F32)
    return emb.reshape(*x.shape[:-1], -1)


def area_downsample(x: np.ndarray, sr: int) -> np.ndarray:
    """F.interpolate(mode="area") from (S*sr) to S on [H,W,C]: mean of sr x sr blocks, summed in raster order."""
    H, W, C = x.shape
    S = H // sr
    acc = np.zeros((S, S, C), F32)
    for sy in range(sr):
        for sx in range(sr):
            acc = acc + x[sy::sr, sx::sr]
    return (acc / F32(sr * sr)).astype(F32)


def cond_pos_enc(verts_2d, offsets_3d, faces, props, face_mask, ray_map, ref_mask, crop_mask, image_size=64,
                 super_resolution=2, positional_channels=42, positional_multiplier=1.0, std_expr_deformation=0.0104,
                 return_fragments=False):
    """CAP4DConditioning.forward(unconditional=False) for flattened views.

    verts_2d/offsets_3d [n,Nv,3]; ray_map [n,3,S,S] or None; ref_mask [n,S,S]; crop_mask [n,S,S] or None.
    Returns pos_enc fp32 [n,S,S,C]."""
    n = verts_2d.shape[0]
    S, sr = image_size, super_resolution
    n_freq = positional_channels // 6
    outs, frags = [], []
    face_props = np.asarray(props, F32)[faces]  # [F,3,3]   mesh2img.py:341
    for v in range(n):
        p2f, _, bary = rasterize(verts_2d[v], faces, S * sr)
        frags.append(p2f)
        img = interpolate_face_attributes(p2f, bary, face_props)
        enc = positional_encoding(img * F32(positional_multiplier), n_freq)
        if offsets_3d is not None:
            off = (np.asarray(offsets_3d[v], F32) / F32(std_expr_deformation)).astype(F32)  # cap4dcond.py:94
            enc = np.concatenate([enc, interpolate_face_attributes(p2f, bary, off[faces])], -1)
        mask = (p2f != -1) & np.asarray(face_mask, bool)[np.clip(p2f, 0, None)]  # mesh2img.py:374-377
        enc = enc * mask[..., None].astype(F32)
        enc = area_downsample(enc, sr)
        parts = [enc]
        if ray_map is not None:
            parts.append(np.transpose(np.asarray(ray_map[v], F32), (1, 2, 0)))
        parts.append(np.asarray(ref_mask[v], F32)[..., None])
        if crop_mask is not None:
            parts.append(np.asarray(crop_mask[v], F32)[..., None])
        outs.append(np.concatenate(parts, -1))
    out = np.stack(outs).astype(F32)
    return (out, np.stack(frags)) if return_fragments else out


def camera_rows(crop_box, intr, extr, ref_extr, target_resolution):
    """The 22 fp64 numbers per view the ray-map kernel takes: load_camera_rays' intrinsics after the crop
    (utils.py:169-173), inv(extr[:3,:3]) (utils.py:183) and ref_extr[:3,:3] (inference_data.py:99)."""
    scale = target_resolution / (crop_box[2] - crop_box[0])
    row = [intr[0, 0] * scale, intr[1, 1] * scale, (intr[0, 2] - crop_box[0]) * scale,
           (intr[1, 2] - crop_box[1]) * scale]
    return np.concatenate([np.asarray(row, np.float64), np.linalg.inv(extr[:3, :3]).reshape(-1),
                           np.asarray(ref_extr[:3, :3], np.float64).reshape(-1)])


def ray_map(crop_box, intr, extr, ref_extr, target_resolution):
    """load_camera_rays(...) then ref_extr[:3,:3] @ rays: fp64 [3,S,S] (the kernel rounds to fp32)."""
    res = target_resolution
    scale = res / (crop_box[2] - crop_box[0])
    fx, fy = intr[0, 0] * scale, intr[1, 1] * scale
    cx, cy = (intr[0, 2] - crop_box[0]) * scale, (intr[1, 2] - crop_box[1]) * scale
    u, v = np.meshgrid(np.arange(res), np.arange(res))
    d = np.stack(((u - cx) / fx, (v - cy) / fy, np.ones_like(u)), axis=0)
    d = d / (np.linalg.norm(d, axis=0, keepdims=True) + 1e-8)
    d = np.linalg.inv(extr[:3, :3]) @ d.reshape(3, -1)
    d = ref_extr[:3, :3] @ d
    return d.reshape(3, res, res)


def verts_to_pytorch3d(verts_2d, crop_box):
    """cap4d/datasets/utils.py:79-89 (returns a new array instead of writing in place)."""
    v = np.array(verts_2d, copy=True)
    v[..., 0] = -((v[..., 0] - crop_box[..., 0]) / (crop_box[..., 2] - crop_box[..., 0]) * 2. - 1.)
    v[..., 1] = -((v[..., 1] - crop_box[..., 1]) / (crop_box[..., 3] - crop_box[..., 1]) * 2. - 1.)
    return v


# ---- synthetic head-like mesh for tests and benches (the FLAME template is a reference asset and is not copied) --
def make_mesh(n_lat=24, n_lon=32, seed=0):
    """UV-sphere-like closed mesh: returns template verts fp32 [Nv,3], faces int32 [F,3], face_mask bool [F]
    (the lower cap is masked out, like the neck faces of the FLAME template)."""
    rng = np.random.default_rng(seed)
    lat = np.linspace(0.08, np.pi - 0.08, n_lat)
    lon = np.linspace(0, 2 * np.pi, n_lon, endpoint=False)
    la, lo = np.meshgrid(lat, lon, indexing="ij")
    r = 1.0 + 0.08 * np.sin(3 * lo) * np.sin(2 * la) + 0.01 * rng.standard_normal(la.shape)
    v = np.stack([r * np.sin(la) * np.cos(lo), r * np.cos(la) * 1.2, r * np.sin(la) * np.sin(lo)], -1).reshape(-1, 3)
    faces = []
    for i in range(n_lat - 1):
        for j in range(n_lon):
            a, b = i * n_lon + j, i * n_lon + (j + 1) % n_lon
            c, d = a + n_lon, b + n_lon
            faces += [(a, c, b), (b, c, d)]
    faces = np.asarray(faces, np.int32)
    face_mask = v[faces][:, :, 1].max(-1) > -0.8
    return v.astype(F32), faces, face_mask


def normalize_props(template_verts):
    """PropRenderer.__init__ (mesh2img.py:360-363): centre, then divide by the overall maximum."""
    p = np.asarray(template_verts, F32)
    p = p - p.mean(axis=-2, keepdims=True, dtype=F32)
    return (p / p.max()).astype(F32)


def make_views(template_verts, n_views, seed=0, scale=0.75):
    """Random poses of the mesh: verts_2d fp32 [n,Nv,3] in pytorch3d NDC with depth z > 0, offsets_3d [n,Nv,3]."""
    rng = np.random.default_rng(seed)
    out, offs = [], []
    for _ in range(n_views):
        a, b = rng.uniform(-0.9, 0.9), rng.uniform(-0.4, 0.4)
        Ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
        Rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
        p = template_verts @ (Ry @ Rx).T
        s = scale * rng.uniform(0.8, 1.25) / np.abs(template_verts).max()
        xy = p[:, :2] * s + rng.uniform(-0.15, 0.15, size=2)
        z = p[:, 2] + 4.0
        out.append(np.concatenate([xy, z[:, None]], -1))
        offs.append(0.0104 * rng.standard_normal(template_verts.shape) * rng.uniform(0.2, 2.0))
    return np.asarray(out, F32), np.asarray(offs, F32)
