"""Conditioning maps on the GPU, through the C ABI (cap4d_b200_cond_pos_enc / _cond_ray_map),
against the fixtures of the unmodified reference `CAP4DConditioning` and against the oracle on the same seeded inputs.

Bars: `pix_to_face` (integer) bit-exact against the oracle; `pos_enc` within 2e-6 + 2e-6*|ref| (fp32 sin/cos of the
device vs the host libm differ in the last bit; the reference's own CPU and CUDA paths differ by the same amount);
pass-through channels (ray map, reference mask, crop mask) bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import cond_oracle as CO

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _close(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b) <= 2e-6 + 2e-6 * np.abs(b)


def _module(faces, props, fmask, **kw):
    from cap4d_b200 import B200CAP4DConditioning

    return B200CAP4DConditioning(torch.as_tensor(faces), torch.as_tensor(props), torch.as_tensor(fmask), **kw)


def _render(cond, dev, verts, offs, ref, ray, crop, p2f=False):
    t = lambda x: None if x is None else torch.as_tensor(x).to(dev)  # noqa: E731
    out = cond.render_pos_enc(t(verts), t(offs), t(ref), t(ray), t(crop), return_pix_to_face=p2f)
    torch.cuda.synchronize()
    return (out[0].cpu().numpy(), out[1].cpu().numpy()) if p2f else out.cpu().numpy()


@pytest.mark.parametrize("name", ["cond_sr2_s32", "cond_sr1_s24_nocrop"])
def test_matches_reference_fixture(cuda_device, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    use_crop = bool(g["use_crop"])
    S, sr = int(g["image_size"]), int(g["super_resolution"])
    cond = _module(g["faces"], g["props"], g["face_mask"], image_size=S, super_resolution=sr, use_crop_mask=use_crop,
                   std_expr_deformation=float(g["std_expr_deformation"]))
    # through the reference's own call: forward(batch, unconditional=False) with [B, T, ...] tensors
    n = g["verts_2d"].shape[0]
    d = lambda x: torch.as_tensor(x).to(cuda_device)[None]  # noqa: E731
    batch = {"verts_2d": d(g["verts_2d"]), "offsets_3d": d(g["offsets_3d"]), "reference_mask": d(g["ref_mask"]),
             "ray_map": d(g["ray_map"]), "out_crop_mask": d(g["crop_mask"])}
    out = cond(batch, unconditional=False)
    pe = out["pos_enc"].cpu().numpy()[0]
    assert pe.shape == g["pos_enc"].shape == (n, S, S, 50 if use_crop else 49)
    assert out["z_input"] is None and np.array_equal(out["ref_mask"].cpu().numpy()[0], g["ref_mask_out"])
    ok = _close(pe, g["pos_enc"])
    assert ok.all(), f"{(~ok).sum()} of {ok.size} differ, max abs {np.abs(pe - g['pos_enc']).max():.3e}"
    assert np.array_equal(pe[..., 45:], g["pos_enc"][..., 45:])  # ray map / masks pass through untouched


@pytest.mark.parametrize("S,sr,n_lat,n_lon,n", [(64, 2, 72, 72, 3), (16, 4, 10, 12, 3), (20, 1, 10, 12, 2),
                                                (12, 2, 6, 8, 5)])
def test_matches_oracle_pix_to_face_bit_exact(cuda_device, S, sr, n_lat, n_lon, n):
    """Production geometry first: 5184 vertices / 10224 faces (FLAME template: 5223 / 10316), 64^2 at 2x."""
    tv, faces, fmask = CO.make_mesh(n_lat, n_lon, seed=S)
    props = CO.normalize_props(tv)
    verts, offs = CO.make_views(tv, n, seed=S + 1)
    rng = np.random.default_rng(S)
    ray = rng.standard_normal((n, 3, S, S)).astype(np.float32)
    ref = (rng.uniform(size=(n, S, S)) > 0.5).astype(np.float32)
    crop = rng.uniform(size=(n, S, S)).astype(np.float32)
    cond = _module(faces, props, fmask, image_size=S, super_resolution=sr, use_crop_mask=True)
    pe, p2f = _render(cond, cuda_device, verts, offs, ref, ray, crop, p2f=True)
    want, want_p2f = CO.cond_pos_enc(verts, offs, faces, props, fmask, ray, ref, crop, S, sr, 42, 1.0, 0.0104,
                                     return_fragments=True)
    assert (want_p2f >= 0).mean() > 0.15
    assert np.array_equal(p2f, want_p2f), f"{(p2f != want_p2f).sum()} pix_to_face entries differ"
    ok = _close(pe, want)
    assert ok.all(), f"{(~ok).sum()} of {ok.size} differ, max abs {np.abs(pe - want).max():.3e}"


def test_optional_channels_and_edge_geometry(cuda_device):
    """No offsets / ray map / crop mask (43 channels); faces behind the camera, zero-area faces, depth ties between
    duplicated faces, a mesh hanging out of the image and an empty batch."""
    tv, faces, fmask = CO.make_mesh(8, 10, seed=4)
    props = CO.normalize_props(tv)
    verts, offs = CO.make_views(tv, 3, seed=9, scale=1.6)  # partly outside [-1, 1]
    verts[1, : tv.shape[0] // 3, 2] = -0.5                # a third of view 1's vertices behind the camera
    faces2 = np.concatenate([faces, faces[::-1], np.asarray([[0, 0, 1], [2, 2, 2]], np.int32)])  # duplicates + degenerate
    fmask2 = np.concatenate([fmask, ~fmask[::-1], [True, True]])
    S = 24
    ref = np.ones((3, S, S), np.float32)
    cond = _module(faces2, props, fmask2, image_size=S, super_resolution=2, use_ray_directions=False,
                   use_expr_deformation=False, use_crop_mask=False)
    pe, p2f = _render(cond, cuda_device, verts, None, ref, None, None, p2f=True)
    want, want_p2f = CO.cond_pos_enc(verts, None, faces2, props, fmask2, None, ref, None, S, 2, 42, 1.0, 0.0104,
                                     return_fragments=True)
    assert pe.shape == (3, S, S, 43)
    assert np.array_equal(p2f, want_p2f) and p2f.max() < faces.shape[0]  # ties go to the first copy of a face
    assert _close(pe, want).all()
    empty = _render(cond, cuda_device, verts[:0], None, ref[:0], None, None)
    assert empty.shape == (0, S, S, 43)


def test_full_workload_properties(cuda_device):
    """840 views at the production geometry (BASELINE config 4), checked through size-independent properties:
    views are independent (a batch equals its views rendered alone), the result does not depend on the order of the
    faces, pass-through channels are exact, background pixels are exactly zero and the positional channels obey
    sin^2 + cos^2 = coverage on pixels whose 2x2 samples are all covered."""
    n, S = 840, 64
    tv, faces, fmask = CO.make_mesh(72, 72, seed=0)
    fmask[:] = True
    props = CO.normalize_props(tv)
    verts, offs = CO.make_views(tv, 12, seed=1)
    reps = n // 12
    verts, offs = np.tile(verts, (reps, 1, 1)), np.tile(offs, (reps, 1, 1))
    verts[:, :, :2] += np.linspace(-0.1, 0.1, n, dtype=np.float32)[:, None, None]
    rng = np.random.default_rng(3)
    ray = rng.standard_normal((n, 3, S, S)).astype(np.float32)
    ref = np.zeros((n, S, S), np.float32)
    crop = np.ones((n, S, S), np.float32)
    cond = _module(faces, props, fmask, image_size=S, super_resolution=2, use_crop_mask=True)
    pe, p2f = _render(cond, cuda_device, verts, offs, ref, ray, crop, p2f=True)
    assert pe.shape == (n, S, S, 50) and np.isfinite(pe).all()
    for i in (0, 417, 839):
        alone = _render(cond, cuda_device, verts[i:i + 1], offs[i:i + 1], ref[i:i + 1], ray[i:i + 1], crop[i:i + 1])
        assert np.array_equal(alone[0], pe[i])
    perm = np.random.default_rng(0).permutation(faces.shape[0])
    cond_p = _module(faces[perm], props, fmask, image_size=S, super_resolution=2, use_crop_mask=True)
    pe_p = _render(cond_p, cuda_device, verts[:24], offs[:24], ref[:24], ray[:24], crop[:24])
    assert np.array_equal(pe_p, pe[:24])
    assert np.array_equal(pe[..., 45:48], ray.transpose(0, 2, 3, 1)) and np.all(pe[..., 48] == 0) and np.all(pe[..., 49] == 1)
    cov = (p2f >= 0).reshape(n, S, 2, S, 2).mean((2, 4))
    assert 0.2 < cov.mean() < 0.9
    assert np.all(pe[..., :45][cov == 0] == 0)
    full = cov == 1
    sc = pe[..., :42].reshape(n, S, S, 3, 2, 7)
    k0 = sc[..., 0, 0] ** 2 + sc[..., 1, 0] ** 2  # lowest frequency: the 4 samples of a pixel are nearly in phase
    assert np.all(k0[full] < 1.0 + 1e-5) and np.median(k0[full]) > 0.99


def test_ray_map_matches_reference_fixture(cuda_device):
    from cap4d_b200.conditioning import camera_rows, ray_maps

    g = np.load(os.path.join(GOLD, "cond_rays.npz"))
    S = int(g["S"])
    rows = camera_rows(g["crop_boxes"], g["intr"], g["extr"], g["ref_extr"], S)
    out = ray_maps(rows, S, cuda_device).cpu().numpy()
    assert out.shape == (4, 3, S, S) and out.dtype == np.float32
    # fp64 on the device, rounded once to fp32: at most one fp32 ulp from the rounded numpy result
    assert np.abs(out.astype(np.float64) - g["rays"]).max() <= 6e-8
    assert (out != g["rays"].astype(np.float32)).mean() < 1e-3
