"""Conditioning maps, CPU side: the oracle against the fixtures the unmodified reference
`CAP4DConditioning` / `load_camera_rays` produced (oracle/make_golden_cond.py), known-answer tests of the restated
pytorch3d rasterisation rules, and the host logic of cap4d_b200/conditioning.py.  No GPU needed."""
import os

import numpy as np
import pytest
import torch

from oracle import cond_oracle as CO

GOLD = os.path.join(os.path.dirname(__file__), "golden")
FIXTURES = ["cond_sr2_s32", "cond_sr1_s24_nocrop"]


def _load(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: g[k] for k in g.files}


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_matches_reference_fixture(name):
    g = _load(name)
    out = CO.cond_pos_enc(g["verts_2d"], g["offsets_3d"], g["faces"], g["props"], g["face_mask"], g["ray_map"],
                          g["ref_mask"], g["crop_mask"] if bool(g["use_crop"]) else None, int(g["image_size"]),
                          int(g["super_resolution"]), 42, 1.0, float(g["std_expr_deformation"]))
    assert out.shape == g["pos_enc"].shape and out.dtype == np.float32
    assert np.abs(out - g["pos_enc"]).max() <= 1e-6
    # the fixture is not degenerate: a third of the image is covered and every channel group is populated
    assert (np.abs(g["pos_enc"][..., :42]).sum(-1) > 0).mean() > 0.25
    assert np.abs(g["pos_enc"][..., 42:45]).max() > 0.5


def test_ray_map_and_vertex_transform_match_reference_fixture():
    g = _load("cond_rays")
    S = int(g["S"])
    for i in range(g["rays"].shape[0]):
        r = CO.ray_map(g["crop_boxes"][i], g["intr"][i], g["extr"][i], g["ref_extr"], S)
        assert np.abs(r - g["rays"][i]).max() <= 1e-15
        assert np.abs(np.linalg.norm(r, axis=0) - 1.0).max() < 1e-6
        v = CO.verts_to_pytorch3d(g["verts_in"][i], g["crop_boxes"][i])
        assert np.array_equal(v, g["verts_out"][i])


# ---- known answers for the rasterisation rules (pytorch3d 0.7.8 conventions) ---------------------------------
def _tri(*pts):
    return np.asarray(pts, np.float32), np.asarray([[0, 1, 2]], np.int32)


def test_pixel_centres_and_axis_orientation():
    # +X points LEFT and +Y UP: a triangle in the (+x, +y) NDC quadrant lands in the top-left image quadrant
    v, f = _tri((0.05, 0.05, 1), (0.95, 0.05, 1), (0.05, 0.95, 1))
    p2f, zbuf, bary = CO.rasterize(v, f, 8)
    ys, xs = np.nonzero(p2f == 0)
    assert ys.size > 0 and ys.max() < 4 and xs.max() < 4
    # pixel centres: row/col i samples NDC 1 - (2 i + 1) / S
    assert np.allclose(CO.pix_to_ndc(np.arange(7, -1, -1), 8), 1 - (2 * np.arange(8) + 1) / 8)
    # (row 3, col 3) has centre (0.125, 0.125): inside; (row 0, col 0) has centre (0.875, 0.875): outside (x + y > 1)
    assert p2f[3, 3] == 0 and p2f[0, 0] == -1
    assert np.allclose(bary[p2f == 0].sum(-1), 1.0, atol=1e-6)
    assert np.allclose(zbuf[p2f == 0], 1.0, atol=1e-6) and np.all(zbuf[p2f < 0] == -1.0)


def test_edges_are_exclusive_and_barycentrics_interpolate():
    # full-image right triangle; pixel centres exactly on the hypotenuse x + y = 0 are NOT covered (w > 0 strictly)
    v, f = _tri((-1, -1, 2), (1, -1, 4), (-1, 1, 6))
    p2f, zbuf, bary = CO.rasterize(v, f, 4)
    ndc = CO.pix_to_ndc(np.arange(3, -1, -1), 4)
    for r in range(4):
        for c in range(4):
            inside = ndc[c] + ndc[r] < 0
            assert (p2f[r, c] == 0) == bool(inside)
            if inside:  # z is affine in (x, y): 2 + (x + 1) + 2 (y + 1)
                assert abs(zbuf[r, c] - (2 + (ndc[c] + 1) + 2 * (ndc[r] + 1))) < 1e-5


def test_depth_order_ties_and_rejections():
    big = [(-1, -1), (3, -1), (-1, 3)]
    verts = np.asarray([(x, y, 5.0) for x, y in big] + [(x, y, 2.0) for x, y in big] + [(x, y, 2.0) for x, y in big]
                       + [(x, y, -1.0) for x, y in big] + [(0, 0, 1), (0.5, 0.5, 1), (1, 1, 1)], np.float32)
    faces = np.arange(15, dtype=np.int32).reshape(5, 3)
    p2f, zbuf, _ = CO.rasterize(verts, faces, 6)
    # face 0 (z 5) is behind faces 1 and 2 (z 2, identical): the tie goes to the smaller index; face 3 has z < eps
    # (skipped entirely); face 4 has zero area
    assert np.all(p2f == 1) and np.allclose(zbuf, 2.0, atol=1e-6)
    # a face with ONE vertex behind the camera is dropped as a whole (z_invalid in CheckPointOutsideBoundingBox)
    verts2 = verts.copy()
    verts2[3:9, 2] = [2.0, 2.0, 0.0, 2.0, 2.0, 0.0]
    assert np.all(CO.rasterize(verts2, faces, 6)[0] == 0)


def test_rasterize_is_face_order_invariant_up_to_relabelling():
    tv, faces, _ = CO.make_mesh(10, 12, seed=1)
    verts, _ = CO.make_views(tv, 1, seed=2)
    a, za, ba = CO.rasterize(verts[0], faces, 48)
    perm = np.random.default_rng(0).permutation(faces.shape[0])
    b, zb, bb = CO.rasterize(verts[0], faces[perm], 48)
    assert np.array_equal(za, zb) and np.array_equal(ba, bb)
    assert np.array_equal(np.where(b >= 0, perm[np.clip(b, 0, None)], -1), a)
    assert 0.2 < (a >= 0).mean() < 0.9


# ---- host logic ------------------------------------------------------------------------------------------------
def _cond(**kw):
    from cap4d_b200 import B200CAP4DConditioning

    tv, faces, fmask = CO.make_mesh(6, 8, seed=0)
    return B200CAP4DConditioning(torch.from_numpy(faces), torch.from_numpy(CO.normalize_props(tv)),
                                 torch.from_numpy(fmask), **kw), tv


def test_load_template_matches_reference_renderer_buffers(tmp_path):
    """faces / props / face_mask of the fixture came out of the unmodified PropRenderer.__init__ reading the same
    .obj text through (a stub of) pytorch3d's load_obj."""
    from cap4d_b200.conditioning import load_template

    g = _load("cond_sr2_s32")
    tv, faces, _ = CO.make_mesh(24, 32, seed=0)
    obj, hv = tmp_path / "t.obj", tmp_path / "head.txt"
    with open(obj, "w") as fh:
        for v in tv:
            fh.write("v %.9g %.9g %.9g\n" % tuple(v))
        for f in faces:
            fh.write("f %d/%d %d/%d %d/%d\n" % tuple(int(i) + 1 for i in np.repeat(f, 2)))
    np.savetxt(hv, np.nonzero(tv[:, 1] > -0.6)[0], fmt="%d")
    f2, props, fmask = load_template(str(obj), str(hv), n_mouth_verts=40)
    assert np.array_equal(f2.numpy(), g["faces"])
    assert np.array_equal(props.numpy(), g["props"])
    assert np.array_equal(fmask.numpy(), g["face_mask"])
    assert np.abs(CO.normalize_props(tv) - g["props"]).max() < 1e-6  # numpy mean vs torch mean: last-bit differences


def test_unconditional_branch_and_channel_count():
    cond, tv = _cond(image_size=16, use_crop_mask=True)
    assert cond.total_channels == 50
    batch = {"verts_2d": torch.zeros(2, 3, tv.shape[0], 3), "offsets_3d": torch.zeros(2, 3, tv.shape[0], 3),
             "reference_mask": torch.ones(2, 3, 16, 16), "z": torch.ones(2, 3, 4, 16, 16)}
    out = cond(batch, unconditional=True)  # cap4dcond.py:78-88: zeros, z * 0, ref_mask with a channel axis
    assert out["pos_enc"].shape == (2, 3, 16, 16, 50) and float(out["pos_enc"].abs().max()) == 0
    assert float(out["z_input"].abs().max()) == 0 and out["ref_mask"].shape == (2, 3, 1, 16, 16)
    vis = cond.get_vis(out["pos_enc"])
    assert set(vis) == {"pose_map_12", "pose_map_13", "expr_disp", "ray_map", "ref_mask", "crop_mask"}
    assert _cond(image_size=16, use_ray_directions=False, use_expr_deformation=False)[0].total_channels == 43


def test_conditioning_has_no_cpu_path_and_validates_arguments():
    from cap4d_b200 import B200CAP4DConditioning
    from cap4d_b200.conditioning import ray_maps

    cond, tv = _cond(image_size=16)
    batch = {"verts_2d": torch.zeros(1, 1, tv.shape[0], 3), "offsets_3d": torch.zeros(1, 1, tv.shape[0], 3),
             "reference_mask": torch.ones(1, 1, 16, 16), "ray_map": torch.zeros(1, 1, 3, 16, 16)}
    with pytest.raises(RuntimeError, match="no CPU path"):
        cond(batch, unconditional=False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ray_maps(np.zeros((1, 22)), 16, "cpu")
    with pytest.raises(ValueError):
        B200CAP4DConditioning(torch.tensor([[0, 1, 9]]), torch.zeros(3, 3), torch.ones(1), super_resolution=2)
    with pytest.raises(ValueError):
        B200CAP4DConditioning(torch.tensor([[0, 1, 2]]), torch.zeros(3, 3), torch.ones(1), super_resolution=3)
    with pytest.raises(AssertionError):
        B200CAP4DConditioning(torch.tensor([[0, 1, 2]]), torch.zeros(3, 3), torch.ones(1), positional_channels=40)


def test_c_abi_rejects_bad_arguments_without_gpu():
    from cap4d_b200 import _lib

    lib = _lib.load()
    one = (np.zeros(16, np.float32)).ctypes.data
    args = lambda sr, pc, n_faces, ws=one, nb=64: (one, None, one, one, one, None, one, None, one, None, 1, 3,  # noqa: E731
                                                   n_faces, 8, sr, pc, 1.0, 0.0104, ws, nb, None)
    assert lib.cap4d_b200_cond_pos_enc(*args(3, 42, 1)) != 0 and "super_resolution" in _lib.last_error()
    assert lib.cap4d_b200_cond_pos_enc(*args(2, 40, 1)) != 0 and "positional_channels" in _lib.last_error()
    assert lib.cap4d_b200_cond_pos_enc(*args(2, 42, 0)) != 0
    assert lib.cap4d_b200_cond_pos_enc(None, None, one, one, one, None, one, None, one, None, 1, 3, 1, 8, 2, 42, 1.0,
                                       0.0104, one, 64, None) != 0
    assert lib.cap4d_b200_cond_pos_enc(*args(2, 42, 4, nb=8)) != 0 and "workspace" in _lib.last_error()
    import ctypes
    nbytes = ctypes.c_size_t()
    assert lib.cap4d_b200_cond_workspace_bytes(840, 10316, ctypes.byref(nbytes)) == 0 and nbytes.value == 840 * 10316 * 4
    assert lib.cap4d_b200_cond_workspace_bytes(3, 10, ctypes.byref(nbytes)) == 0 and nbytes.value == 3 * 10 * 4
    assert lib.cap4d_b200_cond_ray_map(None, one, 1, 8, None) != 0


def test_camera_rows_match_oracle():
    from cap4d_b200.conditioning import camera_rows

    g = _load("cond_rays")
    rows = camera_rows(g["crop_boxes"], g["intr"], g["extr"], g["ref_extr"], int(g["S"]))
    assert rows.shape == (4, 22) and rows.dtype == np.float64
    for i in range(4):
        assert np.array_equal(rows[i], CO.camera_rows(g["crop_boxes"][i], g["intr"][i], g["extr"][i], g["ref_extr"],
                                                      int(g["S"])))


def test_get_condition_from_dataloader_glue():
    """Host glue of cap4d/inference/utils.py:64-100 with stand-in stages (the real ones need a GPU): key set, the
    'b t ... -> (b t) ...' flattening, z hand-over, flame parameter unpacking."""
    from cap4d_b200.conditioning import concat_frames, get_condition_from_dataloader

    S = 8

    class Cond:
        def __call__(self, hint, unconditional=True):
            B, T = hint["verts_2d"].shape[:2]
            pe = torch.zeros(B, T, S, S, 50) if unconditional else torch.ones(B, T, S, S, 50) * hint["verts_2d"].mean()
            return {"pos_enc": pe, "z_input": hint["z"] * (0. if unconditional else 1.),
                    "ref_mask": hint["reference_mask"][:, :, None]}

        def get_vis(self, enc):
            return {"ray_map": enc[..., 45:48]}

    class Vae:
        def encode_first_stage(self, x):
            assert x.shape[2] == 3  # b t c h w
            return x[:, :, :1, ::8, ::8].repeat(1, 1, 4, 1, 1) + 1.0

    def frame(i, is_ref):
        return {"jpg": torch.full((1, 1, 64, 64, 3), float(i)),
                "hint": {"verts_2d": torch.full((1, 1, 5, 3), float(i)), "offsets_3d": torch.zeros(1, 1, 5, 3),
                         "reference_mask": torch.full((1, 1, S, S), float(is_ref))},
                "flame_params": {"fx": torch.full((1, 1, 1), 100.0 + i), "extr": torch.eye(4)[None, None] * i}}

    out = get_condition_from_dataloader(Cond(), Vae(), [frame(0, True), frame(1, False), frame(2, False)], "cpu",
                                        visualize=True)
    cf, uf = concat_frames(out["cond_frames"]), concat_frames(out["uncond_frames"])
    assert set(cf) == {"pos_enc", "z_input", "ref_mask"}
    assert cf["pos_enc"].shape == (3, S, S, 50) and cf["z_input"].shape == (3, 4, S, S) and cf["ref_mask"].shape == (3, 1, S, S)
    assert [float(cf["pos_enc"][i].mean()) for i in range(3)] == [0.0, 1.0, 2.0]
    assert [float(cf["z_input"][i].mean()) for i in range(3)] == [1.0, 2.0, 3.0] and float(uf["z_input"].abs().max()) == 0
    assert float(uf["pos_enc"].abs().max()) == 0 and torch.equal(uf["ref_mask"], cf["ref_mask"])
    assert len(out["flame_params"]) == 3 and out["flame_params"][2]["fx"].shape == (1, 1)
    assert float(out["flame_params"][1]["fx"][0, 0]) == 101.0
    assert out["cond_vis_frames"]["ray_map"][0].shape == (1, 8 * S, 8 * S, 3)
