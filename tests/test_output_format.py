"""Output hand-off: cap4d_b200/output.py against the files the unmodified reference wrote
(cap4d/inference/utils.py:117-137; fixture from oracle/make_golden_output.py).  Host-side only."""
import os

import numpy as np
import pytest
import torch

from cap4d_b200 import output as OUT

GOLD = os.path.join(os.path.dirname(__file__), "golden", "output_files.npz")


@pytest.fixture(scope="module")
def gold():
    g = np.load(GOLD)
    return {k: g[k] for k in g.files}


class _StubDecoder:
    """Stands in for B200VAEDecoder (which needs a GPU): 'latent' i decodes to fixture image i."""

    def __init__(self, x):
        self.x = torch.from_numpy(x)
        self.calls = []

    def decode_to_uint8_bgr(self, z, batch=8):
        idx = z.flatten(1)[:, 0].long()
        self.calls.append(len(idx))
        return torch.from_numpy(OUT.to_uint8_bgr(self.x[idx]))


def test_quantisation_matches_reference_pixels(gold):
    assert np.array_equal(OUT.to_uint8_bgr(torch.from_numpy(gold["x_samples"])), gold["pixels_bgr"])
    # the clip and the truncation are both exercised by the fixture
    assert gold["pixels_bgr"].min() == 0 and gold["pixels_bgr"].max() == 255


def test_convert_and_save_latent_images_writes_the_reference_files(gold, tmp_path):
    n = gold["x_samples"].shape[0]
    dec = _StubDecoder(gold["x_samples"])
    latents = torch.arange(n, dtype=torch.float32).view(n, 1, 1, 1)
    assert OUT.convert_and_save_latent_images(latents, dec, "cuda:0", tmp_path, batch=2, writers=3) == n
    assert dec.calls == [2, 2, 1]
    files = sorted(os.listdir(tmp_path / "images"))
    assert files == list(gold["file_names"])
    assert np.array_equal(OUT.read_output_images(tmp_path)[..., ::-1], gold["pixels_bgr"])
    if OUT.cv2 is not None and OUT.cv2.__version__ == str(gold["cv2_version"]):
        for i, f in enumerate(files):  # same encoder, same settings: identical bytes
            assert open(tmp_path / "images" / f, "rb").read() == gold[f"png_{i}"].tobytes()
    # an MMLDM carrying the decoder as `b200_vae` is accepted like the reference's `model` argument
    holder = type("M", (), {"b200_vae": dec})()
    (tmp_path / "b").mkdir()
    OUT.convert_and_save_latent_images(latents[:1], holder, "cuda:0", tmp_path / "b")
    assert os.listdir(tmp_path / "b" / "images") == ["00000.png"]
    with pytest.raises(RuntimeError, match="no CPU path"):
        OUT.convert_and_save_latent_images(latents, object(), "cpu", tmp_path)


def test_builtin_png_encoder_round_trips(gold, tmp_path):
    img = gold["pixels_bgr"][0]
    data = OUT.encode_png_bgr(img)
    assert np.array_equal(OUT._decode_png_rgb(data)[..., ::-1], img)
    if OUT.cv2 is not None:
        p = tmp_path / "x.png"
        p.write_bytes(data)
        assert np.array_equal(OUT.cv2.imread(str(p)), img)
    with pytest.raises(ValueError):
        OUT.encode_png_bgr(img.astype(np.float32))


def test_flame_params_and_directory_layout(gold, tmp_path):
    ref_dir, gen_dir = OUT.make_output_dirs(tmp_path / "out")
    assert ref_dir.name == "reference_images" and gen_dir.name == "generated_images" and gen_dir.is_dir()
    names = list(gold["flame_file_names"])
    keys = sorted({k.split("_", 2)[2] for k in gold if k.startswith("flame_0_")})
    flame = [{k: gold[f"flame_{i}_{k}"] for k in keys} for i in range(len(names))]
    OUT.save_flame_params(flame, gen_dir)
    assert sorted(os.listdir(gen_dir / "flame")) == names
    for i, f in enumerate(names):
        got = np.load(gen_dir / "flame" / f)
        assert sorted(got.files) == keys
        for k in keys:
            assert np.array_equal(got[k], gold[f"flame_{i}_{k}"]) and got[k].dtype == gold[f"flame_{i}_{k}"].dtype


def test_save_visualization_layout(tmp_path):
    if OUT.cv2 is None:
        pytest.skip("cv2 not importable")
    vis = {"ray_map": [torch.zeros(1, 16, 16, 3), torch.ones(1, 16, 16, 3)]}
    OUT.save_visualization(vis, tmp_path)
    assert sorted(os.listdir(tmp_path / "condition_vis" / "ray_map")) == ["00000.jpg", "00001.jpg"]
    a = OUT.cv2.imread(str(tmp_path / "condition_vis" / "ray_map" / "00000.jpg"))
    assert a.shape == (16, 16, 3) and abs(int(a.mean()) - 127) <= 1


def test_ranks_write_disjoint_blocks_with_global_names(gold, tmp_path):
    """One process per GPU: every rank writes its block of the views; together they produce the reference's files."""
    n = gold["x_samples"].shape[0]
    latents = torch.arange(n, dtype=torch.float32).view(n, 1, 1, 1)
    counts = [OUT.convert_and_save_latent_images(latents, _StubDecoder(gold["x_samples"]), "cuda:0", tmp_path, batch=2,
                                                 rank=r, world=3) for r in range(3)]
    assert counts == [2, 2, 1] and sum(counts) == n
    assert sorted(os.listdir(tmp_path / "images")) == list(gold["file_names"])
    assert np.array_equal(OUT.read_output_images(tmp_path)[..., ::-1], gold["pixels_bgr"])
    (tmp_path / "solo").mkdir()
    assert OUT.convert_and_save_latent_images(latents, _StubDecoder(gold["x_samples"]), "cuda:0", tmp_path / "solo",
                                              rank=7, world=8) == 0  # more ranks than views: nothing to do
    with pytest.raises(ValueError):
        OUT.convert_and_save_latent_images(latents, _StubDecoder(gold["x_samples"]), "cuda:0", tmp_path, rank=3, world=3)
