"""Generate tests/golden/output_files.npz: the files the UNMODIFIED reference writes
(cap4d/inference/utils.py:117-137 `save_flame_params`, `convert_and_save_latent_images`) for seeded images, with a
stand-in model whose `decode_first_stage` returns them.  TEST INFRASTRUCTURE ONLY; needs /root/reference.

    python oracle/make_golden_output.py
"""
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import as RI  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    RI.install_stubs()
    sys.modules["omegaconf"].OmegaConf = object  # imported at the top of cap4d/inference/utils.py, unused here
    if "decord" not in sys.modules:
        dec = types.ModuleType("decord")
        dec.VideoReader = object
        sys.modules["decord"] = dec
    sys.path.insert(0, RI.REFERENCE_ROOT)
    import cv2
    import cap4d.inference.utils as U

    g = torch.Generator().manual_seed(0)
    n, H, W = 5, 24, 40
    yy, xx = torch.meshgrid(torch.linspace(-1.3, 1.3, H), torch.linspace(-1.3, 1.3, W), indexing="ij")
    x = torch.stack([torch.stack([xx * (i + 1) / n, yy, xx * yy]) for i in range(n)]) + 0.05 * torch.randn(n, 3, H, W, generator=g)

    class Model:  # decode_first_stage(latents[None, [i]])[0, 0] -> image i
        def decode_first_stage(self, z):
            return x[int(z.flatten()[0])][None, None]

    latents = torch.arange(n, dtype=torch.float32).view(n, 1, 1, 1)
    flame = [{"fx": np.full((1, 1), 1000.0 + i), "extr": np.eye(4)[None] * (i + 1), "expr": np.arange(6, dtype=np.float32)[None] * i}
             for i in range(3)]
    with tempfile.TemporaryDirectory() as d:
        U.convert_and_save_latent_images(latents, Model(), "cpu", Path(d))
        U.save_flame_params(flame, Path(d))
        files = sorted(os.listdir(os.path.join(d, "images")))
        png = [np.frombuffer(open(os.path.join(d, "images", f), "rb").read(), np.uint8) for f in files]
        pixels = np.stack([cv2.imread(os.path.join(d, "images", f)) for f in files])
        flame_files = sorted(os.listdir(os.path.join(d, "flame")))
        fl = [dict(np.load(os.path.join(d, "flame", f))) for f in flame_files]
    save = {"x_samples": x.numpy(), "file_names": np.array(files), "pixels_bgr": pixels, "cv2_version": cv2.__version__,
            "flame_file_names": np.array(flame_files)}
    for i, p in enumerate(png):
        save[f"png_{i}"] = p
    for i, f in enumerate(fl):
        for k, v in f.items():
            save[f"flame_{i}_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "output_files.npz"), **save)
    print("output_files:", files, flame_files, "png bytes", [len(p) for p in png])


if __name__ == "__main__":
    main()
