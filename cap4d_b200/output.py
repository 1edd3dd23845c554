"""Output hand-off: the on-disk contract `gaussianavatars/train.py` reads
(gaussianavatars/scene/dataset_readers.py:74-131): `<out>/{reference_images,generated_images}/images/%05d.png`
(uint8, channels swapped to BGR for cv2, cap4d/inference/utils.py:125-137) and `.../flame/%05d.npz`
(cap4d/inference/utils.py:117-122), optionally `.../condition_vis/<key>/%05d.jpg` (utils.py:103-114).

Same function names and arguments as cap4d/inference/utils.py; the difference is the data plane: latents are decoded
`batch` at a time by the B200 VAE decoder straight to the uint8 BGR arrays (a quarter of the fp32 bytes cross PCIe)
and the PNG files are encoded by a pool of host threads while the GPU decodes the next batch, instead of one
decode -> .cpu() -> imwrite round trip per view.  Files are written by cv2.imwrite like the reference when cv2 is
importable (identical bytes), else by the built-in PNG encoder below (identical pixels).
"""
from __future__ import annotations

import os
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Mapping, Sequence

import numpy as np
import torch

try:  # the reference's writer; present in the reference environment and in this image
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def encode_png_bgr(img_bgr: np.ndarray) -> bytes:
    """Minimal PNG encoder (8-bit RGB, no interlace, filter 0) for uint8 [H,W,3] BGR arrays."""
    if img_bgr.dtype != np.uint8 or img_bgr.ndim != 3 or img_bgr.shape[2] != 3:
        raise ValueError("expected uint8 [H,W,3]")
    h, w, _ = img_bgr.shape
    rgb = np.ascontiguousarray(img_bgr[..., ::-1])
    raw = np.concatenate([np.zeros((h, 1), np.uint8), rgb.reshape(h, w * 3)], axis=1).tobytes()

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
            + chunk(b"IDAT", zlib.compress(raw, 1)) + chunk(b"IEND", b""))


def write_png_bgr(path, img_bgr: np.ndarray) -> None:
    """cv2.imwrite(path, img) of utils.py:136; asserts success like utils.py:137."""
    if cv2 is not None:
        success = cv2.imwrite(str(path), img_bgr)
        assert success, f"failed to save image to {path}"
        return
    with open(path, "wb") as fh:
        fh.write(encode_png_bgr(img_bgr))


def to_uint8_bgr(x_samples: torch.Tensor) -> np.ndarray:
    """utils.py:133-136 on the host for fp32 images [N,3,H,W] in about [-1,1] -> uint8 [N,H,W,3] BGR
    (the device kernel behind B200VAEDecoder.decode_to_uint8_bgr computes the same bytes)."""
    img = ((x_samples + 1.) / 2.).clip(0., 1.)
    img = img.permute(0, 2, 3, 1).cpu().numpy() * 255.
    return np.ascontiguousarray(img[..., [2, 1, 0]].astype(np.uint8))


def convert_and_save_latent_images(latents: torch.Tensor, model, device, output_dir, batch: int = 8,
                                   writers: int = 8, start_index: int = 0, rank: int = 0, world: int = 1) -> int:
    """convert_and_save_latent_images(latents, model, device, output_dir) of cap4d/inference/utils.py:125-137.

    `model` is a B200VAEDecoder, or an MMLDM whose first stage has been installed with `install_vae` / carries a
    `b200_vae` attribute; `device` is accepted for signature compatibility (the decoder is bound to its GPU).
    With one process per GPU every rank holds all latents after the sampler's last all-gather: rank r of `world`
    decodes and writes the contiguous block r of the views (file names keep the global index; no collective).
    Returns the number of files this rank wrote."""
    vae = getattr(model, "b200_vae", model)
    if not hasattr(vae, "decode_to_uint8_bgr"):
        raise RuntimeError("cap4d_b200: convert_and_save_latent_images needs a B200VAEDecoder (there is no CPU path)")
    out_img_dir = Path(output_dir) / "images"
    out_img_dir.mkdir(exist_ok=True)
    if not (0 <= rank < world):
        raise ValueError("rank must be in [0, world)")
    per = (latents.shape[0] + world - 1) // world
    lo, hi = min(rank * per, latents.shape[0]), min((rank + 1) * per, latents.shape[0])
    pending = []
    with ThreadPoolExecutor(max_workers=max(1, writers)) as pool:
        for i in range(lo, hi, batch):
            imgs = vae.decode_to_uint8_bgr(latents[i:min(i + batch, hi)], batch=batch).numpy()
            for j in range(imgs.shape[0]):
                pending.append(pool.submit(write_png_bgr, out_img_dir / f"{start_index + i + j:05d}.png", imgs[j]))
        for f in pending:
            f.result()
    return hi - lo


def save_flame_params(flame_params: Sequence[Mapping[str, np.ndarray]], output_dir) -> None:
    """cap4d/inference/utils.py:117-122."""
    out_flame_dir = Path(output_dir) / "flame"
    out_flame_dir.mkdir(exist_ok=True)
    for frame_id, flame_item in enumerate(flame_params):
        np.savez(out_flame_dir / f"{frame_id:05d}.npz", **flame_item)


def save_visualization(vis_frames: Mapping[str, Sequence[torch.Tensor]], output_dir) -> None:
    """cap4d/inference/utils.py:103-114: condition_vis/<key>/%05d.jpg of the first view of every frame."""
    if cv2 is None:
        raise RuntimeError("save_visualization writes JPEG files through cv2, which is not importable")
    condition_base_dir = Path(output_dir) / "condition_vis"
    condition_base_dir.mkdir(exist_ok=True)
    for key in vis_frames:
        out_dir = condition_base_dir / f"{key}"
        out_dir.mkdir(exist_ok=True)
        for frame_id, vis_img in enumerate(vis_frames[key]):
            vis_img = vis_img[0]
            cv2.imwrite(str(out_dir / f"{frame_id:05d}.jpg"),
                        (((vis_img[..., [2, 1, 0]].cpu().numpy() + 1.) / 2.) * 255).astype(np.uint8))


def make_output_dirs(output_path):
    """generate_images.py:31-36: returns (reference_images, generated_images) directories."""
    output_path = Path(output_path)
    output_path.mkdir(exist_ok=True, parents=True)
    ref, gen = output_path / "reference_images", output_path / "generated_images"
    ref.mkdir(exist_ok=True)
    gen.mkdir(exist_ok=True)
    return ref, gen


def read_output_images(output_dir) -> np.ndarray:
    """What the downstream reader sees (gaussianavatars/scene/dataset_readers.py loads images/%05d.png in file
    order): uint8 [N,H,W,3] in RGB."""
    files = sorted(os.listdir(Path(output_dir) / "images"))
    out = []
    for f in files:
        p = str(Path(output_dir) / "images" / f)
        if cv2 is not None:
            out.append(cv2.imread(p)[..., ::-1])
        else:  # pragma: no cover
            out.append(_decode_png_rgb(open(p, "rb").read()))
    return np.stack(out)


def _decode_png_rgb(data: bytes) -> np.ndarray:
    """Decoder for files written by encode_png_bgr (filter 0 only)."""
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(data):
        ln, tag = struct.unpack(">I", data[pos:pos + 4])[0], data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + ln]
        if tag == b"IHDR":
            w, h = struct.unpack(">II", body[:8])
        elif tag == b"IDAT":
            idat += body
        pos += 12 + ln
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 3 * w)
    assert not raw[:, 0].any()
    return raw[:, 1:].reshape(h, w, 3).copy()
