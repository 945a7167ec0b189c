"""Deterministic synthetic JPEG inputs of the BASELINE.json shapes.

There is no network and no dataset in the image, so every bitstream the tests
and bench.py decode is produced here: a seeded procedural picture (smooth
gradients + band-limited noise — white noise would inflate the bitstream far
beyond photographic bit rates) encoded at quality 90 with
the libjpeg-turbo encoders bundled in Pillow / OpenCV, baseline, standard
Huffman tables, with or without restart intervals.
"""
from __future__ import annotations

import io
import numpy as np

CSS_NAMES = ("444", "440", "422", "420", "411", "400")


def synth_image(width: int, height: int, seed: int = 0) -> np.ndarray:
    """Seeded RGB uint8 picture, shape (height, width, 3)."""
    rng = np.random.default_rng(seed)
    u = (np.arange(width, dtype=np.float32) / max(width, 1))[None, :]
    v = (np.arange(height, dtype=np.float32) / max(height, 1))[:, None]
    fx, fy = rng.uniform(0.5, 3.0, 2) * 2 * np.pi
    ph = rng.uniform(0, 2 * np.pi, 6)
    img = np.empty((height, width, 3), dtype=np.float32)
    img[..., 0] = 128 + 90 * np.sin(fx * u + ph[0]) * np.cos(fy * v + ph[1])
    img[..., 1] = 128 + 90 * np.cos(0.5 * fx * u + ph[3]) * np.sin(fy * v + ph[2])
    img[..., 2] = 128 + 90 * np.sin(1.5 * fx * u + ph[4]) * np.cos(0.7 * fy * v + ph[5])
    # band-limited noise: low-resolution Gaussian noise, upsampled and box-blurred
    f = 3
    lh, lw = (height + f - 1) // f + 2, (width + f - 1) // f + 2
    lo = rng.normal(0.0, 12.0, (lh, lw, 3)).astype(np.float32)
    up = np.repeat(np.repeat(lo, f, axis=0), f, axis=1)
    up = (up[:-2] + up[1:-1] + up[2:]) * np.float32(1.0 / 3.0)
    up = (up[:, :-2] + up[:, 1:-1] + up[:, 2:]) * np.float32(1.0 / 3.0)
    img += up[:height, :width]
    # a few hard edges so high-frequency coefficients and long codes occur
    ex = (np.arange(width) // 37)[None, :]
    ey = (np.arange(height) // 53)[:, None]
    img += (((ex + ey) % 5) == 0).astype(np.float32)[..., None] * np.float32(30.0)
    np.clip(img, 0, 255, out=img)
    return img.astype(np.uint8)


def encode_jpeg(img: np.ndarray, css: str = "420", quality: int = 90, restart_mcus: int = 0,
                restart_rows: int = 0) -> bytes:
    """Baseline JPEG bytes. `restart_mcus` = DRI in MCUs; `restart_rows` = DRI in MCU rows."""
    assert css in CSS_NAMES
    if css in ("440", "411"):   # Pillow offers 4:4:4 / 4:2:2 / 4:2:0 only: OpenCV's libjpeg-turbo writes these two
        import cv2

        params = [cv2.IMWRITE_JPEG_QUALITY, quality, cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                  cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440 if css == "440" else cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411]
        if restart_rows and not restart_mcus:
            restart_mcus = restart_rows * ((img.shape[1] + (7 if css == "440" else 31)) // (8 if css == "440" else 32))
        if restart_mcus:
            params += [cv2.IMWRITE_JPEG_RST_INTERVAL, int(restart_mcus)]
        ok, buf = cv2.imencode(".jpg", img[..., ::-1], params)
        assert ok
        return buf.tobytes()
    from PIL import Image

    kw = dict(format="JPEG", quality=quality, optimize=False, progressive=False)
    if css == "400":
        r, g, b = img[..., 0].astype(np.float32), img[..., 1].astype(np.float32), img[..., 2].astype(np.float32)
        gray = np.clip(0.299 * r + 0.587 * g + 0.114 * b + 0.5, 0, 255).astype(np.uint8)
        im = Image.fromarray(gray, mode="L")
    else:
        im = Image.fromarray(img, mode="RGB")
        kw["subsampling"] = {"444": 0, "422": 1, "420": 2}[css]
    if restart_mcus:
        kw["restart_marker_blocks"] = int(restart_mcus)
    elif restart_rows:
        kw["restart_marker_rows"] = int(restart_rows)
    bio = io.BytesIO()
    im.save(bio, **kw)
    return bio.getvalue()


def make_jpeg(width: int, height: int, css: str, seed: int = 0, quality: int = 90, restart_mcus: int = 0,
              restart_rows: int = 0) -> bytes:
    return encode_jpeg(synth_image(width, height, seed), css, quality, restart_mcus, restart_rows)


def workload(name: str, n: int | None = None):
    """Named BASELINE.json workloads -> (list of jpeg bytes, output format name)."""
    if name == "c2":  # single 1920x1080 4:2:0 -> RGB
        return [make_jpeg(1920, 1080, "420", seed=2)], "rgb"
    if name == "c3":  # 256 ImageNet-shaped, mixed 444/422/420 -> RGB_PLANAR
        n = n or 256
        return [make_jpeg(500, 375, ("444", "422", "420")[i % 3], seed=100 + i) for i in range(n)], "rgb_planar"
    if name == "c3j":  # jittered sizes, odd values included
        n = n or 256
        rng = np.random.default_rng(7)
        out = []
        for i in range(n):
            w, h = int(rng.integers(300, 641)), int(rng.integers(224, 501))
            out.append(make_jpeg(w, h, ("444", "422", "420")[i % 3], seed=300 + i))
        return out, "rgb_planar"
    if name in ("c4_dri", "c4_nodri"):  # 64 x 3840x2160 4:2:2 -> YUV_PLANAR
        n = n or 64
        rows = 1 if name == "c4_dri" else 0
        base = [make_jpeg(3840, 2160, "422", seed=400 + i, restart_rows=rows) for i in range(min(n, 8))]
        return [base[i % len(base)] for i in range(n)], "yuv_planar"
    if name == "c5_400":
        return [make_jpeg(8192, 8192, "400", seed=5)], "y"
    if name == "c5_440":
        return [make_jpeg(8192, 8192, "440", seed=6)], "native"
    raise ValueError(name)
