"""CPU oracle for the dataset's colour pyramid.  TEST INFRASTRUCTURE ONLY.

The reference builds, per frame, ``("color", f, i) = Resize((H >> i, W >> i), Image.ANTIALIAS)(("color", f, i-1))``
on PIL uint8 images and then ``ToTensor`` (datasets/mono_dataset.py:84-111).  The arithmetic lives in two
third-party libraries that are absent from /root/reference and unpinned by it (no requirements file):

* Pillow's ``ImagingResample`` (src/libImaging/Resample.c; version here: 12.2.0).  ``Image.ANTIALIAS`` is
  ``Image.LANCZOS`` (support 3).  Published algorithm, restated below in numpy:
  - ``precompute_coeffs``: ``scale = in/out``; ``filterscale = max(scale, 1)``; ``support = 3 * filterscale``;
    for output ``xx``: ``center = (xx + 0.5) * scale``, ``xmin = max(int(center - support + 0.5), 0)``,
    ``xmax = min(int(center + support + 0.5), in)``, weights ``lanczos((x + xmin - center + 0.5) / filterscale)``
    normalised to sum 1 (all in double);
  - ``normalize_coeffs_8bpc``: ``k = int(+-0.5 + w * 2**22)``;
  - horizontal pass over the rows the vertical pass needs, then vertical pass, each
    ``clip8((2**21 + sum(u8 * k)) >> 22)`` with a uint8 intermediate image.
* torchvision's ``ToTensor``: HWC uint8 -> CHW float32 ``.div(255)``.

Pinning: ``tests/golden/make_golden_pyramid.py`` runs the reference's own ``MonoDataset.preprocess`` unbound
(which calls the real Pillow and torchvision) and commits inputs + outputs; ``tests/test_pyramid.py`` checks
this restatement against those fixtures bit for bit, and against the installed Pillow on random sizes.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
LANCZOS_SUPPORT = 3.0


def _sinc(x: float) -> float:
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def lanczos(x: float) -> float:
    """Resample.c lanczos_filter: truncated sinc, support 3."""
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3.0)
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for box (0, in_size).
    -> (bounds [out,2] = (xmin, count), coeffs [out, ksize] int32)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = LANCZOS_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One pass of ImagingResampleHorizontal/Vertical_8bpc along ``axis`` of an HWC uint8 array."""
    in_size = img.shape[axis]
    bounds, kk = precompute_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        xmin, cnt = bounds[xx]
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(cnt):
            acc += src[xmin + x] * kk[xx, x]
        out[xx] = _clip8(acc)
    return np.moveaxis(out, 0, axis)


def resize_lanczos(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """``PIL.Image.fromarray(img).resize((out_w, out_h), Image.LANCZOS)`` for an HWC uint8 array:
    horizontal pass first, then vertical (ImagingResample), each skipped when the size is unchanged."""
    out = img
    if out_w != img.shape[1]:
        out = _resample_axis(out, out_w, 1)
    if out_h != img.shape[0]:
        out = _resample_axis(out, out_h, 0)
    return out


def to_tensor(img: np.ndarray) -> np.ndarray:
    """torchvision ToTensor on a uint8 HWC image: CHW float32, value / 255 (fp32 division)."""
    return (np.ascontiguousarray(np.moveaxis(img, 2, 0)).astype(np.float32) / np.float32(255.0)).astype(np.float32)


def pyramid(frames: np.ndarray, n_scales: int):
    """mono_dataset.py:99-111 for the loss-side colours: frames uint8 [N,H,W,3] (scale 0) ->
    list over scales of float32 [N,3,H>>s,W>>s]; scale s is resized from scale s-1."""
    N, H, W, _ = frames.shape
    levels = [frames]
    for s in range(1, n_scales):
        prev = levels[-1]
        levels.append(np.stack([resize_lanczos(prev[n], H // 2 ** s, W // 2 ** s) for n in range(N)], 0))
    return [np.stack([to_tensor(lv[n]) for n in range(N)], 0) for lv in levels], levels
