"""ORACLE (test infrastructure only -- never imported by the product path): CPU restatement of the reference's frame
preprocessing, `frame_transform` (mm_utils/utils.py:153-183) as used by `create_inputs` (inference.py:69-88):

    ToPILImage -> Resize(size, BICUBIC)  [shortest edge, aspect preserving] -> CenterCrop(size) -> convert('RGB')
    -> ToTensor (/255, float32) -> Normalize(mean, std)

The resize is Pillow's `ImagingResample` for 8-bit images (third-party, Pillow; torchvision.transforms.Resize on a PIL image calls
`Image.resize(size, BICUBIC)`), restated here from its published algorithm: separable two-pass convolution (horizontal, then
vertical) with an 8-bit intermediate image, bicubic kernel a = -0.5 with support 2 * max(scale, 1), coefficients normalised in
double precision and converted to 22-bit fixed point, accumulators start at 1 << 21, result = clamp(acc >> 22, 0, 255).
Pinned bit-exact against Pillow / torchvision themselves in tests/test_preprocess.py (they are installed in this image).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size, out_size):
    """Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc -> (ksize, bounds[out,2] (xmin, n), kk[out,ksize] int32)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass(img, out_size, axis):
    """One resampling pass over `axis` of a uint8 array [..., H, W]."""
    in_size = img.shape[axis]
    _, bounds, kk = precompute_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, -1).astype(np.int64)
    out = np.empty(src.shape[:-1] + (out_size,), dtype=np.uint8)
    for xx in range(out_size):
        xmin, n = bounds[xx]
        acc = (src[..., xmin:xmin + n] * kk[xx, :n].astype(np.int64)).sum(-1) + (1 << (PRECISION_BITS - 1))
        out[..., xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, -1, axis)


def resize_bicubic_u8(img, out_h, out_w):
    """img: uint8 [..., H, W]. Pillow order: horizontal pass first (skipped when the width is unchanged), then vertical."""
    if img.shape[-1] != out_w:
        img = _pass(img, out_w, -1)
    if img.shape[-2] != out_h:
        img = _pass(img, out_h, -2)
    return img


def resized_size(h, w, size):
    """torchvision _compute_resized_output_size for an int size: shortest edge -> size, long edge = int(size * long / short)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)     # (new_h, new_w)


def center_crop_offsets(h, w, size):
    """torchvision center_crop: int(round((h - size) / 2.0)) with Python's round-half-even."""
    return int(round((h - size) / 2.0)), int(round((w - size) / 2.0))


def frame_transform(frames, size, mean, std):
    """frames: uint8 [N, 3, H, W] (what read_frames_decord hands to the processor). Returns float32 [N, 3, size, size]."""
    n, c, h, w = frames.shape
    nh, nw = resized_size(h, w, size)
    r = resize_bicubic_u8(frames, nh, nw)
    top, left = center_crop_offsets(nh, nw, size)
    r = r[:, :, top:top + size, left:left + size]
    x = r.astype(np.float32) / np.float32(255)
    m = np.asarray(mean, dtype=np.float32).reshape(1, 3, 1, 1)
    s = np.asarray(std, dtype=np.float32).reshape(1, 3, 1, 1)
    return ((x - m) / s).astype(np.float32)
