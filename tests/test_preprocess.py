"""Frame preprocessing: oracle restatement vs the reference's own stack (Pillow + torchvision) and vs the
committed golden vectors, bit-exact; the CUDA kernels vs the oracle, bit-exact (-m gpu)."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as PO

INTERNVIDEO = ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
OPENAI = ((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _frames(h, w, n=2, seed=0):
    rng = np.random.default_rng(seed + h * 1000 + w)
    fr = rng.integers(0, 256, (n, 3, h, w), dtype=np.uint8)
    fr[-1] = (np.linspace(0, 255, w)[None, None, :] * np.ones((3, h, 1))).astype(np.uint8)
    return fr


def test_oracle_matches_golden_vectors(gold_dir):
    z = np.load(os.path.join(gold_dir, "preprocess_golden.npz"))
    names = sorted(k[:-3] for k in z.files if k.endswith("_in"))
    assert len(names) == 5
    for n in names:
        cfg = z[n + "_cfg"]
        got = PO.frame_transform(z[n + "_in"], int(cfg[0]), tuple(cfg[1:4]), tuple(cfg[4:7]))
        assert np.array_equal(_bits(got), _bits(z[n + "_out"])), n


@pytest.mark.parametrize("h,w,size,ms", [(336, 336, 224, INTERNVIDEO), (336, 336, 336, OPENAI), (360, 640, 224, INTERNVIDEO),
                                         (360, 640, 336, OPENAI), (241, 317, 224, INTERNVIDEO), (100, 150, 224, INTERNVIDEO)])
def test_oracle_matches_pillow_torchvision(h, w, size, ms):
    """The reference itself, run here: the Compose of frame_transform (mm_utils/utils.py:174-183) on PIL images."""
    pytest.importorskip("PIL")
    tv = pytest.importorskip("torchvision.transforms")
    tf = tv.Compose([tv.ToPILImage(), tv.Resize(size, interpolation=tv.InterpolationMode.BICUBIC), tv.CenterCrop(size),
                     lambda im: im.convert("RGB"), tv.ToTensor(), tv.Normalize(mean=ms[0], std=ms[1])])
    fr = _frames(h, w)
    ref = np.stack([tf(torch.from_numpy(f)).numpy() for f in fr])
    assert np.array_equal(_bits(PO.frame_transform(fr, size, *ms)), _bits(ref))


def test_size_rules_match_torchvision():
    from gvl import preprocess
    for h, w, s in [(360, 640, 224), (640, 360, 224), (241, 317, 224), (500, 375, 336), (336, 336, 336), (225, 399, 224)]:
        assert preprocess.resized_size(h, w, s) == PO.resized_size(h, w, s)
        nh, nw = PO.resized_size(h, w, s)
        assert preprocess.center_crop_offsets(nh, nw, s) == PO.center_crop_offsets(nh, nw, s)
    assert PO.center_crop_offsets(224, 399, 224) == (0, 88) and PO.center_crop_offsets(224, 397, 224) == (0, 86)   # half-even


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,size,ms", [(336, 336, 224, INTERNVIDEO), (336, 336, 336, OPENAI), (360, 640, 224, INTERNVIDEO),
                                         (360, 640, 336, OPENAI), (640, 360, 224, INTERNVIDEO), (241, 317, 224, INTERNVIDEO),
                                         (100, 150, 224, INTERNVIDEO), (720, 1280, 336, OPENAI)])
def test_cuda_frame_transform_bit_exact(h, w, size, ms):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gvl import preprocess
    fr = _frames(h, w, n=3)
    out = preprocess.frame_transform(size, mean=ms[0], std=ms[1])(torch.from_numpy(fr).cuda())
    ref = PO.frame_transform(fr, size, *ms)
    assert out.shape == (3, 3, size, size)
    assert np.array_equal(_bits(out.cpu().numpy()), _bits(ref))


@pytest.mark.gpu
def test_cuda_golden_vectors_and_create_pixel_inputs(gold_dir):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gvl import preprocess
    z = np.load(os.path.join(gold_dir, "preprocess_golden.npz"))
    for n in sorted(k[:-3] for k in z.files if k.endswith("_in")):
        cfg = z[n + "_cfg"]
        out = preprocess.frame_transform(int(cfg[0]), mean=tuple(cfg[1:4]), std=tuple(cfg[4:7]))(torch.from_numpy(z[n + "_in"]).cuda())
        assert np.array_equal(_bits(out.cpu().numpy()), _bits(z[n + "_out"])), n
    # the video half of create_inputs (inference.py:69-88): 96 raw frames -> both pixel tensors, key frames 4, 12, ..., 92
    raw = torch.randint(0, 256, (96, 3, 336, 336), dtype=torch.uint8, generator=torch.Generator().manual_seed(1234))
    d = preprocess.create_pixel_inputs(raw.cuda(), 96, 12)
    assert d["temporal_pixel_values"].shape == (1, 96, 3, 224, 224) and d["spatial_pixel_values"].shape == (1, 12, 3, 336, 336)
    key = [4 + 8 * i for i in range(12)]
    ref_sp = PO.frame_transform(raw[key].numpy(), 336, *OPENAI)
    assert np.array_equal(_bits(d["spatial_pixel_values"][0].cpu().numpy()), _bits(ref_sp))
    ref_tp = PO.frame_transform(raw[:8].numpy(), 224, *INTERNVIDEO)
    assert np.array_equal(_bits(d["temporal_pixel_values"][0, :8].cpu().numpy()), _bits(ref_tp))
    with pytest.raises(ValueError):
        preprocess.frame_transform(224)(raw[:2])          # CPU tensor: no CPU path
