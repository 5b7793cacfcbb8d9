"""Generates tests/golden/preprocess_golden.npz from the reference's own preprocessing stack (torchvision transforms on PIL
images, exactly the Compose built by frame_transform, mm_utils/utils.py:153-183). Run in a container where Pillow and
torchvision are installed:  python oracle/make_golden_preprocess.py"""
import os

import numpy as np
import torch
from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToPILImage, ToTensor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INTERNVIDEO = ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
OPENAI = ((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))


def reference_transform(size, mean, std):
    return Compose([ToPILImage(), Resize(size, interpolation=InterpolationMode.BICUBIC), CenterCrop(size),
                    lambda im: im.convert("RGB"), ToTensor(), Normalize(mean=mean, std=std)])


def main():
    rng = np.random.default_rng(20240229)
    out = {}
    for name, (h, w, size, ms) in {"down_landscape": (37, 53, 24, INTERNVIDEO), "down_portrait": (61, 40, 32, OPENAI),
                                    "up": (20, 27, 32, INTERNVIDEO), "same": (32, 32, 32, OPENAI),
                                    "odd_crop": (45, 64, 32, INTERNVIDEO)}.items():
        fr = rng.integers(0, 256, (2, 3, h, w), dtype=np.uint8)
        fr[1] = (np.linspace(0, 255, w)[None, None, :] * np.ones((3, h, 1))).astype(np.uint8)
        tf = reference_transform(size, *ms)
        ref = np.stack([tf(torch.from_numpy(f)).numpy() for f in fr])
        out[name + "_in"] = fr
        out[name + "_out"] = ref
        out[name + "_cfg"] = np.array([size] + list(ms[0]) + list(ms[1]), dtype=np.float64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "preprocess_golden.npz"), **out)
    print("wrote", len(out) // 3, "cases")


if __name__ == "__main__":
    main()
