"""GPU mirror of the reference's frame preprocessing.

`frame_transform(image_size, mean=, std=)` has the reference's signature (mm_utils/utils.py:153-183) and returns a callable
that takes uint8 frames `[N, 3, H, W]` (or one frame `[3, H, W]`) on the GPU and returns float32 `[N, 3, size, size]`,
bit-identical to `ToPILImage -> Resize(BICUBIC) -> CenterCrop -> ToTensor -> Normalize` applied frame by frame on the host.
`create_pixel_inputs` is the video half of `create_inputs` (inference.py:69-88): temporal stream = every sampled frame at
224 with the InternVideo2 statistics, spatial stream = the middle frame of each segment at 336 with the OpenAI CLIP ones.
All arithmetic happens in libgvl.so (csrc/preprocess.cu); this file only reproduces torchvision's integer size rules.
"""
import ctypes

import torch

from . import _lib

OPENAI_DATASET_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_DATASET_STD = (0.26862954, 0.26130258, 0.27577711)
INTERNVIDEO_MEAN = (0.485, 0.456, 0.406)
INTERNVIDEO_STD = (0.229, 0.224, 0.225)


def resized_size(h, w, size):
    """torchvision Resize(int): shortest edge -> size, long edge -> int(size * long / short)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def center_crop_offsets(h, w, size):
    """torchvision CenterCrop: int(round((h - size) / 2.0)) -- Python's round-half-even, kept literally."""
    return int(round((h - size) / 2.0)), int(round((w - size) / 2.0))


class FrameTransform:
    def __init__(self, image_size, mean, std):
        self.size = int(image_size)
        self.mean = (ctypes.c_float * 3)(*[float(m) for m in mean])
        self.std = (ctypes.c_float * 3)(*[float(s) for s in std])

    def __call__(self, frames):
        lib = _lib.load()
        single = frames.dim() == 3
        if single:
            frames = frames[None]
        if not frames.is_cuda:
            raise ValueError("frames must be a CUDA tensor (gvl has no CPU path)")
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[1] != 3:
            raise ValueError("frames must be uint8 [N, 3, H, W], got %s %s" % (frames.dtype, tuple(frames.shape)))
        frames = frames.contiguous()
        n, _, h, w = frames.shape
        nh, nw = resized_size(h, w, self.size)
        if nh < self.size or nw < self.size:
            raise ValueError("resized frame %dx%d is smaller than the crop %d" % (nh, nw, self.size))
        top, left = center_crop_offsets(nh, nw, self.size)
        ws_bytes = lib.gvl_frame_transform_workspace(n, h, w, nh, nw)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=frames.device)
        out = torch.empty((n, 3, self.size, self.size), dtype=torch.float32, device=frames.device)
        rc = lib.gvl_frame_transform(ctypes.c_void_p(frames.data_ptr()), n, h, w, nh, nw, top, left, self.size, self.mean, self.std,
                                     ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(ws.data_ptr()), ws_bytes,
                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "gvl_frame_transform")
        return out[0] if single else out


def frame_transform(image_size, rescale_factor=1.0, mean=None, std=None):
    """Same signature as the reference's frame_transform (rescale_factor is unused there as well)."""
    mean = mean or OPENAI_DATASET_MEAN
    if not isinstance(mean, (list, tuple)):
        mean = (mean,) * 3
    std = std or OPENAI_DATASET_STD
    if not isinstance(std, (list, tuple)):
        std = (std,) * 3
    if isinstance(image_size, (list, tuple)) and image_size[0] == image_size[1]:
        image_size = image_size[0]
    return FrameTransform(image_size, mean, std)


def create_pixel_inputs(pixel_values, num_frames=96, num_segs=12):
    """pixel_values: uint8 [num_frames, 3, H, W] on the GPU (what read_frames_decord returns, inference.py:73-77).
    Returns {'temporal_pixel_values': [1, num_frames, 3, 224, 224], 'spatial_pixel_values': [1, num_segs, 3, 336, 336]}."""
    if pixel_values.shape[0] != num_frames:
        raise ValueError("expected %d frames, got %d" % (num_frames, pixel_values.shape[0]))
    video_processor = frame_transform(image_size=224, mean=INTERNVIDEO_MEAN, std=INTERNVIDEO_STD)
    image_processor = frame_transform(image_size=336, mean=OPENAI_DATASET_MEAN, std=OPENAI_DATASET_STD)
    temporal = video_processor(pixel_values)[None]
    per_seg = int(num_frames // num_segs)
    idx = [(i * per_seg) + int(per_seg / 2) for i in range(num_segs)]            # inference.py:82-83
    spatial = image_processor(pixel_values[idx])[None]
    return {"temporal_pixel_values": temporal, "spatial_pixel_values": spatial}
