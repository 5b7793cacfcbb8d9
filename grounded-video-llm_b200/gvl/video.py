"""Frame sampling front end (host side): the reference's `read_frames_decord` (mm_utils/video_utils.py:56-96) with
OpenCV as the container / H.264 decoder (decord and PyAV are third-party decoders that are not part of this image; NVDEC is the
planned GPU replacement). Everything that is arithmetic is the reference's: vlen, fps, duration = vlen / fps, the `clip` window,
`get_frame_indices` (gvl.hostlogic, bit-exact with the reference's function), frames returned as uint8 [T, 3, H, W] RGB.
The decoded pixels themselves come from FFmpeg through OpenCV instead of FFmpeg through decord.
"""
import numpy as np
import torch

from . import hostlogic


def read_frames(video_path, num_frames, sample="middle", clip=None):
    """Returns (frames uint8 [T,3,H,W], frame_indices, fps, vlen, duration) like read_frames_decord."""
    import cv2
    cap = cv2.VideoCapture(video_path)
    if not cap.isOpened():
        raise FileNotFoundError("cannot open video %s" % video_path)
    try:
        vlen = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
        fps = float(cap.get(cv2.CAP_PROP_FPS))
        if vlen <= 0 or fps <= 0:
            raise ValueError("no frame count / fps in %s" % video_path)
        duration = vlen / float(fps)
        start_index = 0
        if clip:
            start, end = clip
            duration = end - start
            vlen = int(duration * fps)
            start_index = int(start * fps)
        idx = hostlogic.get_frame_indices(num_frames, vlen, sample=sample)
        if clip:
            idx = [f + start_index for f in idx]
        frames, want, pos = [], sorted(set(idx)), 0
        got = {}
        cap.set(cv2.CAP_PROP_POS_FRAMES, want[0])
        pos = want[0]
        for target in want:
            while pos < target:                       # sequential grab: exact frame positions without seek rounding
                if not cap.grab():
                    raise ValueError("video ended at frame %d (wanted %d)" % (pos, target))
                pos += 1
            ok, bgr = cap.read()
            if not ok:
                raise ValueError("cannot decode frame %d of %s" % (target, video_path))
            pos += 1
            got[target] = bgr[:, :, ::-1]             # BGR -> RGB
        for f in idx:
            frames.append(got[f])
    finally:
        cap.release()
    arr = np.ascontiguousarray(np.stack(frames, axis=0))          # [T, H, W, 3]
    return torch.from_numpy(arr).permute(0, 3, 1, 2).contiguous(), idx, float(fps), vlen, duration
