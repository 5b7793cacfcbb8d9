"""ORACLE SUPPORT -- TEST INFRASTRUCTURE ONLY.

Populates oracle/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box with the snapshot like libgvl.so)
with the reference's own Python files for the hot path, copied UNMODIFIED from /root/reference at build time:

    models/*.py  mm_utils/*.py  datasets/chat/base_template.py  inference.py

Nothing under oracle/_ref/ is committed and nothing in the product imports it. It exists so that the `-m gpu` parity
tests (tests/test_gpu_vs_reference.py), bench.py's reference arm and `extra.gpu_reference` can run the REFERENCE's
modules (random-init, named architecture) on the B200 box, where /root/reference does not exist.
Called from __graft_entry__.build(); a no-op (keeps what is there) when /root/reference is absent.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("GVL_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")

FILES = ["inference.py", "datasets/chat/base_template.py", "mm_utils/utils.py", "mm_utils/video_utils.py",
         "models/internvideo2.py", "models/llava_next_video.py", "models/modeling_clip.py", "models/modeling_llama.py",
         "models/modeling_phi3.py"]


def build_ref(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "models")):
        if verbose:
            print("oracle/_ref: reference not present at %s (%s)" % (
                SRC, "keeping the existing copy" if os.path.isdir(DST) else "nothing to install"))
        return DST if os.path.isdir(DST) else None
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
    if verbose:
        print("oracle/_ref: installed %d reference files from %s" % (len(FILES), SRC))
    return DST


if __name__ == "__main__":
    build_ref()
