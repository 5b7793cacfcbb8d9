"""ORACLE SUPPORT -- TEST INFRASTRUCTURE ONLY (runs where /root/reference exists -- this container -- or where
oracle/build_ref.py has installed the unmodified reference files under oracle/_ref -- the GPU box).

Makes the reference (WHB139426/Grounded-Video-LLM @ e26da4e) importable under the container's newer stack
(python 3.12, transformers 5.5, no timm/decord/av/peft) WITHOUT modifying or copying it:
  * a 3-symbol `timm.models.layers` stub (internvideo2.py:15 imports DropPath, to_2tuple, trunc_normal_),
  * `import_models()` puts /root/reference on sys.path and imports models.modeling_clip / internvideo2 /
    modeling_phi3 / modeling_llama,
  * `extract()` pulls ONE function or method out of a reference file with `ast` and exec's it in a namespace we
    control, so that e.g. inference.parse_time_interval or LLAVA_NEXT_VIDEO.reshape_hd_patches_2x2merge_phi3 can
    be called as the reference wrote them even though their modules do not import here.
Used by oracle/make_golden.py and tests/test_oracle_vs_reference.py.
"""
import ast
import importlib
import importlib.machinery
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_ref():
    """/root/reference in the build container; oracle/_ref (installed by oracle/build_ref.py, travels with the gpurun
    snapshot) on the GPU box."""
    for cand in (os.environ.get("GVL_REFERENCE"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "models")):
            return cand
    return "/root/reference"


REF = _find_ref()


def available():
    return os.path.isdir(os.path.join(REF, "models"))


def _install_timm_stub():
    if "timm.models.layers" in sys.modules:
        return
    import torch
    from torch import nn

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            return x  # eval-mode identity

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    def trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return torch.nn.init.trunc_normal_(t, mean=mean, std=std, a=a, b=b)

    for name in ("timm", "timm.models", "timm.models.layers"):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
        m.__path__ = []
        sys.modules[name] = m
    lay = sys.modules["timm.models.layers"]
    lay.DropPath, lay.to_2tuple, lay.trunc_normal_ = DropPath, to_2tuple, trunc_normal_
    sys.modules["timm"].models = sys.modules["timm.models"]
    sys.modules["timm.models"].layers = lay


_mods = {}


def import_models():
    """Returns dict(clip=..., iv2=..., phi3=..., llama=...) of the reference's model modules."""
    if _mods:
        return _mods
    if not available():
        raise RuntimeError("reference not present at %s" % REF)
    import transformers  # noqa: F401  (must be imported before the timm stub is registered)
    _install_timm_stub()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    _mods["clip"] = importlib.import_module("models.modeling_clip")
    _mods["iv2"] = importlib.import_module("models.internvideo2")
    _mods["phi3"] = importlib.import_module("models.modeling_phi3")
    _mods["llama"] = importlib.import_module("models.modeling_llama")
    return _mods


def extract(rel_path, name, class_name=None, namespace=None):
    """exec the source of function `name` (optionally a method of `class_name`) from a reference file."""
    path = os.path.join(REF, rel_path)
    src = open(path).read()
    tree = ast.parse(src)
    body = tree.body
    if class_name is not None:
        cls = [n for n in body if isinstance(n, ast.ClassDef) and n.name == class_name][0]
        body = cls.body
    fn = [n for n in body if isinstance(n, ast.FunctionDef) and n.name == name][0]
    fn.decorator_list = []
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {} if namespace is None else dict(namespace)
    exec(compile(mod, path, "exec"), ns)
    return ns[name]


def std_namespace():
    import copy
    import math
    import re

    import einops
    import numpy as np
    import torch
    from torch import nn
    return dict(torch=torch, nn=nn, einops=einops, rearrange=einops.rearrange, math=math, re=re, np=np, copy=copy,
                IMAGE_TOKEN_INDEX=-200, IGNORE_INDEX=-100, DEFAULT_IMAGE_TOKEN="<image>",
                GROUNDING_TOKEN="<timestamp_grounding>", random=__import__("random"))
