"""Load the real reference module (build container only).  TEST INFRASTRUCTURE ONLY.

``/root/reference`` exists in the build container and NOT on the GPU box, so this loader is
used only by ``oracle/gen_golden.py`` and by the optional ``tests/test_oracle_vs_reference.py``
(skipped when the reference is absent).  The reference's ``ufvideo/model/layer.py`` imports
nothing but torch, so it is executed standalone by file path; importing the ``ufvideo`` package
would pull in decord / moviepy / pycocotools / timm, none of which are installed.
"""
from __future__ import annotations

import importlib.util
import os
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# $UFV_REF, the build container's read-only tree, or an install of the reference under baseline/_ref
# (git-ignored; nothing in this repository puts reference sources there)
_CANDIDATES = (os.environ.get("UFV_REF", ""), "/root/reference", os.path.join(_REPO, "baseline", "_ref"))


def reference_layer_path():
    for root in _CANDIDATES:
        if root:
            path = os.path.join(root, "ufvideo", "model", "layer.py")
            if os.path.isfile(path):
                return path
    return None


def load_reference_layer():
    """Return the reference's layer module, or None when the reference tree is absent."""
    path = reference_layer_path()
    if path is None:
        return None
    spec = importlib.util.spec_from_file_location("ufvideo_reference_layer", path)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


def reference_config(mm_hidden_size: int = 1152, hidden_size: int = 3584):
    """The only two config attributes the region encoder reads (layer.py:55-58)."""
    return types.SimpleNamespace(mm_hidden_size=mm_hidden_size, hidden_size=hidden_size)
