"""ORACLE helper — loads the reference's OWN Unet3D / VideoDecoder code (unmodified, from
/root/reference/dalle2_video/dalle2_video.py) on top of the restated primitives.

Search order: $DV_REFERENCE_ROOT, /root/reference (this container only), then <repo>/baseline/_ref — the
unmodified reference package installed there with
    python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>
(git-ignored, but it travels to the GPU box, where /root/reference does not exist).
Used by oracle/make_golden.py, by CPU tests that cross-check the restatement, and by `bench.py --impl reference` /
the `cpu_baseline` leg (kind "reference"); never by the product package.
"""
from __future__ import annotations

import importlib.util
import os
import sys

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_cached = None


def _find_root():
    roots = [os.environ.get("DV_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for r in roots:
        if r and os.path.isfile(os.path.join(r, "dalle2_video", "dalle2_video.py")):
            return r
    return os.environ.get("DV_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "dalle2_video", "dalle2_video.py"))


def load():
    """Return the reference module object (dalle2_video.dalle2_video), imported from where it lies."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    for p in (_REPO, _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    path = os.path.join(REFERENCE_ROOT, "dalle2_video", "dalle2_video.py")
    spec = importlib.util.spec_from_file_location("_reference_dalle2_video", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_reference_dalle2_video"] = mod
    spec.loader.exec_module(mod)
    _cached = mod
    return mod
