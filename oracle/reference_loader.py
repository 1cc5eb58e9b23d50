"""TEST INFRASTRUCTURE ONLY — import the UNMODIFIED reference `depth_pro` package.

Loads `/root/reference/src/depth_pro` under the alias ``ref_depth_pro`` (so it can live
next to the product's own drop-in ``depth_pro`` package in one process) with the `timm`
shim and the `pillow_heif` stub from this directory on ``sys.path``.  `/root/reference`
exists only in the build container: `available()` is False on the GPU box, and nothing
in the `-m gpu` tests, `smoke()` or `bench.py` may call `load()` there.
"""

from __future__ import annotations

import importlib.util
import os
import sys

REFERENCE_SRC = os.environ.get("DEPTHPRO_REFERENCE_SRC", "/root/reference/src")
_ALIAS = "ref_depth_pro"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "depth_pro", "__init__.py"))


def load():
    """Return the reference package (module object), importing it on first use."""
    if _ALIAS in sys.modules:
        return sys.modules[_ALIAS]
    if not available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)  # exposes the timm shim + pillow_heif stub
    pkg_dir = os.path.join(REFERENCE_SRC, "depth_pro")
    spec = importlib.util.spec_from_file_location(
        _ALIAS, os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_ALIAS] = mod
    spec.loader.exec_module(mod)
    return mod


def build_reference_model(state_dict=None):
    """Reference model on CPU fp32 (`depth_pro.py:72-151`), optionally loading a state_dict."""
    import dataclasses

    ref = load()
    dp = sys.modules[_ALIAS + ".depth_pro"]
    cfg = dataclasses.replace(dp.DEFAULT_MONODEPTH_CONFIG_DICT, checkpoint_uri=None)
    model, transform = ref.create_model_and_transforms(config=cfg)
    model.eval()
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    return model, transform
