"""Import the UNMODIFIED reference `models.*` from /root/reference (build container) or its staged copy (GPU box).

TEST INFRASTRUCTURE.  The reference imports `mcubes` and `icecream` at module top
(models/renderer.py:6-7); both are absent here and neither is used on the hot path,
so two empty `sys.modules` stubs make it importable.  The reference
package is loaded under the private name `_rnb_reference_models` so it never collides
with this repo's own drop-in `models` package.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "RNb-NeuS-fork")


def _find_root():
    """RNB_REFERENCE_ROOT, else the build container's /root/reference, else the copy oracle/stage_reference.py put under
    baseline/_ref (git-ignored; it travels to the GPU box with the snapshot)."""
    env = os.environ.get("RNB_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _STAGED):
        if os.path.isfile(os.path.join(cand, "models", "renderer.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "renderer.py"))


def load():
    """Returns a namespace with .fields, .renderer, .embedder of the reference."""
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if "_rnb_reference_models" in sys.modules:
        return sys.modules["_rnb_reference_models"]
    if "mcubes" not in sys.modules:
        sys.modules["mcubes"] = types.ModuleType("mcubes")
    if "icecream" not in sys.modules:
        ic = types.ModuleType("icecream")
        ic.ic = lambda *a, **k: None
        sys.modules["icecream"] = ic
    # The reference's `models/` has no __init__.py (namespace package); this repo's drop-in `models`
    # is a regular package and would win any path-based import, so bind the name explicitly.
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "models" or k.startswith("models.")}
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "models")]
    sys.modules["models"] = pkg
    try:
        emb = importlib.import_module("models.embedder")
        fields = importlib.import_module("models.fields")
        renderer = importlib.import_module("models.renderer")
    finally:
        ref_mods = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "models" or k.startswith("models.")}
        sys.modules.update(saved)
    ns = types.ModuleType("_rnb_reference_models")
    ns.embedder, ns.fields, ns.renderer = emb, fields, renderer
    ns._mods = ref_mods
    sys.modules["_rnb_reference_models"] = ns
    return ns
