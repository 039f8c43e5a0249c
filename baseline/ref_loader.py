"""Import the staged, unmodified reference (baseline/_ref) the way SURVEY.md §8(c) prescribes: matplotlib / plotly are
imported at module level by utils/SIMCA.py:8-9 but never used on the timed path, so empty stub modules are registered
when the real packages are missing.  Used by bench.py's CPU arms only (never by the product)."""
from __future__ import annotations

import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF_DIR, "utils", "SIMCA.py"))


def load():
    """Returns (utils package, vae_model module) of the reference, or raises ImportError."""
    if not available():
        raise ImportError("baseline/_ref is not staged (python baseline/stage_reference.py in the build container)")
    for name in ("matplotlib", "matplotlib.pyplot", "plotly", "plotly.graph_objects"):
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules.setdefault(name, types.ModuleType(name))
    if isinstance(sys.modules.get("matplotlib"), types.ModuleType) and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if isinstance(sys.modules.get("plotly"), types.ModuleType) and not hasattr(sys.modules["plotly"], "graph_objects"):
        sys.modules["plotly"].graph_objects = sys.modules["plotly.graph_objects"]
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "utils" or k.startswith("utils.") or k == "vae_model"}
    sys.path.insert(0, REF_DIR)
    try:
        ref_utils = importlib.import_module("utils")
        ref_vae = importlib.import_module("vae_model")
    finally:
        sys.path.remove(REF_DIR)
        # keep the reference under private names so that it cannot shadow anything else called `utils`
        for k in list(sys.modules):
            if k == "utils" or k.startswith("utils.") or k == "vae_model":
                sys.modules["_ocm_ref_" + k] = sys.modules.pop(k)
        sys.modules.update(saved)
    return ref_utils, ref_vae
