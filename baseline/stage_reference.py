"""Stage the UNMODIFIED reference files of the hot path into baseline/_ref/ (git-ignored, shipped to the GPU box by
gpurun) so that `bench.py --impl reference` and the `cpu_baseline` legs time the reference's own code on the box's host
cores.  Runs in the build container only (where /root/reference exists); nothing of the reference enters git history.

    python baseline/stage_reference.py

Files: utils/{__init__,SIMCA,CVSIMCA,data_utils}.py and vae_model.py, byte for byte.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("OCM_REFERENCE_DIR", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["utils/__init__.py", "utils/SIMCA.py", "utils/CVSIMCA.py", "utils/data_utils.py", "vae_model.py", "LICENSE"]


def stage(verbose=True):
    """Copy the files; returns True when baseline/_ref is usable afterwards."""
    if not os.path.isdir(REF):
        return os.path.exists(os.path.join(DST, "utils", "SIMCA.py"))
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"staged {len(manifest)} reference files into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
