"""TEST INFRASTRUCTURE ONLY - loader for the *unmodified* reference, used to pin the oracle.

Only `oracle/make_golden.py` (run in the build container, where `/root/reference`
exists) uses this module.  Nothing in the product path, `bench.py`, `smoke()` or the
`-m gpu` tests imports it: `/root/reference` does not exist on the GPU box.

The reference cannot be imported from its mount point:
  * `src/config/path_config.py:10-12` mkdirs `data/{logs,datasets,gan_outs}` next to the
    package at import time, and `src/classifier.py:15` opens a log file there;
  * `src/tmg_gan.py:6` imports `matplotlib`, `src/ctgan.py:1` imports `context`
    (neither exists in this image).
So we copy the tree to a scratch directory and put two empty stub modules first on
`sys.path` (SURVEY.md section 8c recipe).  No reference source is copied into the repo.
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile
import types

REFERENCE_ROOT = "/root/reference"


def load_reference(scratch: str | None = None):
    """Import the reference package `src` from a scratch copy and return the module."""
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"{REFERENCE_ROOT} is not present (only exists in the build container)")
    if "src" in sys.modules and hasattr(sys.modules["src"], "CVAEGAN"):
        return sys.modules["src"]
    scratch = scratch or tempfile.mkdtemp(prefix="cvaegan_ref_")
    dst = os.path.join(scratch, "ref")
    if not os.path.isdir(dst):
        shutil.copytree(REFERENCE_ROOT, dst)
        for root, dirs, files in os.walk(dst):
            for n in dirs + files:
                os.chmod(os.path.join(root, n), 0o755)
    # stub modules the reference imports at top level but never needs on this path
    for name in ("matplotlib", "matplotlib.pyplot", "context", "seaborn"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    sys.dont_write_bytecode = True
    sys.path.insert(0, dst)
    import src  # noqa: E402  (the reference package)

    src.config.device = "cpu"
    return src
