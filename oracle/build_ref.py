"""Recipe for oracle/_ref: the UNMODIFIED Python reference made importable next to the repo.

TEST / BENCH INFRASTRUCTURE (never imported by the product package).  The reference is pure Python over torch, so
"building" it is copying its three packages -- models/, trainers/, utils/ -- as they lie under /root/reference into
oracle/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box like a built .so, where /root/reference
does not exist).  Nothing is edited.  `load()` imports it with the one shim SURVEY.md 8(c) documents:
`utils.evaluator` needs TensorFlow, which no box here has, so a stub module with `Evaluator = None` is pre-seeded.

    python oracle/build_ref.py          # in the authoring container; __graft_entry__.build() calls build()
"""
from __future__ import annotations

import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("DD_REFERENCE", "/root/reference")
PACKAGES = ("models", "trainers", "utils")


def build() -> bool:
    """Copy the reference's packages into oracle/_ref (only when the source tree is present).  Returns availability."""
    if os.path.isdir(os.path.join(SRC, "models")):
        for pkg in PACKAGES:
            dst = os.path.join(DST, pkg)
            if os.path.isdir(dst):
                shutil.rmtree(dst)
            shutil.copytree(os.path.join(SRC, pkg), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.ipynb"))
        with open(os.path.join(DST, "SOURCE.txt"), "w") as f:
            f.write(f"verbatim copy of {SRC}/{{{','.join(PACKAGES)}}} made by oracle/build_ref.py; not tracked by git\n")
    return available()


def available() -> bool:
    return os.path.isfile(os.path.join(DST, "models", "__init__.py"))


def load():
    """Import the reference from oracle/_ref; returns its `models` package with `.EMA` attached (trainers/ema.py)."""
    if not available():
        raise ImportError("oracle/_ref is absent: run python oracle/build_ref.py where /root/reference exists")
    if DST not in sys.path:
        sys.path.insert(0, DST)
    if "utils.evaluator" not in sys.modules:
        stub = types.ModuleType("utils.evaluator")
        stub.Evaluator = None
        sys.modules["utils.evaluator"] = stub
    import models as ref_models                     # noqa: E402
    from trainers.ema import EMA as RefEMA          # noqa: E402
    ref_models.EMA = RefEMA
    return ref_models


if __name__ == "__main__":
    print("oracle/_ref available:", build())
