"""Place the UNMODIFIED reference modules of the hot path under ``oracle/_ref/`` so that
``bench.py --impl reference`` and the ``cpu_baseline`` leg time the reference itself
(``kind: "reference"``) on the GPU box, where ``/root/reference`` does not exist.

    python oracle/build_ref.py          (also run by __graft_entry__.build() when the tree is present)

``oracle/_ref/`` is git-ignored (no reference source enters this repository's history) but NOT
gpurun-ignored, so it travels to the GPU box with the snapshot like the built ``.so``.  Files are
byte-for-byte copies; their sha256 is recorded in ``oracle/_ref/MANIFEST.json``.  The only
non-reference file is the 12-line ``easydict.py`` stand-in for the uninstalled package that
``miscc/config.py`` imports.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/AttnGAN2/code"
DST = os.path.join(ROOT, "oracle", "_ref")
FILES = ["GlobalAttention.py", "miscc/__init__.py", "miscc/config.py", "miscc/losses.py"]

EASYDICT = '''"""Stand-in for the `easydict` package (not installed): attribute access on a dict."""


class EasyDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v
'''


def build_ref() -> str | None:
    """Copy the reference files; returns the directory, or None when the reference tree is absent."""
    if not os.path.isdir(REF):
        return DST if os.path.exists(os.path.join(DST, "MANIFEST.json")) else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(DST, "easydict.py"), "w") as f:
        f.write(EASYDICT)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "sha256": manifest}, f, indent=1)
    return DST


def load_ref():
    """Import the vendored reference modules (GlobalAttention, miscc.losses, cfg) or return None.
    The copies are verified against the manifest first: an edited file is not 'the reference'."""
    mpath = os.path.join(DST, "MANIFEST.json")
    if not os.path.exists(mpath):
        return None
    manifest = json.load(open(mpath))["sha256"]
    for rel, digest in manifest.items():
        if hashlib.sha256(open(os.path.join(DST, rel), "rb").read()).hexdigest() != digest:
            raise RuntimeError(f"oracle/_ref/{rel} differs from the reference file it was copied from")
    if DST not in sys.path:
        sys.path.insert(0, DST)
    import importlib
    ga = importlib.import_module("GlobalAttention")
    if os.path.dirname(os.path.abspath(ga.__file__)) != DST:
        raise RuntimeError("a different GlobalAttention module is already imported")
    losses = importlib.import_module("miscc.losses")
    cfg = importlib.import_module("miscc.config").cfg
    cfg.CUDA = False
    return ga, losses, cfg


if __name__ == "__main__":
    print(build_ref())
