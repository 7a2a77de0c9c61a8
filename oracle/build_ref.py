"""Recipe for oracle/_ref: the UNMODIFIED reference, placed where it can travel to the GPU box.

TEST INFRASTRUCTURE / CPU BASELINE ONLY -- never imported by the product.

The reference is pure Python (no native code to compile), so "building" it is a verbatim
file copy of its own sources from where they lie under /root/reference into oracle/_ref/
(git-ignored: the sources never enter this repository's history; not gpurun-ignored: the
directory ships to the GPU box like the built .so).  A manifest with the SHA-256 of every
copied file is written next to them so a run can state exactly what it timed.

    python oracle/build_ref.py            # in the build container (needs /root/reference)

`load()` imports the copied modules under private names (they are called config / utils /
prior_model / test_model, like this repo's drop-in shims, so they are imported in an
isolated sys.modules window and handed back as a namespace).
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("RECOMBINER_REFERENCE", "/root/reference")
FILES = ["config.py", "utils.py", "prior_model.py", "test_model.py", "main_compression.py", "main_prior_training.py",
         "data/load_data.py", "data/image.py", "data/audio.py", "data/video.py", "data/protein.py", "LICENSE"]
MODULES = ("config", "utils", "prior_model", "test_model")


def build(source: str = SOURCE, dest: str = DEST) -> bool:
    """Copy the reference's sources verbatim.  Returns False (and leaves an existing copy alone)
    when the reference tree is not present -- on the GPU box only the prebuilt copy is used."""
    if not os.path.isdir(source):
        return False
    manifest = {}
    for rel in FILES:
        src = os.path.join(source, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(dest, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        os.chmod(dst, 0o644)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": source, "sha256": manifest}, f, indent=1, sort_keys=True)
    return True


def available(dest: str = DEST) -> bool:
    return all(os.path.exists(os.path.join(dest, m + ".py")) for m in MODULES)


_cached = None


def load(dest: str = DEST) -> types.SimpleNamespace:
    """Namespace (config, utils, prior_model, test_model) of the unmodified reference modules."""
    global _cached
    if _cached is not None:
        return _cached
    if not available(dest):
        raise ImportError("oracle/_ref is missing: run `python oracle/build_ref.py` in the build container")
    saved = {m: sys.modules.pop(m) for m in list(sys.modules) if m in MODULES or m == "data" or m.startswith("data.")}
    sys.path.insert(0, dest)
    try:
        importlib.invalidate_caches()
        mods = {m: importlib.import_module(m) for m in MODULES}
    finally:
        sys.path.remove(dest)
        for m in list(sys.modules):
            if m in MODULES or m == "data" or m.startswith("data."):
                del sys.modules[m]
        sys.modules.update(saved)
    _cached = types.SimpleNamespace(**mods)
    return _cached


if __name__ == "__main__":
    ok = build()
    print("oracle/_ref:", "copied from " + SOURCE if ok else "reference tree not found, nothing done")
