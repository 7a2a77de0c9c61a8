"""Recipe for oracle/_ref: the UNMODIFIED reference, placed where it can travel to the GPU box.

TEST INFRASTRUCTURE / CPU BASELINE ONLY -- never imported by the product.

The reference is pure Python (no native code to compile), so "building" it is packing its
own sources, byte for byte, from where they lie under /root/reference into ONE archive,
oracle/_ref/reference_py.zip (git-ignored: the sources never enter this repository's history;
not gpurun-ignored: the artefact ships to the GPU box like the built .so).  Python imports
modules straight from the archive (zipimport), so nothing is unpacked.  A manifest with the
SHA-256 of every packed file is written next to it so a run can state exactly what it timed.

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
import sys
import types
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("RECOMBINER_REFERENCE", "/root/reference")
FILES = ["config.py", "utils.py", "prior_model.py", "test_model.py", "main_compression.py", "main_prior_training.py",
         "data/load_data.py", "data/image.py", "data/audio.py", "data/video.py", "data/protein.py", "LICENSE"]
MODULES = ("config", "utils", "prior_model", "test_model")
ARCHIVE = "reference_py.zip"


def build(source: str = SOURCE, dest: str = DEST) -> bool:
    """Pack the reference's sources verbatim.  Returns False (and leaves an existing archive alone)
    when the reference tree is not present -- on the GPU box only the prebuilt archive is used."""
    if not os.path.isdir(source):
        return False
    os.makedirs(dest, exist_ok=True)
    manifest = {}
    tmp = os.path.join(dest, ARCHIVE + ".tmp")
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for rel in FILES:
            src = os.path.join(source, rel)
            if not os.path.exists(src):
                continue
            with open(src, "rb") as f:
                blob = f.read()
            manifest[rel] = hashlib.sha256(blob).hexdigest()
            z.writestr(zipfile.ZipInfo(rel, date_time=(2020, 1, 1, 0, 0, 0)), blob)      # fixed timestamp: reproducible archive
    os.replace(tmp, os.path.join(dest, ARCHIVE))
    for stale in FILES + ["data"]:                      # loose copies of an earlier layout
        path = os.path.join(dest, stale)
        if os.path.isfile(path):
            os.remove(path)
        elif os.path.isdir(path) and not os.listdir(path):
            os.rmdir(path)
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": source, "archive": ARCHIVE, "sha256": manifest}, f, indent=1, sort_keys=True)
    return True


def archive_path(dest: str = DEST) -> str:
    return os.path.join(dest, ARCHIVE)


def available(dest: str = DEST) -> bool:
    return os.path.exists(archive_path(dest))


def read_source(rel: str, dest: str = DEST) -> str:
    """Text of one packed reference file (e.g. 'data/image.py')."""
    with zipfile.ZipFile(archive_path(dest)) as z:
        return z.read(rel).decode("utf-8")


_cached = None


def load(dest: str = DEST) -> types.SimpleNamespace:
    """Namespace (config, utils, prior_model, test_model) of the unmodified reference modules."""
    global _cached
    if _cached is not None:
        return _cached
    if not available(dest):
        raise ImportError("oracle/_ref is missing: run `python oracle/build_ref.py` in the build container")
    saved = {m: sys.modules.pop(m) for m in list(sys.modules) if m in MODULES or m == "data" or m.startswith("data.")}
    zpath = archive_path(dest)
    sys.path.insert(0, zpath)
    try:
        importlib.invalidate_caches()
        mods = {m: importlib.import_module(m) for m in MODULES}
    finally:
        sys.path.remove(zpath)
        for m in list(sys.modules):
            if m in MODULES or m == "data" or m.startswith("data."):
                del sys.modules[m]
        sys.modules.update(saved)
    _cached = types.SimpleNamespace(**mods)
    return _cached


if __name__ == "__main__":
    ok = build()
    print("oracle/_ref:", "packed from " + SOURCE if ok else "reference tree not found, nothing done")
