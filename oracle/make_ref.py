#!/usr/bin/env python
"""Stage the UNMODIFIED reference operator package under oracle/_ref/ (test / baseline infrastructure only).

The reference is pure Python: its hot path is the package ``quantization_utils`` (quant_utils.py, quant_modules.py,
imported as ``from quantization_utils.quant_modules import *`` by main_direct.py:21 and trainer_direct.py:19).
``/root/reference`` does not exist on the GPU box, so the build container copies those two files, byte for byte,
into ``oracle/_ref/quantization_utils/`` -- git-ignored (reference sources never enter the history), not
gpurun-ignored (they travel to the box like the built ``.so``).  ``MANIFEST.json`` records the sha256 of every file
so a reader can check that nothing was edited.  Consumers: ``bench.py --impl reference`` (CPU arm, ``kind:
"reference"``), the ``gpu_eager_baseline`` leg (the same modules on ``cuda``), and the live-reference tests.

    python oracle/make_ref.py            # copies when /root/reference is mounted; otherwise reports what is staged
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("OODFQ_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = ("quantization_utils/quant_utils.py", "quantization_utils/quant_modules.py")


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(verbose=True):
    """Copy the reference's operator package; returns the staging directory or None when nothing is available."""
    if os.path.isdir(REF_ROOT):
        manifest = {"source": REF_ROOT, "files": {}}
        for rel in FILES:
            src, dst = os.path.join(REF_ROOT, rel), os.path.join(OUT, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            manifest["files"][rel] = _sha(dst)
        with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
            json.dump(manifest, f, indent=1)
        if verbose:
            print(f"oracle/_ref: staged {len(FILES)} unmodified files from {REF_ROOT}")
        return OUT
    if available():
        if verbose:
            print("oracle/_ref: reference tree not mounted; using the files staged earlier")
        return OUT
    if verbose:
        print("oracle/_ref: reference tree not mounted and nothing staged", file=sys.stderr)
    return None


def available():
    """True when every staged file exists and still matches the manifest written at staging time."""
    man = os.path.join(OUT, "MANIFEST.json")
    if not os.path.exists(man):
        return False
    with open(man) as f:
        files = json.load(f)["files"]
    return all(os.path.exists(os.path.join(OUT, rel)) and _sha(os.path.join(OUT, rel)) == sha
               for rel, sha in files.items()) and set(files) == set(FILES)


def load():
    """Import the staged package as ``oodfq_ref_quantization_utils.quant_modules`` (a private name: the product's
    drop-in registers itself as top-level ``quantization_utils`` and the two must not collide).  Returns the
    ``quant_modules`` module, whose star-import also exposes everything of ``quant_utils``."""
    import importlib.util
    if not available():
        raise ImportError("oracle/_ref is not staged (run `python oracle/make_ref.py` where /root/reference is mounted)")
    pkg_name = "oodfq_ref_quantization_utils"
    if pkg_name + ".quant_modules" in sys.modules:
        return sys.modules[pkg_name + ".quant_modules"]
    pkg_dir = os.path.join(OUT, "quantization_utils")
    # the reference package has no __init__.py (namespace package): synthesise the package object
    spec = importlib.util.spec_from_loader(pkg_name, loader=None, is_package=True)
    pkg = importlib.util.module_from_spec(spec)
    pkg.__path__ = [pkg_dir]
    sys.modules[pkg_name] = pkg
    mods = {}
    for name in ("quant_utils", "quant_modules"):
        full = f"{pkg_name}.{name}"
        s = importlib.util.spec_from_file_location(full, os.path.join(pkg_dir, name + ".py"))
        m = importlib.util.module_from_spec(s)
        sys.modules[full] = m
        s.loader.exec_module(m)
        setattr(pkg, name, m)
        mods[name] = m
    return mods["quant_modules"]


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
