#!/usr/bin/env python
"""Install the UNMODIFIED reference (dmar-bonn/ipp-marl) under baseline/_ref/ so that it travels to the GPU box.

The reference is pure Python with no setup.py / pyproject, so "installing" it is a verbatim copy of its
``marl_framework/`` package (``*.py`` + ``params.yaml``; IDE debris, ``*.save`` and ``__pycache__`` are skipped).
``baseline/_ref/`` is git-ignored (never part of this repo's history) but NOT gpurun-ignored: the snapshot that goes
to the B200 box carries it next to the built ``.so``.  Nothing under ``ipp_marl_b200/`` imports it; it is used by

  * ``bench.py --impl reference`` and the ``cpu_baseline`` leg (the reference's own env loop timed on the host cores),
  * the ``-m gpu`` drop-in tests (the unchanged ``coma_wrapper.py`` / ``IG_baseline.py`` / ``lawn_mower.py`` running on
    top of ``ipp_marl_b200.facade``),
  * the oracle re-checks against the live reference.

Run by ``__graft_entry__.build()`` whenever the source tree (``/root/reference`` or ``$IPP_REFERENCE_SRC``) exists.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
KEEP_EXT = (".py", ".yaml", ".md")
SKIP_DIRS = {"__pycache__", ".idea", ".git"}


def source_root():
    return os.environ.get("IPP_REFERENCE_SRC", "/root/reference")


def installed():
    return os.path.isfile(os.path.join(DEST, "marl_framework", "coma_wrapper.py"))


def install(src=None, dest=DEST, quiet=False):
    """Copy the reference into ``dest``; returns the manifest (file -> sha256) or None if there is no source."""
    src = src or source_root()
    if not os.path.isdir(os.path.join(src, "marl_framework")):
        return None
    manifest = {}
    for d, dirs, files in os.walk(src):
        dirs[:] = [x for x in dirs if x not in SKIP_DIRS]
        for f in files:
            if not f.endswith(KEEP_EXT) and f != ".SUBMODULES.json":
                continue
            s = os.path.join(d, f)
            rel = os.path.relpath(s, src)
            t = os.path.join(dest, rel)
            os.makedirs(os.path.dirname(t), exist_ok=True)
            with open(s, "rb") as fh:
                data = fh.read()
            manifest[rel] = hashlib.sha256(data).hexdigest()
            if not os.path.exists(t) or open(t, "rb").read() != data:
                shutil.copyfile(s, t)
    commit = None
    sub = os.path.join(src, ".SUBMODULES.json")
    if os.path.exists(sub):
        with open(sub) as fh:
            commit = json.load(fh).get("commit")
    with open(os.path.join(dest, "INSTALLED.json"), "w") as fh:
        json.dump({"source": src, "commit": commit, "files": manifest}, fh, indent=1, sort_keys=True)
    if not quiet:
        print("reference installed: %d files -> %s (commit %s)" % (len(manifest), dest, commit))
    return manifest


def verify(dest=DEST):
    """True when every installed file still matches the manifest written at install time (i.e. is unmodified)."""
    path = os.path.join(dest, "INSTALLED.json")
    if not os.path.exists(path):
        return False
    with open(path) as fh:
        man = json.load(fh)["files"]
    for rel, digest in man.items():
        p = os.path.join(dest, rel)
        if not os.path.exists(p):
            return False
        with open(p, "rb") as fh:
            if hashlib.sha256(fh.read()).hexdigest() != digest:
                return False
    return True


if __name__ == "__main__":
    m = install()
    if m is None:
        print("no reference source tree at %s" % source_root())
        sys.exit(0 if installed() else 1)
