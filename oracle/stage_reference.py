#!/usr/bin/env python
"""Stage the UNMODIFIED reference package under ``oracle/_ref/`` (git-ignored, travels to the GPU box with the
snapshot) so that ``bench.py --impl reference`` and ``cpu_baseline`` can time the reference's own module instead of
the numpy port.  TEST / MEASUREMENT INFRASTRUCTURE ONLY -- nothing under ``arbinterp_b200/`` may import it.

The reference is pure Python, so "building" it is an offline ``pip install --target``; no reference source is
copied into the tracked tree.  Two attempts, in this order:

  1. the source tree (from a copy under /tmp, the original is read-only):
         pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of /root/reference>
     -- fails in this image: ``pyproject.toml`` names the poetry-core build backend, which is not installed;
  2. the reference's own built wheel, ``dist/ARBInterp-1.8-py3-none-any.whl`` (package ``ARBTools``; its
     ``ARBTools/ARBInterp.py`` is byte-identical to ``src/ARBInterp/ARBInterp.py`` -- checked below):
         pip install --no-index --no-deps --target oracle/_ref <wheel>

Run by ``__graft_entry__.build()`` when ``/root/reference`` exists; ``oracle/_ref/STAGED.json`` records what happened.
"""
import filecmp
import glob
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")


def module_path(root=DEST):
    """Path of the reference module inside a staged / installed tree, or None."""
    for rel in ("ARBTools/ARBInterp.py", "ARBInterp/ARBInterp.py", "src/ARBInterp/ARBInterp.py"):
        p = os.path.join(root, rel)
        if os.path.isfile(p):
            return p
    return None


def stage(ref="/root/reference", dest=DEST, quiet=True):
    if not os.path.isdir(ref):
        return None
    src_module = os.path.join(ref, "src", "ARBInterp", "ARBInterp.py")
    have = module_path(dest)
    if have and os.path.isfile(src_module) and filecmp.cmp(have, src_module, shallow=False):
        return have                                         # already staged and identical
    shutil.rmtree(dest, ignore_errors=True)
    os.makedirs(dest, exist_ok=True)
    pip = [sys.executable, "-m", "pip", "install", "--no-index", "--no-deps", "--disable-pip-version-check",
           "--find-links", "/opt/wheelhouse", "--target", dest]
    log = {"reference": ref, "attempts": []}
    with tempfile.TemporaryDirectory() as tmp:
        copy = os.path.join(tmp, "reference")
        shutil.copytree(ref, copy, ignore=shutil.ignore_patterns("dist", ".git", "docs"))
        r = subprocess.run(pip + ["--no-build-isolation", copy], capture_output=True, text=True)
        log["attempts"].append({"what": "source tree", "rc": r.returncode, "tail": (r.stderr or r.stdout)[-300:]})
    if module_path(dest) is None:
        for whl in sorted(glob.glob(os.path.join(ref, "dist", "*.whl"))):
            r = subprocess.run(pip + [whl], capture_output=True, text=True)
            log["attempts"].append({"what": os.path.basename(whl), "rc": r.returncode,
                                    "tail": (r.stderr or r.stdout)[-300:]})
            if module_path(dest):
                break
    got = module_path(dest)
    log["module"] = got and os.path.relpath(got, dest)
    log["identical_to_src"] = bool(got and os.path.isfile(src_module) and filecmp.cmp(got, src_module, shallow=False))
    with open(os.path.join(dest, "STAGED.json"), "w") as f:
        json.dump(log, f, indent=1)
    if not quiet:
        print(json.dumps(log, indent=1))
    if got and not log["identical_to_src"]:
        raise RuntimeError("staged reference module differs from /root/reference/src/ARBInterp/ARBInterp.py")
    return got


if __name__ == "__main__":
    print(stage(quiet=False))
