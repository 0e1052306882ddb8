#!/usr/bin/env python
"""Install the unmodified reference into baseline/_ref (git-ignored; travels to the GPU box with the snapshot).

    python baseline/install_ref.py [--no-fi]

1. ``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>``.
   Outcome (recorded in DESIGN.md): the wheel builds, but the reference's pyproject lists
   ``packages = ["scratchpad"]`` without its sub-packages, so the installed tree holds ``scratchpad/__init__.py``
   only -- an import of ``scratchpad.nn.attention`` fails on the *installed* reference whatever the environment.
2. Therefore the package tree is completed from the same source (``*.py`` under ``/root/reference/scratchpad``),
   byte for byte.  Nothing is edited; ``baseline/ref_loader.py`` imports single modules from it with stubs for the
   dependencies that are absent from this image (triteia, the engine's utils package).
3. The FlashInfer JIT modules the reference's call pattern needs are compiled here (nvcc cross-compiles without a
   GPU) into ``baseline/_ref/fi_cache`` so the GPU box loads them instead of spending box minutes on nvcc
   (``baseline/fi_prebuild.py``).

Runs only where ``/root/reference`` exists (the build container); a no-op elsewhere.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")


def install(prebuild_flashinfer: bool = True, verbose: bool = True) -> bool:
    if not os.path.isdir(os.path.join(REF_SRC, "scratchpad")):
        if verbose:
            print(f"install_ref: {REF_SRC} not present; using whatever baseline/_ref already holds")
        return os.path.isdir(os.path.join(DST, "scratchpad", "nn"))
    os.makedirs(DST, exist_ok=True)
    note = []
    marker = os.path.join(DST, "scratchpad", "nn", "attention", "flashinfer_backend.py")
    if not os.path.exists(marker):
        with tempfile.TemporaryDirectory() as tmp:
            src = os.path.join(tmp, "reference")
            shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns(".git", "docs", "node_modules"))
            cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                   "--find-links", "/opt/wheelhouse", "--target", DST, "--upgrade", src]
            r = subprocess.run(cmd, capture_output=True, text=True)
            note.append("$ " + " ".join(cmd))
            note.append(f"rc={r.returncode}")
            note.append(r.stdout[-1500:])
            note.append(r.stderr[-1500:])
        installed = []
        for root, _, files in os.walk(os.path.join(DST, "scratchpad")):
            installed += [os.path.relpath(os.path.join(root, f), DST) for f in files if f.endswith(".py")]
        note.append(f"python files installed by the wheel: {sorted(installed)}")
        # complete the tree the wheel omits (pyproject: packages = ["scratchpad"], no sub-packages)
        n = 0
        for root, dirs, files in os.walk(os.path.join(REF_SRC, "scratchpad")):
            dirs[:] = [d for d in dirs if d != "__pycache__"]
            for f in files:
                if not f.endswith(".py"):
                    continue
                s = os.path.join(root, f)
                d = os.path.join(DST, os.path.relpath(s, REF_SRC))
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copyfile(s, d)
                n += 1
        note.append(f"completed the package tree from {REF_SRC}/scratchpad: {n} python files copied unmodified")
        with open(os.path.join(DST, "INSTALL_NOTE.txt"), "w") as fh:
            fh.write("\n".join(note) + "\n")
        if verbose:
            print(f"install_ref: reference installed into {DST} ({n} files)")
    if prebuild_flashinfer:
        r = subprocess.run([sys.executable, os.path.join(HERE, "fi_prebuild.py")])
        if r.returncode != 0 and verbose:
            print("install_ref: FlashInfer prebuild failed; the GPU box will JIT-compile what is missing")
    return True


if __name__ == "__main__":
    ok = install(prebuild_flashinfer="--no-fi" not in sys.argv)
    sys.exit(0 if ok else 1)
