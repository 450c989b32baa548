"""Put the UNMODIFIED reference modules of the hot path under ``baseline/_ref/`` (git-ignored, but it
travels to the GPU box with the tree) so that ``bench.py --impl reference`` and the ``cpu_baseline``
leg time the reference's own code there.

Only the pure NumPy/SciPy modules of viloca/local_haplotype_inference are needed (update_eqs,
elbo_eqs, initialization, cavi of both modes); the files are copied byte for byte and their SHA-256
digests are recorded in ``baseline/_ref/MANIFEST.json``.  Runs where /root/reference is mounted (the
build container); on the GPU box the prebuilt copy is used.

    python baseline/install_ref.py
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
REFERENCE = os.environ.get("VILOCA_REFERENCE", "/root/reference")
PACKAGE = "viloca/local_haplotype_inference"
MODES = ("use_quality_scores", "learn_error_params")


def install(reference: str = REFERENCE, dest: str = DEST) -> dict | None:
    """Copy the modules; returns the manifest, or None when the reference is not mounted."""
    src_pkg = os.path.join(reference, PACKAGE)
    if not os.path.isdir(src_pkg):
        return None
    manifest = {"reference": reference, "files": {}}
    for mode in MODES:
        out_dir = os.path.join(dest, PACKAGE, mode)
        os.makedirs(out_dir, exist_ok=True)
        for name in sorted(os.listdir(os.path.join(src_pkg, mode))):
            if not name.endswith(".py"):
                continue
            src = os.path.join(src_pkg, mode, name)
            dst = os.path.join(out_dir, name)
            shutil.copyfile(src, dst)
            with open(dst, "rb") as fh:
                manifest["files"][os.path.join(PACKAGE, mode, name)] = hashlib.sha256(fh.read()).hexdigest()
    init = os.path.join(reference, "viloca", "__init__.py")
    if os.path.exists(init):
        shutil.copyfile(init, os.path.join(dest, "viloca", "__init__.py"))
    with open(os.path.join(dest, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    return manifest


def available(dest: str = DEST) -> bool:
    return os.path.exists(os.path.join(dest, "MANIFEST.json"))


def import_reference(mode: str, dest: str = DEST):
    """(update_eqs, elbo_eqs, initialization) of the unmodified reference for ``mode`` = "uqs" | "lep"."""
    import importlib
    if dest not in sys.path:
        sys.path.insert(0, dest)
    pkg = "viloca.local_haplotype_inference." + ("use_quality_scores" if mode == "uqs" else "learn_error_params")
    return tuple(importlib.import_module(pkg + "." + m) for m in ("update_eqs", "elbo_eqs", "initialization"))


if __name__ == "__main__":
    m = install()
    print("reference not mounted: nothing installed" if m is None else f"{len(m['files'])} files -> {DEST}")
