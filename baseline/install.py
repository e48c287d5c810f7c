#!/usr/bin/env python
"""Copy the reference's four source files + configs, verbatim, into git-ignored ``baseline/_ref/``.

    python -m baseline.install            # in the build container (needs /root/reference)

The reference is not a package (12 loose files, no setup.py / pyproject), so ``pip install --target baseline/_ref
/root/reference`` has nothing to build: the install IS this copy.  The files are data for the baseline arm and the
drop-in tests; nothing here is product source and nothing is edited (a sha256 manifest is written next to them).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ("models.py", "utils.py", "main.py", "dataset.py")


def install(ref_dir: str | None = None, dst: str = DST) -> str | None:
    ref_dir = ref_dir or os.environ.get("VML_REFERENCE_DIR", "/root/reference")
    if not os.path.isdir(ref_dir):
        return dst if os.path.exists(os.path.join(dst, "models.py")) else None
    os.makedirs(os.path.join(dst, "config"), exist_ok=True)
    manifest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(ref_dir, f), os.path.join(dst, f))
    for f in sorted(os.listdir(os.path.join(ref_dir, "config"))):
        shutil.copyfile(os.path.join(ref_dir, "config", f), os.path.join(dst, "config", f))
    for base, _, names in os.walk(dst):
        for n in sorted(names):
            if n == "MANIFEST.json" or n.endswith(".pyc"):
                continue
            p = os.path.join(base, n)
            with open(p, "rb") as fh:
                manifest[os.path.relpath(p, dst)] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as fh:
        json.dump({"source": ref_dir, "sha256": manifest}, fh, indent=1, sort_keys=True)
    return dst


if __name__ == "__main__":
    out = install(sys.argv[1] if len(sys.argv) > 1 else None)
    print(out if out else "reference not found; nothing installed")
