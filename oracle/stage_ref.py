"""Stage the reference files the lifted CPU arm needs into git-ignored ``oracle/_ref/``  --  TEST INFRASTRUCTURE.

    python oracle/stage_ref.py          (build container only: needs /root/reference)

``/root/reference`` does not exist on the GPU box, but ``oracle/_ref/`` travels there with the snapshot (it is
git-ignored, not gpurun-ignored), so ``bench.py --impl reference`` and ``cpu_baseline`` can time the reference's OWN
functions - ``slide_process`` / ``senet`` / ``evaluation`` lifted from ``main_moc.py`` plus its two selector / pooling
modules, byte-for-byte as shipped - instead of the oracle port.  Nothing staged here is tracked or imported by the
product; ``oracle/ref_loader.py`` falls back to this directory when the checkout is absent.
``__graft_entry__.build()`` runs this whenever the checkout is present.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("MOC_REFERENCE_ROOT", "/root/reference")
FILES = ("main_moc.py", "utils/patch_selection_classifier.py", "utils/patch_selection_classifier_index.py", "LICENSE")


def stage(source: str = SOURCE, dest: str = DEST) -> dict:
    """Copy FILES from ``source`` to ``dest`` (only when they changed); returns {relative path: sha256}."""
    if not os.path.isfile(os.path.join(source, "main_moc.py")):
        raise FileNotFoundError("reference checkout not found at %s" % source)
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(source, rel), os.path.join(dest, rel)
        if not os.path.isfile(src):
            continue
        data = open(src, "rb").read()
        manifest[rel] = hashlib.sha256(data).hexdigest()
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.isfile(dst) and open(dst, "rb").read() == data):
            shutil.copyfile(src, dst)
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": source, "sha256": manifest}, f, indent=1, sort_keys=True)
    return manifest


if __name__ == "__main__":
    for rel, digest in stage().items():
        print(digest[:16], rel)
