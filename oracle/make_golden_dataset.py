"""Golden description of the reference's loader on a small synthetic dataset directory  --  TEST INFRASTRUCTURE.

    python oracle/make_golden_dataset.py        (build container only: needs /root/reference)

Runs datasets/dataset_generic.py UNMODIFIED (loaded by file path under a stand-in ``datasets`` package, because the
HuggingFace ``datasets`` distribution shadows the reference's namespace directory) through the calls main_moc.py makes
(:268-289): Generic_MIL_Dataset(...), load_from_h5, load_full_path, return_splits(from_id=False, csv_path, repeat_num).
``h5py`` does not exist in this image; the module's ``import h5py`` is served by a shim whose ``File`` is
moc_b200.h5bag.H5File (the reader under test elsewhere) - the dataset / split logic is entirely the reference's.
pandas 3 needs ``future.infer_string = False`` for the reference's ``df_prep`` (dataset_generic.py:125-127).
Writes tests/golden/dataset_splits.json.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402


def load_reference_dataset_module():
    import pandas as pd
    pd.set_option("future.infer_string", False)
    from moc_b200.h5bag import H5File
    shim = types.ModuleType("h5py")
    shim.File = H5File
    sys.modules["h5py"] = shim
    ref = ref_loader.REFERENCE_ROOT
    saved = sys.modules.get("datasets")
    pkg = types.ModuleType("datasets")
    pkg.__path__ = [os.path.join(ref, "datasets")]
    sys.modules["datasets"] = pkg
    sys.path.insert(0, ref)
    try:
        spec = importlib.util.spec_from_file_location("datasets.dataset_generic", os.path.join(ref, "datasets", "dataset_generic.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["datasets.dataset_generic"] = mod
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(ref)
        if saved is not None:
            sys.modules["datasets"] = saved
        else:
            sys.modules.pop("datasets", None)
    return mod


def main():
    assert ref_loader.has_checkout()
    from tests import dataset_fixture as fx
    mod = load_reference_dataset_module()
    out = {}
    for repeat_num in (9, 4, None):
        with tempfile.TemporaryDirectory() as root:
            dataset, splits = fx.make(mod, root, repeat_num=repeat_num)
            out["repeat_%s" % repeat_num] = fx.describe(dataset, splits, root)
    path = os.path.join(ROOT, "tests", "golden", "dataset_splits.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path, {k: {s: (v[s] or {}).get("len") for s in ("train", "val", "test")} for k, v in out.items()})


if __name__ == "__main__":
    main()
