"""Test infrastructure: a small CLAM-style dataset directory (dataset csv, split csv, h5_files/, pt_files/) built
deterministically, and the description of a dataset object that the loader parity tests compare."""
from __future__ import annotations

import os

import numpy as np
import torch

from tests.h5_writer import write_h5

LABEL_DICT = {"KICH": 0, "KIRC": 1, "KIRP": 2}
# (case_id, slide_id, label): zero-padded numeric ids must survive as strings; two slides of one patient; an
# ignored label; csv order differs from the split file's order
ROWS = [("p01", "007", "KIRC"), ("p02", "S-b", "KICH"), ("p03", "S-c", "KIRP"), ("p03", "S-d", "KIRP"),
        ("p04", "0042", "KICH"), ("p05", "S-f", "OTHER"), ("p06", "S-g", "KIRC"), ("p07", "S-h", "KIRC"),
        ("p08", "S-i", "KIRP"), ("p09", "S-j", "KICH"), ("p10", "S-k", "KIRC"), ("p11", "S-l", "KIRP"),
        ("p12", "S-m", "KICH"), ("p13", "S-n", "KIRC")]
SPLITS = {"train": ["S-i", "007", "S-b", "S-g", "S-c", "0042"],          # not in csv order
          "val": ["S-h", "S-j", "S-zz"],                                  # S-zz is not in the dataset csv
          "test": ["S-n", "S-m", "S-l", "S-k", "S-d", "S-f"]}             # S-f carries the ignored label


def n_patches(slide_id: str) -> int:
    return 5 + (sum(ord(c) for c in slide_id) * 7) % 37


def bag_of(slide_id: str):
    g = np.random.RandomState(sum(ord(c) for c in slide_id))
    n = n_patches(slide_id)
    feats = g.standard_normal((n, 512)).astype(np.float32)
    coords = g.randint(0, 50000, size=(n, 2)).astype(np.int64)
    return feats, coords


def build(root: str) -> dict:
    os.makedirs(os.path.join(root, "feats", "h5_files"), exist_ok=True)
    os.makedirs(os.path.join(root, "feats", "pt_files"), exist_ok=True)
    with open(os.path.join(root, "dataset.csv"), "w") as f:
        f.write("case_id,slide_id,label\n")
        for r in ROWS:
            f.write("%s,%s,%s\n" % r)
    n = max(len(v) for v in SPLITS.values())
    with open(os.path.join(root, "splits_0.csv"), "w") as f:
        f.write(",train,val,test\n")
        for i in range(n):
            f.write("%d,%s\n" % (i, ",".join(SPLITS[k][i] if i < len(SPLITS[k]) else "" for k in ("train", "val", "test"))))
    for _, sid, _ in ROWS:
        feats, coords = bag_of(sid)
        write_h5(os.path.join(root, "feats", "h5_files", sid + ".h5"), {"features": feats, "coords": coords})
        torch.save(torch.from_numpy(feats), os.path.join(root, "feats", "pt_files", sid + ".pt"))
    return {"csv": os.path.join(root, "dataset.csv"), "splits": os.path.join(root, "splits_0.csv"),
            "data_dir": os.path.join(root, "feats")}


def describe_split(ds, root: str) -> dict:
    """Everything the loops (and a user) can observe about a split object."""
    out = {"real_len": int(ds.real_len()), "len": int(len(ds)), "repeat_num": ds.repeat_num,
           "slide_ids": [str(s) for s in ds.slide_data["slide_id"]], "labels": [int(v) for v in ds.slide_data["label"]],
           "slide_cls_ids": [[int(i) for i in a] for a in ds.slide_cls_ids], "items": []}
    for idx in range(len(ds)):
        feats, label, coords, full_path = ds[idx]
        out["items"].append({"shape": list(feats.shape), "dtype": str(feats.dtype), "sum": float(feats.double().sum()),
                             "label": int(label), "coords_sum": int(np.asarray(coords).sum()),
                             "coords_shape": list(np.asarray(coords).shape),
                             "path": os.path.relpath(full_path, root)})
    try:
        ds[len(ds)]
        out["index_error"] = False
    except IndexError:
        out["index_error"] = True
    return out


def describe(dataset, splits, root: str) -> dict:
    out = {"num_classes": int(dataset.num_classes), "len": int(len(dataset)),
           "slide_ids": [str(s) for s in dataset.slide_data["slide_id"]],
           "labels": [int(v) for v in dataset.slide_data["label"]],
           "slide_cls_ids": [[int(i) for i in a] for a in dataset.slide_cls_ids],
           "patient_cls_ids": [[int(i) for i in a] for a in dataset.patient_cls_ids],
           "patient_case_ids": [str(s) for s in dataset.patient_data["case_id"]],
           "patient_labels": [int(v) for v in dataset.patient_data["label"]]}
    for name, sp in zip(("train", "val", "test"), splits):
        out[name] = None if sp is None else describe_split(sp, root)
    return out


def make(module, root: str, repeat_num=9):
    """main_moc.py:268-289 against ``module`` (the reference's datasets.dataset_generic or moc_b200.datasets)."""
    paths = build(root)
    dataset = module.Generic_MIL_Dataset(csv_path=paths["csv"], data_dir=paths["data_dir"], shuffle=False, seed=1,
                                         print_info=False, label_dict=LABEL_DICT, patient_strat=False, ignore=["OTHER"])
    dataset.load_from_h5(True)
    dataset.load_full_path(True)
    splits = dataset.return_splits(from_id=False, csv_path=paths["splits"], repeat_num=repeat_num)
    for sp in splits:
        if sp is not None:
            sp.load_full_path(True)
            sp.load_from_h5(True)
    return dataset, splits
