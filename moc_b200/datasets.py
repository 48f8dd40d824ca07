"""Slide manifests and few-shot splits for the MOC loops: the part of the reference's ``datasets/dataset_generic.py``
that ``main_moc.py:268-293`` touches, written from its observable behaviour (pinned by
``tests/golden/dataset_splits.json``, which the reference's own module produced) and backed by the native HDF5 reader
and the GPU-resident ragged store.

What the driver does with these classes::

    dataset = Generic_MIL_Dataset(csv_path=..., data_dir=..., shuffle=False, seed=1, print_info=True,
                                  label_dict={...}, patient_strat=False, ignore=[])
    dataset.load_from_h5(True); dataset.load_full_path(True)
    train, val, test = dataset.return_splits(from_id=False, csv_path=splits_csv, repeat_num=shot * n_classes)
    split.load_from_h5(True); split.load_full_path(True); DataLoader(split, batch_size=1, ...)

and what the loops read: ``real_len()``, ``repeat_num`` (get / set), ``len()``, items.  Behaviour kept (reference
line in brackets):

* manifest: every csv column is a string; rows whose label is in ``ignore`` are dropped, the rest are mapped through
  ``label_dict`` and an unmapped label raises ``KeyError`` [dataset_generic.py:117-128]; ``filter_dict`` keeps the rows
  whose column values are listed [:130-138]; per-class row positions ``slide_cls_ids`` / ``patient_cls_ids`` [:82-92];
  one label per ``case_id`` (sorted ids) by ``max`` or majority vote, ties to the smallest label [:94-114];
* splits: a split is the manifest rows (manifest order, re-indexed from 0) whose ``slide_id`` occurs in the split
  file's column; blank padding and unknown ids vanish; an empty column yields ``None``; ``repeat_num`` goes to the
  train split only [:201-215, :233-267];
* virtual length: a truthy ``repeat_num`` is the length of a dataset, any non-``None`` one the length of a split; items
  wrap modulo ``real_len()`` and ``IndexError`` ends iteration [:380-396, :500-504];
* items: ``(features, label)`` from ``pt_files/<slide_id>.pt``, ``(slide_id, label)`` without a data directory, or
  ``(features, label, coords[, full_path])`` from ``h5_files/<slide_id>.h5`` [:398-433]; ``toggle_label_revert`` flips
  binary labels [:372-373, :404-405].

Left out because ``main_moc.py`` never reaches it: split generation, bag sub-sampling (``bag_size``), ``preselect``
dictionaries, the ViLa two-scale variants, id look-ups, ``shuffle=True`` and per-source directory dictionaries.

Addition: ``split.store`` / ``split.to_store(device)`` - the whole split read once into a :class:`RaggedBagStore`,
which :mod:`moc_b200.loops` uses instead of iterating a DataLoader.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from .bag_store import RaggedBagStore

_VOTES = {
    "max": lambda labels: labels.max(),
    "maj": lambda labels: labels.mode().iloc[0],     # Series.mode() is ascending: ties go to the smallest label
}


def _positions_by_class(labels, num_classes: int):
    lab = np.asarray(labels, dtype=np.int64)
    return [np.flatnonzero(lab == c) for c in range(num_classes)]


def _read_manifest(csv_path, label_col, label_dict, ignore, filter_dict):
    import pandas as pd
    table = pd.read_csv(csv_path, dtype=str)
    for column, allowed in (filter_dict or {}).items():
        table = table[table[column].isin(allowed)]
    table = table[~table[label_col].isin(list(ignore))].reset_index(drop=True)
    mapped = table[label_col].map(label_dict)
    if mapped.isna().any():
        raise KeyError(table[label_col][mapped.isna()].iloc[0])
    table["label"] = mapped.astype(np.int64)
    return table


class Generic_WSI_Classification_Dataset(torch.utils.data.Dataset):
    """The slide manifest: ``slide_data`` (DataFrame with ``case_id``, ``slide_id``, integer ``label``),
    ``patient_data`` ({"case_id", "label"} arrays), ``num_classes``, ``slide_cls_ids``, ``patient_cls_ids``."""

    def __init__(self, csv_path="dataset_csv/ccrcc_clean.csv", shuffle=False, seed=7, print_info=True, label_dict={},
                 filter_dict={}, ignore=[], patient_strat=False, label_col=None, patient_voting="max"):
        if shuffle:
            raise NotImplementedError("shuffle=True is not supported; main_moc.py passes shuffle=False")
        if patient_voting not in _VOTES:
            raise NotImplementedError
        self.label_dict, self.seed, self.print_info, self.patient_strat = label_dict, seed, print_info, patient_strat
        self.label_col = label_col or "label"
        self.num_classes = len(set(label_dict.values()))
        self.train_ids = self.val_ids = self.test_ids = None
        self.data_dir = None
        self.slide_data = _read_manifest(csv_path, self.label_col, label_dict, ignore, filter_dict)
        per_case = self.slide_data.groupby("case_id", sort=True)["label"].agg(_VOTES[patient_voting])
        self.patient_data = {"case_id": per_case.index.to_numpy(), "label": per_case.to_numpy()}
        self.slide_cls_ids = _positions_by_class(self.slide_data["label"], self.num_classes)
        self.patient_cls_ids = _positions_by_class(self.patient_data["label"], self.num_classes)
        if print_info:
            self.summarize()

    def summarize(self):
        print("label column: {}".format(self.label_col))
        print("label dictionary: {}".format(self.label_dict))
        print("number of classes: {}".format(self.num_classes))
        print("slide-level counts: ", "\n", self.slide_data["label"].value_counts(sort=False))
        for c, (pat, sl) in enumerate(zip(self.patient_cls_ids, self.slide_cls_ids)):
            print("Patient-LVL; Number of samples registered in class %d: %d" % (c, len(pat)))
            print("Slide-LVL; Number of samples registered in class %d: %d" % (c, len(sl)))

    def __len__(self):
        return len(self.patient_data["case_id"] if self.patient_strat else self.slide_data)

    def __getitem__(self, idx):
        return None

    # ---- splits ----------------------------------------------------------------------------------------------
    def get_split_from_df(self, all_splits, split_key="train", bag_size=None, repeat_num=None):
        wanted = set(all_splits[split_key].dropna())
        if not wanted:
            return None
        rows = self.slide_data[self.slide_data["slide_id"].isin(wanted)].reset_index(drop=True)
        return Generic_Split(rows, data_dir=self.data_dir, num_classes=self.num_classes, repeat_num=repeat_num)

    def return_splits(self, from_id=True, csv_path=None, bag_size=None, repeat_num=None, vila=False):
        if from_id or vila or bag_size:
            raise NotImplementedError("only return_splits(from_id=False, csv_path=..., repeat_num=...) is supported: "
                                      "the reference ships its split files (splits/*_fewshot)")
        assert csv_path
        import pandas as pd
        columns = pd.read_csv(csv_path, dtype=str)
        return (self.get_split_from_df(columns, "train", repeat_num=repeat_num),
                self.get_split_from_df(columns, "val"), self.get_split_from_df(columns, "test"))


class Generic_MIL_Dataset(Generic_WSI_Classification_Dataset):
    """Manifest + feature directory: items are whole bags."""

    def __init__(self, data_dir, bag_size=None, repeat_num=None, label_revert=False, **kwargs):
        if bag_size:
            raise NotImplementedError("bag sub-sampling (bag_size) is not part of the MOC path")
        super().__init__(**kwargs)
        self._attach(data_dir, repeat_num, label_revert)

    def _attach(self, data_dir, repeat_num, label_revert=False):
        if isinstance(data_dir, dict):
            raise NotImplementedError("per-source data_dir dictionaries are not supported")
        self.data_dir, self.repeat_num, self.label_revert = data_dir, repeat_num, label_revert
        self.use_h5 = self.return_full_path = False
        self._store = None

    def toggle_label_revert(self, toggle):
        self.label_revert = toggle

    def load_from_h5(self, toggle):
        self.use_h5 = toggle

    def load_full_path(self, toggle):
        self.return_full_path = toggle

    def real_len(self):
        return len(self.slide_data)

    def __len__(self):
        return self.repeat_num if self.repeat_num else super().__len__()

    def _label_of(self, row: int) -> int:
        label = self.slide_data["label"].iat[row]
        return 1 - label if self.label_revert else label

    def _bag_path(self, slide_id: str) -> str:
        sub, ext = ("h5_files", "h5") if self.use_h5 else ("pt_files", "pt")
        return os.path.join(self.data_dir, sub, "%s.%s" % (slide_id, ext))

    def __getitem__(self, idx):
        n = self.real_len()
        if idx >= (self.repeat_num or n):
            raise IndexError
        row = idx % n if self.repeat_num else idx
        slide_id, label = self.slide_data["slide_id"].iat[row], self._label_of(row)
        if self.use_h5:
            from .h5bag import H5File
            path = self._bag_path(slide_id)
            with H5File(path, "r") as f:
                item = (torch.from_numpy(f["features"][:]), label, f["coords"][:])
            return item + (path,) if self.return_full_path else item
        if not self.data_dir:
            return slide_id, label
        return torch.load(self._bag_path(slide_id)), label

    # ---- addition: the split as a GPU-resident ragged store ----------------------------------------------------
    def bag_sizes(self):
        """Patch count of every slide in manifest order without reading the features (h5: dataset header only)."""
        ids = self.slide_data["slide_id"].astype(str).tolist()
        if self.use_h5:
            from .h5bag import H5File
            sizes = []
            for s in ids:
                with H5File(self._bag_path(s)) as f:
                    sizes.append(int(f["features"].shape[0]))
            return sizes
        return [int(torch.load(self._bag_path(s), map_location="cpu").size(0)) for s in ids]

    def to_store(self, device="cuda", rows=None) -> RaggedBagStore:
        """The slides (all of them, or manifest positions ``rows`` - one rank's shard) in manifest order, read once
        (native HDF5 reader or torch.load) into one buffer on ``device``.  With h5 bags the store's slide ids are the
        ``full_path`` strings the reference's loops see."""
        rows = range(self.real_len()) if rows is None else list(rows)
        ids = [str(self.slide_data["slide_id"].iat[r]) for r in rows]
        labels = [int(self._label_of(r)) for r in rows]
        if not self.use_h5:
            return RaggedBagStore.from_pt_dir(self.data_dir, ids, labels, device)
        st = RaggedBagStore.from_h5_dir(self.data_dir, ids, labels, device)
        st.slide_ids = [self._bag_path(s) for s in ids]
        return st

    @property
    def store(self) -> Optional[RaggedBagStore]:
        if not torch.cuda.is_available():
            return None
        if self._store is None:
            self._store = self.to_store(torch.device("cuda", torch.cuda.current_device()))
        return self._store


class Generic_Split(Generic_MIL_Dataset):
    """A subset of the manifest rows (built by ``return_splits``); any non-None ``repeat_num`` is its length."""

    def __init__(self, slide_data, data_dir=None, num_classes=2, bag_size=None, repeat_num=None):
        if bag_size:
            raise NotImplementedError("bag sub-sampling (bag_size) is not part of the MOC path")
        self.slide_data, self.num_classes = slide_data, num_classes
        self.slide_cls_ids = _positions_by_class(slide_data["label"], num_classes)
        self._attach(data_dir, repeat_num)

    def __len__(self):
        return len(self.slide_data) if self.repeat_num is None else self.repeat_num
