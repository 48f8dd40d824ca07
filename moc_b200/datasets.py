"""The dataset classes main_moc.py builds its loaders from (datasets/dataset_generic.py of the reference), backed by
the native HDF5 reader and the GPU-resident ragged store.

The reference's driver does (main_moc.py:268-293)

    dataset = Generic_MIL_Dataset(csv_path=..., data_dir=..., shuffle=False, seed=1, print_info=True,
                                  label_dict={...}, patient_strat=False, ignore=[])
    dataset.load_from_h5(True); dataset.load_full_path(True)
    train, val, test = dataset.return_splits(from_id=False, csv_path=splits_csv, repeat_num=shot * n_classes)
    ... .load_from_h5(True) / .load_full_path(True) on each split; DataLoader(split, batch_size=1, ...)

and its loops touch ``real_len()``, ``repeat_num``, ``len()`` and the item tuple.  The classes here keep those names,
constructor arguments, return values and quirks (file:line cited at each) for that slice of the API:

* label mapping and the ``ignore`` list (``df_prep``, :117-128), ``filter_dict`` (:130-138), per-class id lists
  (:82-92), patient-level labels with ``max`` / ``maj`` voting (:94-114), ``summarize`` (:147-154);
* ``return_splits(from_id=False, csv_path=...)`` / ``get_split_from_df`` (:201-215, :233-267): a split keeps the rows
  of the **dataset csv** whose slide_id appears in the split column, in dataset-csv order; NaN padding is dropped;
  an empty column gives ``None``; only the train split receives ``bag_size`` / ``repeat_num``;
* ``Generic_Split`` (:484-504) with the virtual length ``repeat_num`` and ``idx % real_len`` indexing (:380-396),
  ``load_from_h5`` / ``load_full_path`` / ``toggle_label_revert`` (:372-378), items (:389-433): ``(features, label)``
  from ``pt_files/<slide_id>.pt``, or ``(features, label, coords[, full_path])`` from ``h5_files/<slide_id>.h5``.

Not built (unused by main_moc.py): split *generation* (``create_splits`` / ``set_splits`` / ``save_split`` - the
reference ships its split files), the ViLa two-scale variants, ``shuffle=True`` (the reference applies
``np.random.shuffle`` to a DataFrame, :71-73), ``preselect`` dictionaries.

Addition: ``split.store`` - the split as one GPU-resident :class:`RaggedBagStore` (read once through the native HDF5
reader or ``torch.load``), which the loops of :mod:`moc_b200.loops` pick up instead of iterating a DataLoader.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from .bag_store import RaggedBagStore


class Generic_WSI_Classification_Dataset(torch.utils.data.Dataset):
    def __init__(self, csv_path="dataset_csv/ccrcc_clean.csv", shuffle=False, seed=7, print_info=True, label_dict={},
                 filter_dict={}, ignore=[], patient_strat=False, label_col=None, patient_voting="max"):
        import pandas as pd
        self.label_dict = label_dict
        self.num_classes = len(set(self.label_dict.values()))
        self.seed = seed
        self.print_info = print_info
        self.patient_strat = patient_strat
        self.train_ids, self.val_ids, self.test_ids = (None, None, None)
        self.data_dir = None
        if not label_col:
            label_col = "label"
        self.label_col = label_col

        slide_data = pd.read_csv(csv_path, dtype=str)
        slide_data = self.filter_df(slide_data, filter_dict)
        slide_data = self.df_prep(slide_data, self.label_dict, ignore, self.label_col)
        if shuffle:
            raise NotImplementedError("shuffle=True is not mirrored (the reference shuffles a DataFrame in place with "
                                      "np.random.shuffle, dataset_generic.py:71-73); main_moc.py passes shuffle=False")
        self.slide_data = slide_data
        self.patient_data_prep(patient_voting)
        self.cls_ids_prep()
        if print_info:
            self.summarize()

    def cls_ids_prep(self):
        """dataset_generic.py:82-92"""
        self.patient_cls_ids = [np.where(self.patient_data["label"] == i)[0] for i in range(self.num_classes)]
        self.slide_cls_ids = [np.where(self.slide_data["label"] == i)[0] for i in range(self.num_classes)]

    def patient_data_prep(self, patient_voting="max"):
        """dataset_generic.py:94-114"""
        patients = np.unique(np.array(self.slide_data["case_id"]))
        patient_labels = []
        for p in patients:
            locations = self.slide_data[self.slide_data["case_id"] == p].index.tolist()
            assert len(locations) > 0
            label = self.slide_data["label"][locations].values
            if patient_voting == "max":
                label = label.max()
            elif patient_voting == "maj":
                vals, counts = np.unique(label.astype(np.int64), return_counts=True)
                label = vals[counts.argmax()]      # scipy.stats.mode: the smallest of the most frequent labels
            else:
                raise NotImplementedError
            patient_labels.append(label)
        self.patient_data = {"case_id": patients, "label": np.array(patient_labels)}

    @staticmethod
    def df_prep(data, label_dict, ignore, label_col):
        """dataset_generic.py:116-128: drop ignored labels, map the remaining label strings through label_dict.  The
        mapped column holds Python ints in an object column, as ``data.at[i, 'label'] = label_dict[key]`` leaves it."""
        if label_col != "label":
            data["label"] = data[label_col].copy()
        mask = data["label"].isin(ignore)
        data = data[~mask]
        data = data.reset_index(drop=True)
        data["label"] = np.array([label_dict[k] for k in data["label"]], dtype=object)   # KeyError on unknown labels
        return data

    def filter_df(self, df, filter_dict={}):
        """dataset_generic.py:130-138"""
        if len(filter_dict) > 0:
            filter_mask = np.full(len(df), True, bool)
            for key, val in filter_dict.items():
                filter_mask = np.logical_and(filter_mask, df[key].isin(val))
            df = df[filter_mask]
        return df

    def __len__(self):
        return len(self.patient_data["case_id"]) if self.patient_strat else len(self.slide_data)

    def summarize(self):
        """dataset_generic.py:147-154"""
        print("label column: {}".format(self.label_col))
        print("label dictionary: {}".format(self.label_dict))
        print("number of classes: {}".format(self.num_classes))
        print("slide-level counts: ", "\n", self.slide_data["label"].value_counts(sort=False))
        for i in range(self.num_classes):
            print("Patient-LVL; Number of samples registered in class %d: %d" % (i, self.patient_cls_ids[i].shape[0]))
            print("Slide-LVL; Number of samples registered in class %d: %d" % (i, self.slide_cls_ids[i].shape[0]))

    def get_split_from_df(self, all_splits, split_key="train", bag_size=None, repeat_num=None):
        """dataset_generic.py:201-215: rows of the dataset csv (in its order) whose slide_id is in the split column."""
        split = all_splits[split_key]
        split = split.dropna().reset_index(drop=True)
        if len(split) > 0:
            mask = self.slide_data["slide_id"].isin(split.tolist())
            df_slice = self.slide_data[mask].reset_index(drop=True)
            return Generic_Split(df_slice, data_dir=self.data_dir, num_classes=self.num_classes, bag_size=bag_size,
                                 repeat_num=repeat_num)
        return None

    def return_splits(self, from_id=True, csv_path=None, bag_size=None, repeat_num=None, vila=False):
        """dataset_generic.py:233-267 with ``from_id=False`` (what main_moc.py:281 calls)."""
        import pandas as pd
        if from_id:
            raise NotImplementedError("return_splits(from_id=True) needs create_splits()/set_splits(), which are not "
                                      "mirrored: the reference ships its split files (splits/*_fewshot)")
        assert csv_path
        all_splits = pd.read_csv(csv_path, dtype=self.slide_data["slide_id"].dtype)
        train_split = self.get_split_from_df(all_splits, "train", bag_size, repeat_num)
        val_split = self.get_split_from_df(all_splits, "val")
        test_split = self.get_split_from_df(all_splits, "test")
        return train_split, val_split, test_split

    def get_list(self, ids):
        return self.slide_data["slide_id"][ids]

    def getlabel(self, ids):
        return self.slide_data["label"][ids]

    def __getitem__(self, idx):
        return None


class Generic_MIL_Dataset(Generic_WSI_Classification_Dataset):
    def __init__(self, data_dir, bag_size=None, repeat_num=None, label_revert=False, **kwargs):
        super().__init__(**kwargs)
        self.data_dir = data_dir
        self.use_h5 = False
        self.return_full_path = False
        self.bag_size = bag_size
        self.repeat_num = repeat_num
        self.use_preselect = None
        self.preselect_dict = None
        self.label_revert = label_revert
        self.selected_index = []

    def toggle_label_revert(self, toggle):
        self.label_revert = toggle

    def load_from_h5(self, toggle):
        self.use_h5 = toggle

    def load_full_path(self, toggle):
        self.return_full_path = toggle

    def __len__(self):
        if self.repeat_num:
            return self.repeat_num
        return super().__len__()

    def real_len(self):
        return len(self.slide_data)

    def _data_dir_of(self, idx):
        if type(self.data_dir) == dict:
            return self.data_dir[self.slide_data["source"][idx]]
        return self.data_dir

    def __getitem__(self, idx):
        """dataset_generic.py:389-433"""
        if self.repeat_num:
            if idx >= self.repeat_num:
                raise IndexError
            idx = idx % len(self.slide_data)
        elif idx >= len(self.slide_data):
            raise IndexError
        slide_id = self.slide_data["slide_id"][idx]
        label = self.slide_data["label"][idx]
        if self.label_revert:
            label = 1 - label
        data_dir = self._data_dir_of(idx)
        if not self.use_h5:
            if self.data_dir:
                full_path = os.path.join(data_dir, "pt_files", "{}.pt".format(slide_id))
                features = torch.load(full_path)
                if self.bag_size and self.selected_index:
                    features = features[self.selected_index[idx]]
                elif self.bag_size and self.preselect is None:
                    features = features[torch.randperm(features.size(0))[:self.bag_size]]
                if self.use_preselect is not None:
                    features = features[self.preselect_dict[slide_id]]
                return features, label
            return slide_id, label
        from .h5bag import H5File
        full_path = os.path.join(data_dir, "h5_files", "{}.h5".format(slide_id))
        with H5File(full_path, "r") as hdf5_file:
            features = hdf5_file["features"][:]
            coords = hdf5_file["coords"][:]
        features = torch.from_numpy(features)
        if self.return_full_path:
            return features, label, coords, full_path
        return features, label, coords

    def get_item_by_id(self, slide_id):
        idx = self.slide_data[self.slide_data["slide_id"] == slide_id].index[0]
        return self.__getitem__(idx)

    preselect = None   # the attribute the reference's bag_size branch tests (:415); its preselect() setter is not mirrored

    # ---- addition: the split as a GPU-resident ragged store ----------------------------------------------------
    def to_store(self, device="cuda") -> RaggedBagStore:
        """Every slide of the split, in split order, read once (native HDF5 reader or torch.load) into one device
        buffer.  ``full_path`` strings are kept as the store's slide ids, as the loops report them."""
        ids = [str(s) for s in self.slide_data["slide_id"]]
        labels = [int(v) for v in self.slide_data["label"]]
        if self.label_revert:
            labels = [1 - v for v in labels]
        if type(self.data_dir) == dict:
            raise NotImplementedError("to_store: per-source data_dir dictionaries are not supported")
        if self.use_h5:
            st = RaggedBagStore.from_h5_dir(self.data_dir, ids, labels, device)
            st.slide_ids = [os.path.join(self.data_dir, "h5_files", "{}.h5".format(s)) for s in ids]
        else:
            st = RaggedBagStore.from_pt_dir(self.data_dir, ids, labels, device)
        return st

    @property
    def store(self) -> Optional[RaggedBagStore]:
        if not torch.cuda.is_available():
            return None
        st = getattr(self, "_store", None)
        if st is None:
            st = self._store = self.to_store(torch.device("cuda", torch.cuda.current_device()))
        return st


class Generic_Split(Generic_MIL_Dataset):
    def __init__(self, slide_data, data_dir=None, num_classes=2, bag_size=None, repeat_num=None):
        """dataset_generic.py:484-498"""
        self.use_h5 = False
        self.return_full_path = False
        self.slide_data = slide_data
        self.data_dir = data_dir
        self.num_classes = num_classes
        self.bag_size = bag_size
        self.slide_cls_ids = [np.where(self.slide_data["label"] == i)[0] for i in range(self.num_classes)]
        self.repeat_num = repeat_num
        self.use_preselect = None
        self.preselect_dict = None
        self.label_revert = False
        self.selected_index = []

    def __len__(self):
        """dataset_generic.py:500-504"""
        if self.repeat_num is not None:
            return self.repeat_num
        return len(self.slide_data)
