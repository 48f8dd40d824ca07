"""The dataset classes main_moc.py builds its loaders from (moc_b200/datasets.py) against the reference's own
datasets/dataset_generic.py run unmodified on the same small dataset directory (tests/golden/dataset_splits.json,
written by oracle/make_golden_dataset.py; compared live as well where the reference checkout exists).  CPU only."""
import json
import os

import pytest
import torch

from moc_b200 import datasets as ours
from tests import dataset_fixture as fx

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "dataset_splits.json")


def _same(a, b, path=""):
    if isinstance(a, dict):
        assert isinstance(b, dict) and sorted(a) == sorted(b), path
        for k in a:
            _same(a[k], b[k], path + "/" + k)
    elif isinstance(a, list):
        assert isinstance(b, list) and len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, "%s[%d]" % (path, i))
    elif isinstance(a, float):
        assert abs(a - b) <= 1e-9 * max(1.0, abs(a)), (path, a, b)
    else:
        assert a == b, (path, a, b)


@pytest.mark.parametrize("repeat_num", [9, 4, None])
def test_splits_match_the_reference_loader(tmp_path, repeat_num):
    """Label mapping, the ignore list, dataset-csv order inside a split, NaN padding and unknown ids dropped,
    zero-padded ids kept as strings, repeat_num only on the train split, virtual length and modulo indexing, item
    tuples from the h5 files, IndexError at the end."""
    golden = json.load(open(GOLDEN))["repeat_%s" % repeat_num]
    dataset, splits = fx.make(ours, str(tmp_path), repeat_num=repeat_num)
    got = fx.describe(dataset, splits, str(tmp_path))
    _same(json.loads(json.dumps(got)), golden)
    train = splits[0]
    assert train.slide_data["slide_id"].tolist() == ["007", "S-b", "S-c", "0042", "S-g", "S-i"]   # dataset-csv order


def test_live_against_reference_checkout(tmp_path):
    from oracle import ref_loader
    if not ref_loader.has_checkout():
        pytest.skip("reference checkout not present (GPU box)")
    from oracle.make_golden_dataset import load_reference_dataset_module
    ref = load_reference_dataset_module()
    a = fx.describe(*fx.make(ref, str(tmp_path / "ref"), repeat_num=7), str(tmp_path / "ref"))
    b = fx.describe(*fx.make(ours, str(tmp_path / "ours"), repeat_num=7), str(tmp_path / "ours"))
    _same(json.loads(json.dumps(b)), json.loads(json.dumps(a)))


def test_pt_branch_label_revert_and_store(tmp_path):
    """The pt_files branch (dataset_generic.py:409-420), toggle_label_revert, and the split as a ragged store
    (host tensors here; the loops use the CUDA one)."""
    paths = fx.build(str(tmp_path))
    ds = ours.Generic_MIL_Dataset(csv_path=paths["csv"], data_dir=paths["data_dir"], print_info=False,
                                  label_dict={"KICH": 0, "KIRC": 1, "KIRP": 1, "OTHER": 0})
    assert ds.num_classes == 2 and len(ds) == len(fx.ROWS)
    feats, label = ds[0]
    assert torch.equal(feats, torch.from_numpy(fx.bag_of("007")[0])) and label == 1
    ds.toggle_label_revert(True)
    assert ds[0][1] == 0
    ds.toggle_label_revert(False)
    _, val, _ = ds.return_splits(from_id=False, csv_path=paths["splits"])
    for use_h5 in (True, False):
        val.load_from_h5(use_h5)
        st = val.to_store("cpu")
        assert st.labels_h == [1, 0] and len(st) == 2
        assert torch.equal(st.bag(0), torch.from_numpy(fx.bag_of("S-h")[0]))
        assert torch.equal(st.bag(1), torch.from_numpy(fx.bag_of("S-j")[0]))
    assert val.real_len() == 2 and val.repeat_num is None
    with pytest.raises(KeyError):
        ours.Generic_MIL_Dataset(csv_path=paths["csv"], data_dir=paths["data_dir"], print_info=False, label_dict=fx.LABEL_DICT)
