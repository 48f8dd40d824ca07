"""Native HDF5 bag reader (moc_b200/csrc/h5_reader.cu, host code): a file libhdf5 itself wrote, and CLAM-shaped
chunked files produced by tests/h5_writer.py.  No GPU needed: the reader launches no kernels."""
import glob
import os

import numpy as np
import pytest
import torch

from moc_b200 import _lib
from moc_b200.bag_store import RaggedBagStore
from moc_b200.h5bag import H5File
from tests.h5_writer import write_h5


def _scipy_hdf5_file():
    try:
        import scipy.io.matlab.tests as t
    except Exception:
        return None
    hits = glob.glob(os.path.join(os.path.dirname(t.__file__), "data", "testhdf5_7.4_GLNX86.mat"))
    return hits[0] if hits else None


def test_reads_a_file_written_by_libhdf5():
    """scipy ships one MATLAB-7.3 file = HDF5 written by libhdf5 (512-byte user block, superblock v0, symbol-table
    root group, contiguous float64 dataset).  scipy's own test states its content: 0 : pi/4 : 2*pi."""
    path = _scipy_hdf5_file()
    if path is None:
        pytest.skip("scipy's MATLAB test data is not installed")
    with H5File(path) as f:
        d = f["testdouble"]
        assert d.shape == (9, 1) and d.dtype == np.float64
        np.testing.assert_array_equal(d[:].ravel(), np.arange(0, np.pi * (2 + 0.25), np.pi / 4)[:9])
        with pytest.raises(_lib.MocError):
            f["no_such_dataset"]


def _bag(n, seed=0, dtype=np.float32):
    g = np.random.default_rng(seed)
    feats = g.standard_normal((n, 512)).astype(dtype)
    coords = g.integers(0, 100000, size=(n, 2)).astype(np.int64)
    return feats, coords


@pytest.mark.parametrize("n", [1, 7, 64, 65, 5000])
def test_clam_shaped_chunked_file(tmp_path, n):
    """Chunk shape (1,512)/(1,2), resizable, interleaved batch appends; 5000 rows need a two-level chunk B-tree."""
    feats, coords = _bag(n, seed=n)
    p = str(tmp_path / "s.h5")
    write_h5(p, {"features": feats, "coords": coords}, layout="chunked", batch=512)
    with H5File(p) as f:
        assert f["features"].shape == (n, 512) and f["features"].dtype == np.float32
        assert f["coords"].shape == (n, 2) and f["coords"].dtype == np.int64
        np.testing.assert_array_equal(f["features"][:], feats)
        np.testing.assert_array_equal(f["coords"][:], coords)


def test_three_level_chunk_tree_and_wide_chunks(tmp_path):
    feats, coords = _bag(3000, seed=3)
    p = str(tmp_path / "s.h5")
    write_h5(p, {"features": feats, "coords": coords}, chunk_fanout=8, batch=256)      # 3000 -> 375 -> 47 -> 6 -> 1
    with H5File(p) as f:
        np.testing.assert_array_equal(f["features"][:], feats)
    write_h5(p, {"features": feats, "coords": coords}, chunk_rows=128, batch=512)       # edge chunk sticks out
    with H5File(p) as f:
        np.testing.assert_array_equal(f["features"][:], feats)
        np.testing.assert_array_equal(f["coords"][:], coords)


@pytest.mark.parametrize("layout", ["contiguous", "compact"])
def test_other_layouts_user_block_and_continuations(tmp_path, layout):
    feats, coords = _bag(20, seed=5)
    p = str(tmp_path / "s.h5")
    write_h5(p, {"features": feats, "coords": coords}, layout=layout, user_block=1024, split_headers=True)
    with H5File(p) as f:
        np.testing.assert_array_equal(f["features"][:], feats)
        np.testing.assert_array_equal(f["coords"][:], coords)
    write_h5(p, {"features": feats, "coords": coords}, layout=layout, superblock_version=1)
    with H5File(p) as f:
        np.testing.assert_array_equal(f["features"][:], feats)


def test_unallocated_chunks_read_as_fill_value(tmp_path):
    feats, coords = _bag(100, seed=6)
    p = str(tmp_path / "s.h5")
    write_h5(p, {"features": feats, "coords": coords}, missing_chunks={"features": [3, 50]})
    want = feats.copy()
    want[[3, 50]] = 0
    with H5File(p) as f:
        np.testing.assert_array_equal(f["features"][:], want)


def test_empty_and_other_dtypes(tmp_path):
    p = str(tmp_path / "s.h5")
    write_h5(p, {"features": np.zeros((0, 512), np.float32), "coords": np.zeros((0, 2), np.int64)})
    with H5File(p) as f:
        assert f["features"].shape == (0, 512) and f["features"][:].shape == (0, 512)
    f16, coords = _bag(33, seed=8, dtype=np.float16)
    write_h5(p, {"features": f16, "coords": coords.astype(np.int32)})
    with H5File(p) as f:
        assert f["features"].dtype == np.float16 and f["coords"].dtype == np.int32
        np.testing.assert_array_equal(f["features"][:], f16)


def test_malformed_files_fail_loudly(tmp_path):
    p = str(tmp_path / "bad.h5")
    with open(p, "wb") as fh:
        fh.write(b"\0" * 4096)
    with pytest.raises(_lib.MocError, match="signature"):
        H5File(p)
    with pytest.raises(_lib.MocError, match="cannot open"):
        H5File(str(tmp_path / "missing.h5"))
    feats, coords = _bag(300, seed=9)
    good = str(tmp_path / "good.h5")
    write_h5(good, {"features": feats, "coords": coords})
    blob = open(good, "rb").read()
    with open(p, "wb") as fh:       # truncated in the middle of the raw data: every address is bounds-checked
        fh.write(blob[:len(blob) // 3])
    with pytest.raises(_lib.MocError):
        with H5File(p) as f:
            f["features"][:]
    with H5File(good) as f:         # destination too small
        d = f["features"]
        buf = np.empty(10, np.float32)
        with pytest.raises(_lib.MocError):
            d.read_into(buf.ctypes.data, buf.nbytes)


def test_cyclic_chunk_tree_fails_instead_of_spinning(tmp_path):
    """A corrupt file whose chunk B-tree points back at an inner node must be rejected, not walked forever."""
    import struct
    feats, coords = _bag(600, seed=11)
    p = str(tmp_path / "cyc.h5")
    write_h5(p, {"features": feats, "coords": coords}, chunk_fanout=8, batch=256)
    blob = bytearray(open(p, "rb").read())
    key_size = 8 + 8 * 3
    patched = 0
    pos = blob.find(b"TREE")
    while pos >= 0:
        ntype, level, used = struct.unpack_from("<BBH", blob, pos + 4)
        if ntype == 1 and level >= 1 and used >= 2:      # inner chunk node: make its last child the node itself
            child_at = pos + 8 + 16 + (used - 1) * (key_size + 8) + key_size
            struct.pack_into("<Q", blob, child_at, pos)
            patched += 1
        pos = blob.find(b"TREE", pos + 4)
    assert patched > 0
    with open(p, "wb") as fh:
        fh.write(bytes(blob))
    with pytest.raises(_lib.MocError, match="cyclic"):
        with H5File(p) as f:
            f["features"][:]


def test_store_from_h5_dir(tmp_path):
    """The loader surface: h5_files/<slide_id>.h5 -> ragged store (host tensors here; pinned + async on CUDA)."""
    os.makedirs(tmp_path / "h5_files")
    bags, ids = [], ["a", "b", "c"]
    for i, s in enumerate(ids):
        feats, coords = _bag([40, 0, 1300][i], seed=20 + i)
        write_h5(str(tmp_path / "h5_files" / (s + ".h5")), {"features": feats, "coords": coords})
        bags.append((feats, coords))
    store, coords = RaggedBagStore.from_h5_dir(str(tmp_path), ids, [0, 1, 0], device="cpu", return_coords=True)
    assert store.offsets_h == [0, 40, 40, 1340] and store.labels_h == [0, 1, 0] and store.slide_ids == ids
    for i in range(3):
        assert torch.equal(store.bag(i), torch.from_numpy(bags[i][0]))
        np.testing.assert_array_equal(coords[i], bags[i][1])


@pytest.mark.parametrize("workers", [1, 2, 5])
def test_store_from_h5_dir_ring_reuse(tmp_path, workers):
    """More bags than staging slots: the read-ahead ring is reused and every bag still lands in its own rows."""
    os.makedirs(tmp_path / "h5_files")
    ids, want = [], []
    for i in range(17):
        feats, coords = _bag([3, 120, 0, 64, 700][i % 5] + i, seed=100 + i)
        write_h5(str(tmp_path / "h5_files" / ("s%02d.h5" % i)), {"features": feats, "coords": coords})
        ids.append("s%02d" % i)
        want.append(feats)
    store = RaggedBagStore.from_h5_dir(str(tmp_path), ids, list(range(17)), device="cpu", workers=workers)
    assert len(store) == 17 and store.total_rows == sum(w.shape[0] for w in want)
    for i in range(17):
        assert torch.equal(store.bag(i), torch.from_numpy(want[i])), i
