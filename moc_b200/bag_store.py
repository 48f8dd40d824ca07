"""GPU-resident ragged bag store and the dataset surface the MOC loops consume.

The reference re-reads every slide's h5 file, pickles it through a DataLoader worker and copies it to the
GPU on *every* pass (datasets/dataset_generic.py:389-433, main_moc.py:382).  Here a split is loaded once
into one contiguous device buffer ``feat [sum N_i, 512]`` (fp32, row-major) with ``offsets [n+1]``; all
kernels take (feat, offsets) and whole splits are processed per launch.  ``BagDataset`` keeps the slice of
``Generic_Split``'s API that train/evaluation/zs_evaluation touch: ``real_len()``, ``repeat_num``
(get/set), ``len()`` and items ``(features, label, coords, full_path)``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

D = 512


class RaggedBagStore:
    def __init__(self, feat: torch.Tensor, offsets_h: Sequence[int], labels_h: Sequence[int],
                 slide_ids: Optional[Sequence[str]] = None):
        assert feat.dim() == 2 and feat.size(1) == D and feat.dtype == torch.float32 and feat.is_contiguous()
        assert len(offsets_h) == len(labels_h) + 1 and offsets_h[-1] == feat.size(0)
        self.feat = feat
        self.offsets_h = [int(v) for v in offsets_h]
        self.labels_h = [int(v) for v in labels_h]
        self.offsets = torch.tensor(self.offsets_h, dtype=torch.int64, device=feat.device)
        self.labels = torch.tensor(self.labels_h, dtype=torch.int64, device=feat.device)
        self.slide_ids = list(slide_ids) if slide_ids is not None else ["slide_%d" % i for i in range(len(labels_h))]

    # ---- construction -------------------------------------------------------------------------
    @staticmethod
    def from_bags(bags: Sequence[torch.Tensor], labels: Sequence[int], device="cuda",
                  slide_ids: Optional[Sequence[str]] = None) -> "RaggedBagStore":
        """Host bags -> one device buffer, staged through pinned memory in <=256 MB pieces."""
        offs = [0]
        for b in bags:
            assert b.dim() == 2 and b.size(1) == D, "a bag is [N,512]"
            offs.append(offs[-1] + b.size(0))
        feat = torch.empty(offs[-1], D, dtype=torch.float32, device=device)
        use_pin = torch.device(device).type == "cuda"
        stage_rows = 128 * 1024
        stage = [torch.empty(stage_rows, D, dtype=torch.float32, pin_memory=use_pin) for _ in range(2)] if use_pin else None
        ev = [None, None]
        k = 0
        for i, b in enumerate(bags):
            b = b.detach()
            if b.is_cuda or not use_pin:
                feat[offs[i]:offs[i + 1]].copy_(b, non_blocking=True)
                continue
            b = b.float()
            for r0 in range(0, b.size(0), stage_rows):
                r1 = min(r0 + stage_rows, b.size(0))
                s = stage[k % 2]
                if ev[k % 2] is not None:
                    ev[k % 2].synchronize()
                s[:r1 - r0].copy_(b[r0:r1])
                feat[offs[i] + r0:offs[i] + r1].copy_(s[:r1 - r0], non_blocking=True)
                ev[k % 2] = torch.cuda.Event()
                ev[k % 2].record()
                k += 1
        if use_pin:
            torch.cuda.current_stream().synchronize()
        return RaggedBagStore(feat, offs, labels, slide_ids)

    @staticmethod
    def synthetic(sizes: Sequence[int], n_classes: int, w_ext: torch.Tensor, cohort_seed: int = 0, device="cuda",
                  labels: Optional[Sequence[int]] = None) -> "RaggedBagStore":
        """Synthetic cohort generated directly into the store on ``device`` (moc_b200.synthetic recipe)."""
        from . import synthetic as syn
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + int(n))
        feat = torch.empty(offs[-1], D, dtype=torch.float32, device=device)
        labs = []
        w_ext = w_ext.to(device)
        for i, n in enumerate(sizes):
            y = int(labels[i]) if labels is not None else i % n_classes
            syn.make_bag(int(n), y, w_ext, n_classes, syn.slide_seed(cohort_seed, i), device=device,
                         out=feat[offs[i]:offs[i + 1]])
            labs.append(y)
        return RaggedBagStore(feat, offs, labs)

    @staticmethod
    def from_pt_dir(data_dir: str, slide_ids: Sequence[str], labels: Sequence[int], device="cuda") -> "RaggedBagStore":
        """CLAM-style ``pt_files/<slide_id>.pt`` bags (datasets/dataset_generic.py:409-410)."""
        import os
        bags = [torch.load(os.path.join(data_dir, "pt_files", "%s.pt" % s), map_location="cpu") for s in slide_ids]
        return RaggedBagStore.from_bags(bags, labels, device, slide_ids)

    @staticmethod
    def from_h5_dir(data_dir: str, slide_ids: Sequence[str], labels: Sequence[int], device="cuda",
                    return_coords: bool = False, workers: int = 8):
        """CLAM-style ``h5_files/<slide_id>.h5`` bags (datasets/dataset_generic.py:424-430), parsed by the native
        reader (moc_b200/h5bag.py; no h5py).  ``workers`` threads read files ahead into a ring of pinned staging
        buffers (the reader is C code called through ctypes, so the GIL is released) while the main thread issues the
        host-to-device copies in slide order.  ``return_coords`` also returns the per-slide ``coords`` (host, numpy)."""
        import os
        from concurrent.futures import ThreadPoolExecutor
        from .h5bag import H5File
        paths = [os.path.join(data_dir, "h5_files", "%s.h5" % s) for s in slide_ids]
        files = [H5File(p) for p in paths]
        try:
            sets = [f["features"] for f in files]
            offs = [0]
            for p, d in zip(paths, sets):
                if len(d.shape) != 2 or d.shape[1] != D:
                    raise ValueError("%s: 'features' has shape %s, expected [N,%d]" % (p, d.shape, D))
                offs.append(offs[-1] + d.shape[0])
            n = len(sets)
            feat = torch.empty(offs[-1], D, dtype=torch.float32, device=device)
            use_pin = torch.device(device).type == "cuda"
            max_rows = max([d.shape[0] for d in sets] + [1])
            workers = max(1, min(int(workers), n))
            ring = min(n, workers + 2) if n else 1
            stage = [torch.empty(max_rows, D, dtype=torch.float32, pin_memory=use_pin) for _ in range(ring)]
            copied = [None] * ring           # event after the H2D copy that last read the slot

            def read(i, k):
                d = sets[i]
                rows = d.shape[0]
                if rows == 0:
                    return
                if d.dtype == np.float32:
                    d.read_into(stage[k].data_ptr(), rows * D * 4)
                else:       # float16 / float64 feature files: convert on the host
                    stage[k][:rows].copy_(torch.from_numpy(d[:].astype(np.float32)))

            with ThreadPoolExecutor(max_workers=workers) as pool:
                pending = {}
                nxt = 0
                for i in range(n):
                    while nxt < n and nxt < i + ring:
                        k = nxt % ring
                        if copied[k] is not None:
                            copied[k].synchronize()      # the slot's previous bag has left for the device
                            copied[k] = None
                        pending[nxt] = pool.submit(read, nxt, k)
                        nxt += 1
                    pending.pop(i).result()
                    k, rows = i % ring, sets[i].shape[0]
                    if rows:
                        feat[offs[i]:offs[i + 1]].copy_(stage[k][:rows], non_blocking=use_pin)
                        if use_pin:
                            copied[k] = torch.cuda.Event()
                            copied[k].record()
            if use_pin:
                torch.cuda.current_stream().synchronize()
            store = RaggedBagStore(feat, offs, labels, slide_ids)
            if return_coords:
                return store, [f["coords"][:] for f in files]
            return store
        finally:
            for f in files:
                f.close()

    # ---- access --------------------------------------------------------------------------------
    def __len__(self) -> int:
        return len(self.labels_h)

    @property
    def device(self):
        return self.feat.device

    @property
    def total_rows(self) -> int:
        return self.offsets_h[-1]

    def n_rows(self, i: int) -> int:
        return self.offsets_h[i + 1] - self.offsets_h[i]

    def bag(self, i: int) -> torch.Tensor:
        return self.feat[self.offsets_h[i]:self.offsets_h[i + 1]]

    def nbytes(self) -> int:
        return self.feat.numel() * 4


class BagDataset:
    """What ``loader.dataset`` looks like to the loops (Generic_Split, dataset_generic.py:380-433,:484-504)."""

    def __init__(self, store: RaggedBagStore, repeat_num: Optional[int] = None):
        self.store = store
        self.repeat_num = repeat_num
        self.num_classes = (max(store.labels_h) + 1) if len(store) else 0
        self.slide_cls_ids = [np.where(np.asarray(store.labels_h) == i)[0] for i in range(self.num_classes)]

    def real_len(self) -> int:
        return len(self.store)

    def __len__(self) -> int:
        return self.repeat_num if self.repeat_num else len(self.store)

    def __getitem__(self, idx: int):
        if self.repeat_num:
            if idx >= self.repeat_num:
                raise IndexError
            idx = idx % len(self.store)
        elif idx >= len(self.store):
            raise IndexError
        n = self.store.n_rows(idx)
        return (self.store.bag(idx), self.store.labels_h[idx], np.zeros((n, 2), dtype=np.int64),
                self.store.slide_ids[idx])


class BagLoader:
    """Stand-in for ``DataLoader(dataset, batch_size=1, shuffle=False)``: yields the same 4-tuples with a
    leading batch dimension, without worker processes or copies (the bags already live on the GPU)."""

    def __init__(self, dataset: BagDataset):
        self.dataset = dataset

    def __len__(self) -> int:
        return len(self.dataset)

    def __iter__(self):
        d = self.dataset
        for k in range(len(d)):
            feats, lbl, coords, path = d[k]
            yield feats.unsqueeze(0), torch.tensor([lbl], device=feats.device), torch.from_numpy(coords).unsqueeze(0), (path,)


class HostChunk:
    def __init__(self, feat: torch.Tensor, offsets_h: Sequence[int], labels_h: Sequence[int], device):
        self.feat = feat                      # pinned [rows,512]
        self.rows = feat.size(0)
        self.offsets_h = [int(v) for v in offsets_h]
        self.labels_h = [int(v) for v in labels_h]
        self.offsets = torch.tensor(self.offsets_h, dtype=torch.int64, device=device)


class HostBags:
    """Bags held in pinned host memory in chunks of whole slides, plus the two device staging buffers the
    engine streams them through (MocEngine.eval_logits_host).  ``chunks`` may repeat the same pinned chunk
    object to model a cohort larger than host RAM should hold."""

    def __init__(self, chunks: Sequence[HostChunk], device="cuda"):
        self.chunks = list(chunks)
        self.device = torch.device(device)
        self.n_slides = sum(len(c.labels_h) for c in self.chunks)
        self.total_rows = sum(c.rows for c in self.chunks)
        max_rows = max(c.rows for c in self.chunks)
        self.staging = [torch.empty(max_rows, D, dtype=torch.float32, device=device) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=device)
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [None, None]
        self.labels = torch.tensor([v for c in self.chunks for v in c.labels_h], dtype=torch.int64, device=device)

    @staticmethod
    def from_bags(bags: Sequence[torch.Tensor], labels: Sequence[int], slides_per_chunk: int = 64, device="cuda"
                  ) -> "HostBags":
        chunks = []
        for lo in range(0, len(bags), slides_per_chunk):
            part = bags[lo:lo + slides_per_chunk]
            offs = [0]
            for b in part:
                offs.append(offs[-1] + b.size(0))
            pinned = torch.empty(offs[-1], D, dtype=torch.float32, pin_memory=True)
            for i, b in enumerate(part):
                pinned[offs[i]:offs[i + 1]].copy_(b)
            chunks.append(HostChunk(pinned, offs, labels[lo:lo + slides_per_chunk], device))
        return HostBags(chunks, device)

    def h2d_bytes(self) -> int:
        return self.total_rows * D * 4
