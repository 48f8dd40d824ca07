"""Load the UNMODIFIED reference functions from the read-only checkout  --  TEST INFRASTRUCTURE.

Used by ``oracle/make_golden*.py`` and the comparison tests in the build container
(read-only checkout at ``/root/reference``), and by ``bench.py``'s CPU legs on the GPU
box, where the three files :func:`load` needs are found in git-ignored ``oracle/_ref/``
(``oracle/stage_ref.py``).  The two selector/pooling modules are
imported by path, and the functions of ``main_moc.py`` are lifted out of its AST
because the script's module-level body (``main_moc.py:47``, ``:133-293``)
parses the CLI and loads a CONCH checkpoint that is not available offline.
"""
from __future__ import annotations

import ast
import importlib.util
import os
import types

_CHECKOUT = os.environ.get("MOC_REFERENCE_ROOT", "/root/reference")
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # written by oracle/stage_ref.py
# the read-only checkout in the build container; on the GPU box the handful of files staged next to this module
REFERENCE_ROOT = _CHECKOUT if os.path.isfile(os.path.join(_CHECKOUT, "main_moc.py")) else _STAGED

_LIFT = ("senet", "slide_process", "train", "zs_evaluation", "evaluation", "ablation_evaluation")


def available() -> bool:
    """The functions :func:`load` lifts can be loaded (full checkout or the staged subset)."""
    return all(os.path.isfile(os.path.join(REFERENCE_ROOT, rel)) for rel in
               ("main_moc.py", "utils/patch_selection_classifier.py", "utils/patch_selection_classifier_index.py"))


def has_checkout() -> bool:
    """The whole read-only checkout is present (golden generation, loader / heads comparisons)."""
    return REFERENCE_ROOT == _CHECKOUT and os.path.isdir(os.path.join(_CHECKOUT, "datasets"))


def _import_by_path(name: str, rel: str) -> types.ModuleType:
    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load() -> types.SimpleNamespace:
    """Namespace with the reference's selectors, poolers and the lifted main_moc functions.

    ``ns.set_weights(W, W_ext)`` injects the two module globals that
    train/evaluation/zs_evaluation read (``main_moc.py:386,427-428,478,537``).
    """
    import numpy as np
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from sklearn.metrics import roc_auc_score
    from tqdm import tqdm

    idx = _import_by_path("_moc_ref_index", "utils/patch_selection_classifier_index.py")
    pool = _import_by_path("_moc_ref_pool", "utils/patch_selection_classifier.py")

    glb = {
        "torch": torch, "nn": nn, "F": F, "np": np, "tqdm": tqdm, "roc_auc_score": roc_auc_score,
        "index_topj_classifier": idx.index_topj_classifier,
        "index_delta_softmax_classifier": idx.index_delta_softmax_classifier,
        "index_delta_diff_classifier": idx.index_delta_diff_classifier,
        "index_bottomk_irrel_classifier": idx.index_bottomk_irrel_classifier,
        "topj_pooling": pool.topj_pooling,
        "delta_softmax_classifier_pooling": pool.delta_softmax_classifier_pooling,
        "delta_diff_classifier_pooling": pool.delta_diff_classifier_pooling,
        "bottomk_irrel_classifier_pooling": pool.bottomk_irrel_classifier_pooling,
        "zeroshot_weights": None, "zeroshot_weights_ext": None,
    }
    path = os.path.join(REFERENCE_ROOT, "main_moc.py")
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in _LIFT]
    assert {n.name for n in body} == set(_LIFT), "reference main_moc.py changed"
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), glb)

    ns = types.SimpleNamespace(index=idx, pool=pool, **{k: glb[k] for k in _LIFT})

    def set_weights(w, w_ext):
        glb["zeroshot_weights"] = w
        glb["zeroshot_weights_ext"] = w_ext

    ns.set_weights = set_weights
    return ns


class RefDataset:
    """Duck-typed stand-in for ``Generic_Split`` as the reference loops consume it
    through ``DataLoader(bs=1)``: items ``(feats[1,N,512], lbl[1], coords[1,N,2], (path,))``."""

    def __init__(self, bags, labels, repeat_num=None):
        self.bags, self.labels, self.repeat_num = bags, labels, repeat_num

    def real_len(self):
        return len(self.bags)

    def __len__(self):
        return self.repeat_num if self.repeat_num else len(self.bags)


class RefLoader:
    def __init__(self, dataset: RefDataset):
        self.dataset = dataset

    def __len__(self):
        return len(self.dataset)

    def __iter__(self):
        import torch
        d = self.dataset
        for k in range(len(d)):
            i = k % len(d.bags)
            x = d.bags[i]
            yield (x.unsqueeze(0), torch.tensor([d.labels[i]]), torch.zeros(1, x.size(0), 2, dtype=torch.int64),
                   ("slide_%d" % i,))
