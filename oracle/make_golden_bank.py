"""Golden vectors for the prompt-bank collapse from the reference's own zero_shot_classifier  --  TEST INFRASTRUCTURE.

    python oracle/make_golden_bank.py          (build container only: needs /root/reference)

``utils/zeroshot_utils.py`` cannot be imported (it pulls in the CONCH model code and timm), so
``zero_shot_classifier`` (:20-51) is lifted out of its AST unmodified and run with a stand-in text tower: a
"tokenizer" that maps every filled-in template to a row number and a "model" whose encode_text returns those rows
of a random bank.  Inputs (bank, prompts per class) and the output matrix go to tests/golden/bank_*.npz.
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def lift_zero_shot_classifier():
    path = os.path.join(ref_loader.REFERENCE_ROOT, "utils", "zeroshot_utils.py")
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "zero_shot_classifier"]
    assert len(body) == 1, "reference zeroshot_utils.py changed"
    rows = {}

    def tokenize(tokenizer, texts):
        return torch.tensor([rows.setdefault(t, len(rows)) for t in texts], dtype=torch.int64)

    glb = {"torch": torch, "F": F, "tokenize": tokenize, "get_tokenizer": lambda: None}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), glb)
    return glb["zero_shot_classifier"], rows


def case(name, names_per_class, n_templates, seed, scale):
    fn, rows = lift_zero_shot_classifier()
    classnames = [["class%d_name%d" % (c, i) for i in range(n)] for c, n in enumerate(names_per_class)]
    templates = ["template %d of CLASSNAME." % t for t in range(n_templates)]
    n_prompts = sum(names_per_class) * n_templates
    g = torch.Generator().manual_seed(seed)
    # un-normalised text embeddings with a common component, so the mean does not cancel; fp16-exact values
    bank = ((torch.randn(n_prompts, 512, generator=g) + 0.7 * torch.randn(1, 512, generator=g)) * scale).half().float()

    class Tower:
        def encode_text(self, token_ids):
            return bank[token_ids]

    w = fn(Tower(), classnames, templates, tokenizer=object(), device="cpu")
    assert len(rows) == n_prompts and list(rows.values()) == list(range(n_prompts))  # class after class, in order
    np.savez_compressed(os.path.join(OUT, name + ".npz"), bank=bank.half().numpy(),
                        prompts_per_class=np.asarray([n * n_templates for n in names_per_class], dtype=np.int64),
                        W=w.numpy())
    print("wrote", name, tuple(w.shape))


if __name__ == "__main__":
    assert ref_loader.has_checkout()
    torch.set_num_threads(1)
    case("bank_rcc_ext", [3, 2, 4, 1, 1, 1, 1], 22, seed=21, scale=1.0)      # RCC-shaped: 3 classes + 4 background
    case("bank_stress", [3, 3, 3], 22, seed=22, scale=0.05)                   # >= 64 prompts per class
