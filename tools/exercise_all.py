"""Small end-to-end exercise of every kernel family on one GPU (compute-sanitizer is closed on this pool, so
memory safety is covered by the ragged-size / tail cases of the parity tests instead)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G  # noqa: E402
import moc_b200  # noqa: E402
from moc_b200 import _lib, ops, synthetic  # noqa: E402


def main():
    G.smoke()  # regw score, masked select, head (tcgen05), CE, backward, Adam at C=2
    dev = "cuda"
    for c in (3, 30):  # unmasked select; tensor-core scoring at C=30
        w, we = synthetic.prompt_matrices(c, device=dev)
        pr = ops.Prompts.pack(w, we)
        sizes = [700, 129, 2050]
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + n)
        feat = torch.cat([synthetic.make_bag(n, i % c, we, c, seed=i, device=dev) for i, n in enumerate(sizes)])
        keys = ops.score_keys(feat, pr)
        sel = ops.select_union(keys, torch.tensor(offs, device=dev), offs, c, 100)
        g = torch.Generator().manual_seed(0)
        prm = ops.HeadParams(*(t.to(dev) for t in ((torch.rand(64, 512, generator=g) - 0.5) * 0.1, torch.zeros(64),
                                                    (torch.rand(4, 64, generator=g) - 0.5) * 0.2, torch.zeros(4))))
        out = ops.head_forward(feat, keys, c, sel, prm, _lib.CLS_ALL, 10, want_gate=True)
        ops.pool_topk(keys, torch.tensor(offs, device=dev), len(sizes), c, 10, 0, 1, 0, 1, False)
        ops.topj_sorted(keys[:c, :700].t().contiguous(), 50)
        pr.check_finite()
        torch.cuda.synchronize()
        print("C=%d ok" % c, out.bag_logits[0, :3].tolist())
    x = torch.randn(300, 512, device=dev)
    ab = moc_b200.CLAM_SB(size_arg="conch", n_classes=2).to(dev).eval()
    print("abmil", ab(x)[0].tolist())
    ada = moc_b200.Conch_CLIP_Ada(512, 4, 2, synthetic.prompt_matrices(2, device=dev)[0], 0.1, 10).to(dev)
    print("clip_ada", ada(x).tolist(), ada.forward_disable_ada(x).tolist())
    print("bank", ops.collapse_prompt_bank(torch.randn(10, 512, device=dev), [4, 6]).shape)
    torch.cuda.synchronize()
    print("SANITIZE_TARGET_OK")


if __name__ == "__main__":
    main()
