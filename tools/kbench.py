"""Developer micro-benchmark: per-kernel CUDA-event timings of the hot path on synthetic resident bags."""
import argparse
import sys
import os
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moc_b200 import _lib, ops, synthetic  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slides", type=int, default=200)
    ap.add_argument("--patches", type=int, default=20000)
    ap.add_argument("--classes", type=int, default=2)
    ap.add_argument("--topj", type=int, default=400)
    ap.add_argument("--topk", type=int, default=10)
    a = ap.parse_args()
    dev = "cuda"
    c = a.classes
    w, we = synthetic.prompt_matrices(c, device=dev)
    pr = ops.Prompts.pack(w, we)
    n, s = a.patches, a.slides
    feat = torch.empty(n * s, 512, device=dev)
    for i in range(s):
        synthetic.make_bag(n, i % c, we, c, seed=i, device=dev, out=feat[i * n:(i + 1) * n])
    offs = [i * n for i in range(s + 1)]
    offs_d = torch.tensor(offs, dtype=torch.int64, device=dev)
    keys = ops.alloc_keys(c, n * s, dev)
    gb = feat.numel() * 4 / 1e9
    t, tmin = timeit(lambda: ops.score_keys(feat, pr, out=keys))
    print("score_keys  C=%d rows=%d  %.3f ms (min %.3f)  %.1f GB/s (best %.1f)" % (c, n * s, t, tmin, gb / t * 1e3, gb / tmin * 1e3))
    base_h = ops.selection_layout(offs, c, a.topj)
    base_d = torch.tensor(base_h, dtype=torch.int64, device=dev)
    sel = ops.select_union(keys, offs_d, offs, c, a.topj, sel_base=base_d, sel_base_h=base_h)
    t, tmin = timeit(lambda: ops.select_union(keys, offs_d, offs, c, a.topj, sel_base=base_d, sel_base_h=base_h))
    print("select_union  %.3f ms (min %.3f)   mean S=%.0f cap=%d" % (t, tmin, float(sel.sel_count.float().mean()), base_h[1]))
    g = torch.Generator().manual_seed(0)
    prm = ops.HeadParams((torch.rand(64, 512, generator=g) * 2 - 1).mul(512 ** -0.5).to(dev), torch.zeros(64, device=dev),
                         (torch.rand(4, 64, generator=g) * 2 - 1).mul(0.125).to(dev), torch.zeros(4, device=dev))
    t, tmin = timeit(lambda: ops.head_forward(feat, keys, c, sel, prm, _lib.CLS_ALL, a.topk))
    print("head_forward  %.3f ms (min %.3f)" % (t, tmin))
    tot = timeit(lambda: (ops.score_keys(feat, pr, out=keys),
                          ops.select_union(keys, offs_d, offs, c, a.topj, sel_base=base_d, sel_base_h=base_h),
                          ops.head_forward(feat, keys, c, sel, prm, _lib.CLS_ALL, a.topk)))[0]
    print("whole pass  %.3f ms  -> %.0f slides/s  %.3f Gpatch/s  %.1f GB/s algorithmic" % (tot, s / tot * 1e3, n * s / tot / 1e6, gb / tot * 1e3))


if __name__ == "__main__":
    main()
