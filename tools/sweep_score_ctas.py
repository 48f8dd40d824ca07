"""Developer tool: HBM read rate of the streaming score kernel against its CTA count, on the visible GPU."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moc_b200 import ops, synthetic  # noqa: E402


def main():
    dev = "cuda"
    c = 2
    w, we = synthetic.prompt_matrices(c, device=dev)
    pr = ops.Prompts.pack(w, we)
    rows = 8_000_000
    feat = torch.randn(rows, 512, device=dev)
    feat /= feat.norm(dim=1, keepdim=True)
    keys = ops.alloc_keys(c, rows, dev)
    out = []
    for cap in [int(v) for v in (sys.argv[1:] or ["148", "140", "136", "132", "128", "124", "116"])]:
        for _ in range(8):
            ops.score_keys(feat, pr, out=keys, max_ctas=cap)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            ops.score_keys(feat, pr, out=keys, max_ctas=cap)
        b.record()
        torch.cuda.synchronize()
        out.append("%d:%.0f" % (cap, rows * 2048 * 10 / a.elapsed_time(b) / 1e6))
    print(torch.cuda.get_device_name(0), os.environ.get("CUDA_VISIBLE_DEVICES", "-"), " ".join(out))


if __name__ == "__main__":
    main()
