"""Developer micro-benchmark (lives under tests/ because it times the CPU oracle beside the kernels): forward time of the secondary MIL heads on one resident bag, with the oracle's
CPU time beside it (torch fp32, all host cores)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import moc_b200  # noqa: E402
from moc_b200 import ops, synthetic  # noqa: E402
from oracle import moc_oracle_heads as H  # noqa: E402


def gpu_ms(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def cpu_ms(fn, iters=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t0) / iters * 1e3


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    dev = "cuda"
    torch.set_num_threads(os.cpu_count() or 1)
    w, we = synthetic.prompt_matrices(2)
    x = synthetic.make_bag(n, 0, we, 2, seed=1)
    xd = x.to(dev)
    for m_in, m_out, act in ((512, 512, "relu"), (512, 768, "tanh"), (512, 128, "relu"), (128, 512, "relu")):
        xx = torch.randn(n, m_in, device=dev)
        ww = torch.randn(m_out, m_in, device=dev) * m_in ** -0.5
        t = gpu_ms(lambda: ops.linear(xx, ww, None, act))
        print("linear %4d -> %4d  N=%d  %.3f ms  %.1f TFLOP/s (useful fp32-accurate)" % (m_in, m_out, n, t, 2.0 * n * m_in * m_out / t / 1e9))
    ada = moc_b200.Conch_CLIP_Ada(512, 4, 2, w.to(dev), 0.1, 10).to(dev)
    sd = {k: v.detach().cpu() for k, v in ada.state_dict().items()}
    print("Conch_CLIP_Ada.forward       N=%d  gpu %.3f ms   cpu-oracle %.1f ms" % (
        n, gpu_ms(lambda: ada.forward(xd)), cpu_ms(lambda: H.clip_ada_forward(sd, w, x, 0.1, 10))))
    print("Conch_CLIP_Ada.disable_ada   N=%d  gpu %.3f ms   cpu-oracle %.1f ms" % (
        n, gpu_ms(lambda: ada.forward_disable_ada(xd)), cpu_ms(lambda: H.clip_ada_forward_disable_ada(w, x, 10))))
    ab = moc_b200.CLAM_SB(size_arg="conch", n_classes=2).eval()
    sd = {k: v.detach().clone() for k, v in ab.state_dict().items()}
    ab = ab.to(dev)
    print("ABMIL (CLAM_SB conch)        N=%d  gpu %.3f ms   cpu-oracle %.1f ms" % (
        n, gpu_ms(lambda: ab(xd)), cpu_ms(lambda: H.abmil_forward(sd, x))))
    # one ABMIL training step as utils/core_utils.py:391-416 runs it: forward, CE, backward (ours), against the same
    # step through torch autograd on the CPU oracle
    abt = moc_b200.CLAM_SB(size_arg="conch", n_classes=2).to(dev).train()
    lab = torch.tensor([1], device=dev)

    def step():
        for p_ in abt.parameters():
            p_.grad = None
        logits = abt(xd)[0]
        torch.nn.functional.cross_entropy(logits, lab).backward()
    sdt = {k: v.detach().cpu().clone() for k, v in abt.state_dict().items()}
    print("ABMIL fwd + CE + bwd         N=%d  gpu %.3f ms   cpu-oracle %.1f ms" % (
        n, gpu_ms(step), cpu_ms(lambda: H.abmil_loss_and_grads(sdt, x, 1))))
    for m_out, k_in in ((768, 512), (512, 512)):
        gg = torch.randn(n, m_out, device=dev)
        xx = torch.randn(n, k_in, device=dev)
        t = gpu_ms(lambda: ops.linear_wgrad(gg, xx))
        print("wgrad  dW[%d,%d] = G^T X  N=%d  %.3f ms  %.1f TFLOP/s (useful fp32-accurate)" % (m_out, k_in, n, t, 2.0 * n * m_out * k_in / t / 1e9))
    mf = moc_b200.MIL_fc().eval()
    sd = {k: v.detach().clone() for k, v in mf.state_dict().items()}
    mf = mf.to(dev)
    x384 = torch.randn(n, 384)
    x384d = x384.to(dev)
    print("MIL_fc                       N=%d  gpu %.3f ms   cpu-oracle %.1f ms" % (
        n, gpu_ms(lambda: mf(x384d)), cpu_ms(lambda: H.mil_fc_forward(sd, x384))))


if __name__ == "__main__":
    main()
