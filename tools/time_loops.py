"""Developer timing of the reference-shaped loops (train epoch / evaluation / zs_evaluation) on synthetic cfg2 bags."""
import os
import sys
import time
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import moc_b200 as M  # noqa: E402
from moc_b200 import loops, synthetic  # noqa: E402


def timed(name, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print("%-34s %8.2f ms" % (name, dt * 1e3))
    return dt


def main():
    dev = torch.device("cuda")
    c, n = 2, 20000
    w, we = synthetic.prompt_matrices(c, device=dev)
    loops.set_prompts(w, we)
    for cache in (False, True):
        args = types.SimpleNamespace(n_classes=c, topj=400, topk=10, discard_classifiers=[], pretrain="conch",
                                     ablation_study="none", cache_scores=cache, disable_tqdm=True)
        mk = lambda k, seed, rep=None: M.BagLoader(M.BagDataset(
            M.RaggedBagStore.synthetic([n] * k, c, we, cohort_seed=seed, device=dev), repeat_num=rep))
        tr, va, te = mk(32, 1, 32), mk(100, 2), mk(1000, 3)
        torch.manual_seed(0)
        model = M.senet(512, 4).to(dev)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
        print("cache_scores =", cache)
        args.cuda_graph = False
        t = timed("train epoch (32 steps), eager", lambda: M.train(model, tr, opt, dev, args))
        print("   -> %.3f ms per step" % (t * 1e3 / 32))
        args.cuda_graph = True
        t = timed("train epoch (32 steps), CUDA graph", lambda: M.train(model, tr, opt, dev, args))
        print("   -> %.3f ms per step" % (t * 1e3 / 32))
        timed("evaluation(train, 32 slides)", lambda: M.evaluation(model, tr, dev, args))
        timed("evaluation(val, 100 slides)", lambda: M.evaluation(model, va, dev, args))
        timed("evaluation(test, 1000 slides)", lambda: M.evaluation(model, te, dev, args))
        timed("zs_evaluation(test, 1000 slides)", lambda: M.zs_evaluation(te, dev, args))
        del tr, va, te
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
