"""Developer tool: cProfile of the host side of the few-shot training loop (which is launch/host-bound)."""
import cProfile, pstats, sys, os, types, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import moc_b200 as M
from moc_b200 import loops, synthetic
dev = torch.device("cuda")
c, n = 2, 20000
w, we = synthetic.prompt_matrices(c, device=dev)
loops.set_prompts(w, we)
args = types.SimpleNamespace(n_classes=c, topj=400, topk=10, discard_classifiers=[], pretrain="conch", ablation_study="none", cache_scores=False, disable_tqdm=True)
tr = M.BagLoader(M.BagDataset(M.RaggedBagStore.synthetic([n] * 32, c, we, cohort_seed=1, device=dev), repeat_num=32))
torch.manual_seed(0)
model = M.senet(512, 4).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
for _ in range(3):
    M.train(model, tr, opt, dev, args)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    M.train(model, tr, opt, dev, args)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
