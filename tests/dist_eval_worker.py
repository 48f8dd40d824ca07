"""torchrun worker: slide-sharded evaluation over NCCL must reproduce the single-GPU metrics exactly."""
import json
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import moc_b200 as M  # noqa: E402
from moc_b200 import loops, synthetic  # noqa: E402
from moc_b200.dist import Shard, init_from_env  # noqa: E402


def _solo_dp(M, model, loader, opt, dev, args, masks):
    """The same micro-batches accumulated by this process alone (a one-rank group for the collective)."""
    import torch.distributed as dist
    me = dist.get_rank()
    groups = [dist.new_group(ranks=[r]) for r in range(dist.get_world_size())]   # every rank must create every group
    return M.train(model, loader, opt, dev, args, masks=masks, dp_microbatch=4, group=groups[me])


def main():
    rank, local, world = init_from_env("nccl")
    dev = torch.device("cuda", local)
    c, j, k = 3, 200, 10
    w, we = synthetic.prompt_matrices(c, device=dev)
    loops.set_prompts(w, we)
    sizes = synthetic.log_uniform_sizes(37, lo=300, hi=6000, seed=3)
    labels = [i % c for i in range(len(sizes))]
    args = types.SimpleNamespace(n_classes=c, topj=j, topk=k, discard_classifiers=[], pretrain="conch",
                                 ablation_study="none", cache_scores=False)
    torch.manual_seed(5)
    model = M.senet(512, 4).to(dev)

    def store_for(ids):
        st = M.RaggedBagStore.synthetic([sizes[i] for i in ids], c, we, device=dev, labels=[labels[i] for i in ids])
        for kk, i in enumerate(ids):
            synthetic.make_bag(sizes[i], labels[i], we, c, synthetic.slide_seed(77, i), device=dev, out=st.bag(kk))
        return st

    full = M.BagLoader(M.BagDataset(store_for(list(range(len(sizes))))))
    ref_eval = M.evaluation(model, full, dev, args)
    ref_zs = M.zs_evaluation(full, dev, args)

    sh = Shard(sizes, rank, world)
    ds = M.BagDataset(store_for(sh.ids))
    ds.shard = sh
    part = M.BagLoader(ds)
    got_eval = M.evaluation(model, part, dev, args)
    got_zs = M.zs_evaluation(part, dev, args)
    assert got_eval == ref_eval, (rank, got_eval, ref_eval)
    assert got_zs == ref_zs, (rank, got_zs, ref_zs)
    import torch.distributed as dist
    # unseeded multi-GPU run (the reference never seeds): sync_seed + broadcast_parameters keep the replicas identical
    # through a training epoch whose half masks come from each rank's own CPU generator
    from moc_b200.dist import broadcast_parameters, sync_seed
    torch.rand(rank * 5 + 1)                       # let the generators drift apart first
    sync_seed(None)
    model2 = M.senet(512, 4).to(dev)
    broadcast_parameters(model2)
    opt = torch.optim.Adam(model2.parameters(), lr=1e-3, weight_decay=1e-4)
    train = M.BagLoader(M.BagDataset(store_for(list(range(6))), repeat_num=6))
    M.train(model2, train, opt, dev, args)
    flat = torch.cat([p.detach().flatten() for p in model2.parameters()])
    parts = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(parts, flat)
    assert all(torch.equal(parts[0], p) for p in parts), "replicas diverged during unseeded training"
    got2 = M.evaluation(model2, part, dev, args)
    ref2 = M.evaluation(model2, full, dev, args)
    assert got2 == ref2, (rank, got2, ref2)
    # data-parallel training mode: micro-batches of 4 slides spread over the ranks, one all-reduce each; must equal
    # the single-process accumulation of the same slides and leave all ranks with identical parameters
    import copy
    masks = [torch.rand(train.dataset.store.n_rows(k % 6), generator=torch.Generator().manual_seed(100 + k)) > 0.5
             for k in range(6)]
    m_dp, m_one = copy.deepcopy(model2), copy.deepcopy(model2)
    o_dp = torch.optim.Adam(m_dp.parameters(), lr=1e-3, weight_decay=1e-4)
    o_one = torch.optim.Adam(m_one.parameters(), lr=1e-3, weight_decay=1e-4)
    l_dp = M.train(m_dp, train, o_dp, dev, args, masks=masks, dp_microbatch=4)
    l_one = _solo_dp(M, m_one, train, o_one, dev, args, masks)
    flat = torch.cat([p.detach().flatten() for p in m_dp.parameters()])
    parts = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(parts, flat)
    assert all(torch.equal(parts[0], p) for p in parts), "dp mode: replicas differ"
    one = torch.cat([p.detach().flatten() for p in m_one.parameters()])
    assert float((flat - one).abs().max()) < 1e-6, float((flat - one).abs().max())
    assert float((l_dp - l_one).abs().max()) < 1e-6
    dist.barrier()
    if rank == 0:
        print("DIST_OK " + json.dumps({"world": world, "eval": got_eval, "shard_sizes": [len(v) for v in sh.all_ids]}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
