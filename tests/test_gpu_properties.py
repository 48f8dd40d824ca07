"""Full-size checks of the hot path through properties that do not need the (slow) CPU oracle: BASELINE.json's bag
sizes - 20 000 patches (cfg2/cfg3), 50 000 patches x 30 classes (cfg4), a 100 000-patch bag (cfg5's upper end).

Scores against a float64 product, exact power-of-two linearity, selection = union of the four top-J sets recomputed
with torch.topk on the key planes (strictly-above-threshold rows must be in, strictly-below rows must be out),
ascending / unique / padded row lists, pooling = mean of the K largest combined scores, permutation invariance,
ragged-batch independence, run-to-run determinism."""
import pytest
import torch

from moc_b200 import _lib, ops, synthetic
from oracle import moc_oracle as O
from tests.helpers import assert_union_set

pytestmark = pytest.mark.gpu
DEV = "cuda"
J, K = 400, 10


def _bags(c, sizes, seed):
    w, we = synthetic.prompt_matrices(c, device=DEV)
    offs = [0]
    for n in sizes:
        offs.append(offs[-1] + n)
    feat = torch.empty(offs[-1], 512, device=DEV)
    for i, n in enumerate(sizes):
        synthetic.make_bag(n, i % c, we, c, seed=seed + i, device=DEV, out=feat[offs[i]:offs[i + 1]])
    return w, we, feat, offs


def _pipeline(feat, offs, w, we, c, prm):
    pr = ops.Prompts.pack(w, we)
    keys = ops.score_keys(feat, pr)
    offs_d = torch.tensor(offs, dtype=torch.int64, device=DEV)
    sel = ops.select_union(keys, offs_d, offs, c, J)
    out = ops.head_forward(feat, keys, c, sel, prm, _lib.CLS_ALL, K)
    return keys, sel, out


def _params(seed=5):
    p = O.SenetParams.init(seed)
    return ops.HeadParams(p.w1.to(DEV), p.b1.to(DEV), p.w2.to(DEV), p.b2.to(DEV))


@pytest.mark.parametrize("c,sizes", [(2, [20000] * 6), (3, [20000, 19999, 20001]), (30, [50000, 50000]), (2, [100000, 1000])])
def test_full_size_pipeline_properties(c, sizes):
    w, we, feat, offs = _bags(c, sizes, seed=4000 + c)
    prm = _params()
    keys, sel, out = _pipeline(feat, offs, w, we, c, prm)
    total = offs[-1]
    assert keys.shape == (ops.num_key_planes(c), total)
    keys_stored = keys
    keys = ops.expand_keys(keys, c)       # the checks below index the full 2C+3-plane layout

    # ---- scores: float64 product of the same inputs; softmax / |top1-top2| / background sum, max from them
    L64 = feat.double() @ w.double()
    Le64 = feat.double() @ we.double()[:, c:]
    assert (keys[:c].t().double() - L64).abs().max().item() < 2e-5
    assert (keys[c:2 * c].t().double() - torch.softmax(L64, dim=1)).abs().max().item() < 2e-5
    top2 = L64.topk(2, dim=1).values
    assert (keys[2 * c].double() - (top2[:, 0] - top2[:, 1])).abs().max().item() < 4e-5
    assert (keys[2 * c + 1].double() - Le64.sum(dim=1)).abs().max().item() < 4e-5
    assert (keys[2 * c + 2].double() - Le64.max(dim=1).values).abs().max().item() < 2e-5

    # ---- exact linearity under a power-of-two scale of the features
    keys2 = ops.expand_keys(ops.score_keys(feat * 2.0, ops.Prompts.pack(w, we)), c)
    if c + 4 <= 8:      # fp32 FMA kernel: scaling by 2 commutes with every rounding
        assert torch.equal(keys2[:c], keys[:c] * 2.0) and torch.equal(keys2[2 * c + 1:], keys[2 * c + 1:] * 2.0)
    else:               # FP16x3 tensor-core kernel: the low halves of small features are subnormal (2^-25 absolute floor)
        assert (keys2[:c] - keys[:c] * 2.0).abs().max().item() < 1e-5

    # ---- selection: ascending, unique, padded; union of the four top-J sets of the key planes
    counts = sel.sel_count.cpu().tolist()
    for i, n in enumerate(sizes):
        lo, cnt = sel.sel_base_h[i], counts[i]
        cap = sel.sel_base_h[i + 1] - lo
        assert 0 < cnt <= cap == min(n, J * (2 * c + 2))
        rows = sel.sel_rows[lo:lo + cap]
        assert bool((rows[cnt:] == -1).all())
        r = rows[:cnt].long()
        assert bool((r[1:] > r[:-1]).all()) and int(r[0]) >= offs[i] and int(r[-1]) < offs[i + 1]
        assert torch.equal(sel.sel_local[lo:lo + cnt].long(), r - offs[i])
        chosen = torch.zeros(n, dtype=torch.bool, device=DEV)
        chosen[r - offs[i]] = True
        kk = keys[:, offs[i]:offs[i + 1]]
        j = min(J, n)
        must = torch.zeros(n, dtype=torch.bool, device=DEV)     # strictly above a rank-J threshold of some criterion
        may = torch.zeros(n, dtype=torch.bool, device=DEV)      # at or above one
        planes = [kk[p] for p in range(2 * c + 1)] + [-kk[2 * c + 1]]   # bottom-J of the background sum
        for v in planes:
            thr = v.topk(j).values[-1]
            must |= v > thr
            may |= v >= thr
        assert bool((chosen | ~must).all()), "a row strictly above a top-J threshold is missing"
        assert bool((may | ~chosen).all()), "a row below every top-J threshold was selected"

    # ---- pooling: mean of the K largest combined scores of the slide's selected rows; finite everywhere
    assert torch.isfinite(out.bag_logits).all()
    for i in range(len(sizes)):
        lo, cnt = sel.sel_base_h[i], counts[i]
        f = out.final[lo:lo + cnt]
        ref = f.topk(min(K, cnt), dim=0).values.double().mean(dim=0)
        assert (out.bag_logits[i].double() - ref).abs().max().item() < 1e-6
        pos = out.pool_pos[i]
        assert int(pos.min()) >= 0 and int(pos.max()) < cnt

    # ---- run-to-run determinism (bit-identical)
    keys_b, sel_b, out_b = _pipeline(feat, offs, w, we, c, prm)
    assert torch.equal(keys_b, keys_stored) and torch.equal(sel_b.sel_rows, sel.sel_rows)
    assert torch.equal(out_b.bag_logits, out.bag_logits) and torch.equal(out_b.pool_pos, out.pool_pos)


def test_full_size_permutation_and_batch_independence():
    """Shuffling the patches of a 20 000-patch bag maps the selected set through the permutation and leaves the bag
    logits unchanged; a slide scored inside a ragged batch gives what it gives alone."""
    c = 2
    w, we, feat, offs = _bags(c, [20000, 20000, 7777], seed=77)
    prm = _params(9)
    keys, sel, out = _pipeline(feat, offs, w, we, c, prm)
    counts = sel.sel_count.cpu().tolist()
    # slide 1 alone
    x1 = feat[offs[1]:offs[2]].contiguous()
    k1, s1, o1 = _pipeline(x1, [0, 20000], w, we, c, prm)
    lo = sel.sel_base_h[1]
    assert int(s1.sel_count[0]) == counts[1]
    assert torch.equal(s1.sel_rows[:counts[1]], sel.sel_rows[lo:lo + counts[1]] - offs[1])
    assert torch.equal(o1.bag_logits[0], out.bag_logits[1])
    # permuted copy of slide 1
    perm = torch.randperm(20000, generator=torch.Generator().manual_seed(3)).to(DEV)
    xp = x1[perm].contiguous()
    _, sp, op = _pipeline(xp, [0, 20000], w, we, c, prm)
    got = set(perm[sp.sel_rows[:int(sp.sel_count[0])].long()].cpu().tolist())
    want = set(s1.sel_rows[:counts[1]].cpu().tolist())
    assert_union_set(got, want, k1.cpu().numpy(), c, J)   # only rank-J ties may move with the permutation
    assert (op.bag_logits - o1.bag_logits).abs().max().item() < 1e-6
