"""CUDA kernels (through the C ABI) against the oracle and the reference-generated golden fixtures."""
import numpy as np
import pytest
import torch

from oracle import moc_oracle as O
from tests.helpers import full_keys, assert_topj_set, assert_union_set, close, params_from_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"
SLIDE_CASES = ["slide_c2", "slide_c2_j64", "slide_c3", "slide_c30"]


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def load_slides(g):
    bags = [T(g["s%d_feat" % i]).float() for i in range(int(g["n_slides"]))]
    offs = [0]
    for b in bags:
        offs.append(offs[-1] + b.size(0))
    return bags, offs


def oracle_keys(x, w, we, c):
    k = O.selection_keys(x, w, we, c)
    return np.concatenate([k["logit"].T, k["softmax"].T, k["delta"][None], k["bg_sum"][None], k["bg_max"][None]], 0)


@pytest.mark.parametrize("name", SLIDE_CASES)
def test_score_keys_golden(golden, name):
    from moc_b200 import ops
    g = golden(name)
    c = int(g["C"])
    w, we = T(g["W"]), T(g["W_ext"])
    bags, offs = load_slides(g)
    feat = torch.cat(bags).to(DEV)
    pr = ops.Prompts.pack(w.to(DEV), we.to(DEV))
    keys = ops.score_keys(feat, pr)
    torch.cuda.synchronize()
    assert keys.shape == (ops.num_key_planes(c), offs[-1])
    keys = full_keys(keys, c)
    for i, x in enumerate(bags):
        ref = oracle_keys(x, w, we, c)
        got = keys[:, offs[i]:offs[i + 1]].cpu().numpy()
        close(got, ref)
        close(got[:c].T, g["s%d_L" % i])
        # tighter: fp32 accumulation order is the only difference
        assert np.abs(got[:c].T - g["s%d_L" % i]).max() < 2e-6


@pytest.mark.parametrize("n_rows", [1, 2, 3, 4, 5, 7, 31, 33, 1000, 4099])
@pytest.mark.parametrize("c,n_ext", [(2, 6), (3, 7), (2, 3), (4, 8), (5, 9), (30, 34), (12, 16), (59, 64)])
def test_score_keys_shapes(n_rows, c, n_ext):
    """Ragged tails (rows not a multiple of the 4-row stage), every register-resident width, the shared-memory path."""
    from moc_b200 import ops
    gen = torch.Generator().manual_seed(n_rows * 131 + c)
    x = torch.randn(n_rows, 512, generator=gen)
    x = x / x.norm(dim=1, keepdim=True)
    wa = torch.randn(n_ext, 512, generator=gen)
    wa = wa / wa.norm(dim=1, keepdim=True)
    w, we = wa[:c].t().contiguous(), wa.t().contiguous()
    keys = full_keys(ops.score_keys(x.to(DEV), ops.Prompts.pack(w.to(DEV), we.to(DEV))), c)
    close(keys.cpu().numpy(), oracle_keys(x, w, we, c), rtol=1e-3, atol=2e-6)


@pytest.mark.parametrize("impl", ["tc", "simt"])
@pytest.mark.parametrize("n_rows", [1, 127, 128, 129, 4099, 60001])
@pytest.mark.parametrize("c,n_ext", [(2, 6), (12, 16), (20, 33), (30, 34), (59, 64)])
def test_score_keys_both_impls(monkeypatch, impl, n_rows, c, n_ext):
    """The tensor-core (tcgen05, FP16 x3 split) and CUDA-core scoring kernels compute the same fp32-level keys;
    60001 rows = several 128-patch tiles per CTA, a ragged last tile and the cross-tile load ring."""
    from moc_b200 import ops
    monkeypatch.setattr(ops, "SCORE_IMPL", impl)
    gen = torch.Generator().manual_seed(n_rows * 7 + n_ext)
    x = torch.randn(n_rows, 512, generator=gen)
    x = x / x.norm(dim=1, keepdim=True) * (0.25 + 4 * torch.rand(n_rows, 1, generator=gen))  # norms 0.25 .. 4.25
    wa = torch.randn(n_ext, 512, generator=gen)
    wa = wa / wa.norm(dim=1, keepdim=True)
    w, we = wa[:c].t().contiguous(), wa.t().contiguous()
    pr = ops.Prompts.pack(w.to(DEV), we.to(DEV))
    assert (pr.tc is not None) == (impl == "tc" or n_ext > 8)
    keys = full_keys(ops.score_keys(x.to(DEV), pr), c).cpu().numpy()
    close(keys, oracle_keys(x, w, we, c), rtol=1e-3, atol=2e-6)
    exact = (x.double() @ w.double()).numpy().T  # fp32-level accuracy of the raw similarities
    assert np.abs(keys[:c] - exact).max() < 4e-6
    pr.check_finite()


def test_score_keys_tc_flags_nonfinite():
    """Out-of-range / non-finite features give non-finite scores (as they do in the reference) and raise the flag."""
    from moc_b200 import ops
    from moc_b200._lib import MocError
    gen = torch.Generator().manual_seed(3)
    wa = torch.randn(34, 512, generator=gen)
    wa = wa / wa.norm(dim=1, keepdim=True)
    pr = ops.Prompts.pack(wa[:30].t().contiguous().to(DEV), wa.t().contiguous().to(DEV))
    x = torch.randn(300, 512, generator=gen)
    ops.score_keys(x.to(DEV), pr)
    pr.check_finite()
    x[17, 5] = float("inf")
    keys = ops.score_keys(x.to(DEV), pr)
    assert not torch.isfinite(keys[:30, 17]).all() and torch.isfinite(keys[:, :17]).all()
    with pytest.raises(MocError):
        pr.check_finite()


def test_score_keys_normalize_flag():
    """normalize=1 == scoring F.normalize(x) (models/model_adapters.py:188); off by default as in MOC."""
    from moc_b200 import ops
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(777, 512, generator=gen) * 3.0
    for c, n_ext in [(2, 6), (30, 34)]:
        wa = torch.randn(n_ext, 512, generator=gen)
        wa = wa / wa.norm(dim=1, keepdim=True)
        w, we = wa[:c].t().contiguous(), wa.t().contiguous()
        keys = full_keys(ops.score_keys(x.to(DEV), ops.Prompts.pack(w.to(DEV), we.to(DEV)), normalize=True), c)
        xn = torch.nn.functional.normalize(x, dim=-1)
        close(keys.cpu().numpy(), oracle_keys(xn, w, we, c), rtol=1e-3, atol=2e-6)


@pytest.mark.parametrize("name", SLIDE_CASES)
def test_select_union_golden(golden, name):
    """Batched selection over all slides of the fixture: per-selector index sets and the ascending union."""
    from moc_b200 import _lib, ops
    g = golden(name)
    c, j = int(g["C"]), int(g["J"])
    w, we = T(g["W"]), T(g["W_ext"])
    bags, offs = load_slides(g)
    feat = torch.cat(bags).to(DEV)
    pr = ops.Prompts.pack(w.to(DEV), we.to(DEV))
    keys = ops.score_keys(feat, pr)
    offs_d = torch.tensor(offs, dtype=torch.int64, device=DEV)
    for di, disc in enumerate(g["discards"]):
        disc = tuple(d for d in str(disc).split("|") if d)
        sel = ops.select_union(keys, offs_d, offs, c, j, _lib.discard_bits(disc))
        cnt = sel.sel_count.cpu().tolist()
        rows = sel.sel_rows.cpu().numpy()
        local = sel.sel_local.cpu().numpy()
        for i in range(len(bags)):
            b = sel.sel_base_h[i]
            got = local[b:b + cnt[i]]
            ref = g["s%d_d%d_selected_index" % (i, di)]
            assert (np.diff(got) > 0).all()
            assert (rows[b:b + cnt[i]] == got + offs[i]).all()
            assert (rows[b + cnt[i]:sel.sel_base_h[i + 1]] == -1).all()
            assert_union_set(got, ref, oracle_keys(bags[i], w, we, c), c, j, disc)


def _ref_union(keys, offs, c, j, mask=None):
    """Top-j per selection plane with ties broken towards lower row indices; union per slide, ascending."""
    out = []
    for i in range(len(offs) - 1):
        lo, hi = offs[i], offs[i + 1]
        keep = np.ones(hi - lo, bool) if mask is None else mask[lo:hi].astype(bool)
        idx = np.nonzero(keep)[0]
        jj = min(j, idx.size)
        sel = set()
        for plane in list(range(2 * c + 1)) + [2 * c + 1]:
            v = keys[plane, lo:hi][idx].astype(np.float64)
            order = np.lexsort((idx, v if plane == 2 * c + 1 else -v))
            sel |= set(idx[order[:jj]].tolist())
        out.append(sorted(sel))
    return out


@pytest.mark.parametrize("kind", ["normal", "quantised", "constant", "two_values", "tiny_range"])
@pytest.mark.parametrize("masked", [False, True])
def test_select_union_ties_and_degenerate_columns(kind, masked):
    """Exact set semantics on crafted key planes: heavy ties at the rank-J value, constant columns (every key in one
    histogram bin: the generic fallback), ragged unaligned slides from 1 row to 70 000 rows, with and without the
    training row mask."""
    from moc_b200 import ops
    c, j = 2, 400
    sizes = [1, 5, 399, 400, 401, 1003, 4096, 5001, 70000, 33]
    offs = [0]
    for n in sizes:
        offs.append(offs[-1] + n)
    total = offs[-1]
    rng = np.random.default_rng(hash(kind) % 1000 + masked)
    keys = rng.standard_normal((2 * c + 3, total)).astype(np.float32)
    if kind == "quantised":
        keys = np.round(keys, 1)
    elif kind == "constant":
        keys[:] = 0.25
    elif kind == "two_values":
        keys = np.where(keys > 1.5, np.float32(1.0), np.float32(-1.0)).astype(np.float32)
    elif kind == "tiny_range":
        keys = (1.0 + keys * 1e-6).astype(np.float32)
    mask = (rng.random(total) > 0.5) if masked else None
    kd = torch.from_numpy(keys).to(DEV)
    md = torch.from_numpy(mask).to(DEV) if masked else None
    offs_d = torch.tensor(offs, dtype=torch.int64, device=DEV)
    sel = ops.select_union(kd, offs_d, offs, c, j, 0, md)
    cnt = sel.sel_count.cpu().tolist()
    rows = sel.sel_rows.cpu().numpy()
    local = sel.sel_local.cpu().numpy()
    ref = _ref_union(keys, offs, c, j, mask)
    for i in range(len(sizes)):
        b = sel.sel_base_h[i]
        got = (rows[b:b + cnt[i]] - offs[i]).tolist()
        assert got == ref[i], "slide %d (n=%d): %d vs %d rows" % (i, sizes[i], len(got), len(ref[i]))
        if masked:
            keep_idx = np.nonzero(mask[offs[i]:offs[i + 1]])[0]
            assert local[b:b + cnt[i]].tolist() == np.searchsorted(keep_idx, got).tolist()
        else:
            assert local[b:b + cnt[i]].tolist() == got


@pytest.mark.parametrize("kind", ["normal", "sorted", "clumped", "spiky", "quantised", "few_distinct", "tiny_range"])
@pytest.mark.parametrize("j", [10, 400, 720, 721])
def test_select_union_one_scan_path(kind, j):
    """Long unmasked columns take the sampled one-scan selection (select_rows_sampled: a provisional threshold from a
    1/16 sample, one scan that parks everything above it, resolution on chip) and fall back to the three-scan path when
    the parked count misses [j, 4096] or the threshold bin is crowded.  Whatever path a column takes the result is the
    exact top-j set with ties to the lowest rows: sizes around the path's limits (8192, the 65 536 stride change), key
    distributions that make the sample unrepresentative (sorted, clumped), stretch the histogram range (spiky) or tie
    heavily (quantised, few_distinct).  j = 721 is past the path's limit (three scans)."""
    from moc_b200 import ops
    c = 2
    sizes = [8191, 8192, 8193, 20000, 50001, 65536, 65537, 100003, 150000]
    offs = [0]
    for n in sizes:
        offs.append(offs[-1] + n)
    total = offs[-1]
    rng = np.random.default_rng(sum(map(ord, kind)) * 1000 + j)
    keys = rng.standard_normal((2 * c + 3, total)).astype(np.float32)
    if kind == "sorted":
        for i in range(len(sizes)):
            keys[:, offs[i]:offs[i + 1]] = np.sort(keys[:, offs[i]:offs[i + 1]], axis=1)
            keys[1::2, offs[i]:offs[i + 1]] = keys[1::2, offs[i]:offs[i + 1]][:, ::-1]
    elif kind == "clumped":      # the largest (and smallest) keys sit in a few contiguous runs that straddle sample sectors
        for i in range(len(sizes)):
            lo, n = offs[i], sizes[i]
            for p in range(keys.shape[0]):
                for _ in range(3):
                    a = int(rng.integers(0, n - 300))
                    keys[p, lo + a + 8:lo + a + 8 + 250] += np.float32(4.0) * (1 if p != 2 * c + 1 else -1)
    elif kind == "spiky":
        idx = rng.integers(0, total, size=40)
        keys[:, idx] *= np.float32(1e6)
    elif kind == "quantised":
        keys = np.round(keys, 2)
    elif kind == "few_distinct":
        keys = np.round(keys * 2).astype(np.float32)
    elif kind == "tiny_range":
        keys = (1.0 + keys * 1e-6).astype(np.float32)
    kd = torch.from_numpy(np.ascontiguousarray(keys)).to(DEV)
    offs_d = torch.tensor(offs, dtype=torch.int64, device=DEV)
    sel = ops.select_union(kd, offs_d, offs, c, j, 0, None)
    cnt = sel.sel_count.cpu().tolist()
    rows = sel.sel_rows.cpu().numpy()
    ref = _ref_union(keys, offs, c, j)
    for i in range(len(sizes)):
        b = sel.sel_base_h[i]
        got = (rows[b:b + cnt[i]] - offs[i]).tolist()
        assert got == ref[i], "slide %d (n=%d): %d vs %d rows" % (i, sizes[i], len(got), len(ref[i]))


@pytest.mark.parametrize("name", SLIDE_CASES)
def test_topj_sorted_golden(golden, name):
    """The selectors' public return value: sorted top-J indices per column."""
    from moc_b200 import ops
    g = golden(name)
    c, j = int(g["C"]), int(g["J"])
    for i in range(int(g["n_slides"])):
        lo = T(g["s%d_L" % i])
        idx = ops.topj_sorted(lo.to(DEV), j).cpu().numpy()
        ref = g["s%d_idx_topj" % i]
        assert idx.shape == ref.shape
        for cc in range(c):
            assert_topj_set(idx[:, cc], ref[:, cc], lo[:, cc].numpy(), j)
            v = lo[:, cc].numpy()[idx[:, cc]]
            assert (np.diff(v) <= 0).all()
        bg = T(g["s%d_Le" % i])[:, c:].sum(dim=1)
        idx_s = ops.topj_sorted(bg.to(DEV), j, largest=False).cpu().numpy()
        assert_topj_set(idx_s, g["s%d_idx_bottomk" % i].flatten(), bg.numpy(), j, largest=False)
        assert (np.diff(bg.numpy()[idx_s]) >= 0).all()


def test_topj_sorted_ties_and_sizes():
    from moc_b200 import ops
    gen = torch.Generator().manual_seed(0)
    v = torch.randint(0, 7, (5000, 3), generator=gen).float()  # massive ties
    for j in (1, 10, 400, 4999, 5000, 9000):
        idx = ops.topj_sorted(v.to(DEV), j).cpu()
        jj = min(j, 5000)
        assert idx.shape == (jj, 3)
        for cc in range(3):
            col = v[:, cc]
            got = col[idx[:, cc]]
            ref = col.topk(jj, 0, True, True)[0]
            assert torch.equal(got, ref)
            assert idx[:, cc].unique().numel() == jj
            # ties resolved towards lower row index
            same = got[1:] == got[:-1]
            assert (idx[1:, cc][same] > idx[:-1, cc][same]).all()


@pytest.mark.parametrize("name", SLIDE_CASES)
def test_head_forward_golden(golden, name):
    from moc_b200 import _lib, ops
    g = golden(name)
    c, j, k = int(g["C"]), int(g["J"]), int(g["K"])
    w, we = T(g["W"]), T(g["W_ext"])
    bags, offs = load_slides(g)
    feat = torch.cat(bags).to(DEV)
    keys = ops.score_keys(feat, ops.Prompts.pack(w.to(DEV), we.to(DEV)))
    offs_d = torch.tensor(offs, dtype=torch.int64, device=DEV)
    prm = params_from_golden(g, "sd_", DEV)
    for di, disc in enumerate(g["discards"]):
        disc = tuple(d for d in str(disc).split("|") if d)
        sel = ops.select_union(keys, offs_d, offs, c, j, _lib.discard_bits(disc))
        out = ops.head_forward(feat, keys, c, sel, prm, _lib.active_bits(disc, "eval"), k, want_gate=True)
        cnt = sel.sel_count.cpu().tolist()
        for i in range(len(bags)):
            q = "s%d_d%d_" % (i, di)
            ref_idx = g[q + "selected_index"]
            b = sel.sel_base_h[i]
            got_idx = sel.sel_local[b:b + cnt[i]].cpu().tolist()
            common = assert_union_set(got_idx, ref_idx, oracle_keys(bags[i], w, we, c), c, j, disc)
            # per-row quantities on the rows both selections hold (all of them unless a rank-J tie was swapped)
            gp = torch.tensor([got_idx.index(r) for r in common], device=DEV) + b
            rp = np.asarray([ref_idx.tolist().index(r) for r in common]) if common != ref_idx.tolist() else slice(None)
            close(out.gate[gp], g[q + "gate"][rp], rtol=1e-4, atol=1e-6)
            close(out.final[gp], g[q + "final"][rp])
            # pooling: mean of the K largest combined scores of OUR selection; equal to the golden when the sets agree
            fin = out.final[b:b + cnt[i]].double()
            close(out.bag_logits[i:i + 1], fin.topk(min(k, cnt[i]), dim=0).values.mean(dim=0, keepdim=True))
            if got_idx == ref_idx.tolist():
                close(out.bag_logits[i:i + 1], g[q + "bag_logits"])
            # planes are the key rows of the selected patches
            rows = sel.sel_rows[gp].long()
            fk = full_keys(keys, c)
            close(fk[:c, rows].t(), g[q + "plane_top"][rp])
            close(fk[c:2 * c, rows].t(), g[q + "plane_dsoftmax"][rp])
            close(fk[2 * c, rows], g[q + "plane_ddiff"][rp][:, 0])
            close(fk[2 * c + 2, rows], g[q + "plane_bottomk"][rp][:, 0])


def test_pool_topk_zero_shot(golden):
    from moc_b200 import ops
    for name in SLIDE_CASES:
        g = golden(name)
        c, k = int(g["C"]), int(g["K"])
        w, we = T(g["W"]), T(g["W_ext"])
        bags, offs = load_slides(g)
        feat = torch.cat(bags).to(DEV)
        keys = ops.score_keys(feat, ops.Prompts.pack(w.to(DEV), we.to(DEV)))
        offs_d = torch.tensor(offs, dtype=torch.int64, device=DEV)
        n = len(bags)
        keys = full_keys(keys, c)                                             # plane indices below: the full layout
        zs = ops.pool_topk(keys, offs_d, n, c, k, 0, 1, 0, 1)                 # topj_pooling
        ds = ops.pool_topk(keys, offs_d, n, c, k, c, 1, 0, 1)                 # delta_softmax pooling
        dd = ops.pool_topk(keys, offs_d, n, c, k, 2 * c, 0, 0, 1)             # delta_diff pooling
        bk = ops.pool_topk(keys, offs_d, n, c, k, 2 * c + 1, 0, 0, 1, smallest=True)  # bottomk_irrel pooling
        for i in range(n):
            close(zs[i:i + 1], g["s%d_pool_topj" % i])
            close(ds[i:i + 1], g["s%d_pool_dsoftmax" % i])
            close(dd[i:i + 1], g["s%d_pool_ddiff" % i])
            close(bk[i:i + 1], g["s%d_pool_bottomk" % i])


def _oracle_step(prm, x, w, we, c, j, k, lbl, mask, disc=()):
    slide = O.slide_process(x, w, we, c, j, discard_classifiers=disc, mask=mask)
    return slide, O.head_forward_backward(prm, slide, lbl, k, O.active_classifiers(disc, "train"))


@pytest.mark.parametrize("c,n,j,k,masked", [(2, 900, 100, 10, True), (2, 37, 400, 10, False), (3, 500, 60, 10, True),
                                            (30, 400, 40, 10, False), (2, 6, 400, 10, True)])
def test_backward_matches_oracle(c, n, j, k, masked):
    from moc_b200 import _lib, ops, synthetic
    w, we = synthetic.prompt_matrices(c)
    x = synthetic.make_bag(n, 1, we, c, seed=77 + n)
    gen = torch.Generator().manual_seed(n)
    mask = (torch.rand(n, generator=gen) > 0.5) if masked else None
    oprm = O.SenetParams.init(3)
    slide, (loss, logits, grads) = _oracle_step(oprm, x, w, we, c, j, k, 1, mask)

    feat = x.to(DEV)
    keys = ops.score_keys(feat, ops.Prompts.pack(w.to(DEV), we.to(DEV)))
    offs = [0, n]
    offs_d = torch.tensor(offs, dtype=torch.int64, device=DEV)
    sel = ops.select_union(keys, offs_d, offs, c, j, 0, row_mask=None if mask is None else mask.to(DEV))
    cnt = int(sel.sel_count[0])
    assert sel.sel_local[:cnt].cpu().tolist() == slide["selected_index"]
    prm = ops.HeadParams(oprm.w1.to(DEV), oprm.b1.to(DEV), oprm.w2.to(DEV), oprm.b2.to(DEV))
    out = ops.head_forward(feat, keys, c, sel, prm, _lib.CLS_ALL, k)
    close(out.bag_logits, logits)
    lbl = torch.tensor([1], device=DEV)
    l, dl, pred = ops.cross_entropy(out.bag_logits, lbl, want_grad=True, want_pred=True)
    close(l, loss.reshape(1), rtol=1e-4, atol=1e-6)
    assert int(pred[0]) == int(logits.argmax())
    gflat = ops.head_backward(feat, keys, c, sel, prm, _lib.CLS_ALL, k, out.pool_pos, dl)
    ref = torch.cat([t.flatten() for t in grads])
    scale = float(ref.abs().max())
    assert np.abs(gflat.cpu().numpy() - ref.numpy()).max() <= 1e-3 * scale + 1e-9

    # Adam: three steps on the same gradient against the oracle's restatement of torch.optim.Adam
    st = O.AdamState()
    p_flat = torch.cat([t.flatten() for t in oprm.tensors()]).to(DEV)
    m = torch.zeros_like(p_flat)
    v = torch.zeros_like(p_flat)
    for step in range(1, 4):
        O.adam_step(oprm, grads, st)
        ops.adam_step(p_flat, gflat, m, v, step)
    ref_p = torch.cat([t.flatten() for t in oprm.tensors()])
    assert np.abs(p_flat.cpu().numpy() - ref_p.numpy()).max() < 5e-6


def test_ragged_batch_equals_per_slide():
    """One launch over a ragged batch == the same slides one at a time (offsets, region bases, -1 padding)."""
    from moc_b200 import _lib, ops, synthetic
    c, j, k = 2, 50, 10
    sizes = [1, 3, 130, 64, 257, 49, 1000]
    w, we = synthetic.prompt_matrices(c)
    bags, _ = synthetic.make_cohort(len(sizes), sizes, c, cohort_seed=9)
    offs = [0]
    for b in bags:
        offs.append(offs[-1] + b.size(0))
    feat = torch.cat(bags).to(DEV)
    pr = ops.Prompts.pack(w.to(DEV), we.to(DEV))
    keys = ops.score_keys(feat, pr)
    offs_d = torch.tensor(offs, dtype=torch.int64, device=DEV)
    sel = ops.select_union(keys, offs_d, offs, c, j)
    oprm = O.SenetParams.init(1)
    prm = ops.HeadParams(oprm.w1.to(DEV), oprm.b1.to(DEV), oprm.w2.to(DEV), oprm.b2.to(DEV))
    out = ops.head_forward(feat, keys, c, sel, prm, _lib.active_bits((), "eval"), k)
    cnt = sel.sel_count.cpu().tolist()
    for i, x in enumerate(bags):
        slide = O.slide_process(x, w, we, c, j)
        b = sel.sel_base_h[i]
        assert sel.sel_local[b:b + cnt[i]].cpu().tolist() == slide["selected_index"]
        g, _ = O.senet_forward(oprm, slide["selected_feat"])
        ref = O.bag_logits(O.combine(g, slide), k)
        close(out.bag_logits[i:i + 1], ref)


def test_errors_are_loud():
    from moc_b200 import ops
    from moc_b200._lib import MocError
    with pytest.raises(MocError):
        ops.score_keys(torch.zeros(4, 512), None)  # CPU tensor: there is no CPU path
    w = torch.zeros(512, 2, device=DEV)
    with pytest.raises(MocError):
        ops.Prompts.pack(w, torch.zeros(512, 2, device=DEV))  # needs background columns
    with pytest.raises(MocError):
        ops.Prompts.pack(w, torch.zeros(512, 80, device=DEV))  # wider than this build supports


def test_tensor_core_head_matches_cuda_core_head():
    """The tcgen05 (3xTF32) gate kernel against the fp32 CUDA-core kernel on the same rows, dense senet mode."""
    import os
    import subprocess
    import sys
    from moc_b200 import ops
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(1000, 512, generator=gen) * 0.3
    x[5] = 0
    x[17] *= 40.0  # large rows: TF32 splitting must stay accurate across magnitudes
    oprm = O.SenetParams.init(11)
    prm = ops.HeadParams(oprm.w1.to(DEV), oprm.b1.to(DEV), oprm.w2.to(DEV), oprm.b2.to(DEV))
    g_tc = ops.senet_forward(x.to(DEV), prm).cpu()
    g_ref, _ = O.senet_forward(oprm, x)
    assert (g_tc - g_ref).abs().max().item() < 2e-6
    # the CUDA-core implementation is selected per process through the environment
    code = ("import torch,sys;sys.path.insert(0,%r);from moc_b200 import ops;from oracle import moc_oracle as O;"
            "g=torch.Generator().manual_seed(3);x=torch.randn(1000,512,generator=g)*0.3;x[5]=0;x[17]*=40.0;"
            "p=O.SenetParams.init(11);q=ops.HeadParams(p.w1.cuda(),p.b1.cuda(),p.w2.cuda(),p.b2.cuda());"
            "torch.save(ops.senet_forward(x.cuda(),q).cpu(),sys.argv[1])") % os.path.dirname(os.path.dirname(__file__))
    out = "/tmp/moc_simt_gate.pt"
    for impl in ("simt", "tf32", "f16"):   # default (auto) above = FP16x3 here; every kernel against the others
        subprocess.run([sys.executable, "-c", code, out], check=True, env=dict(os.environ, MOC_HEAD_IMPL=impl), timeout=120)
        g_other = torch.load(out)
        assert (g_tc - g_other).abs().max().item() < 2e-6, impl


def test_head_f16_domain():
    """The FP16x3 gate kernel splits x * 16 into two halves: |x| < 4094 is its domain.  Inside it, rows of very
    different magnitudes (1e-6 .. 2000) stay fp32-accurate; outside it the gate comes out non-finite (loud), which is
    what MOC_HEAD_IMPL=tf32 is for."""
    from moc_b200 import ops
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(512, 512, generator=gen)
    scale = torch.logspace(-6, 2.7, 512).unsqueeze(1)          # row magnitudes 1e-6 .. 500 (entries up to ~2000)
    x = x * scale
    oprm = O.SenetParams.init(13)
    prm = ops.HeadParams(oprm.w1.to(DEV), oprm.b1.to(DEV), oprm.w2.to(DEV), oprm.b2.to(DEV))
    g32, _ = O.senet_forward(oprm, x)
    h64 = torch.relu(x.double() @ oprm.w1.double().t() + oprm.b1.double())
    g64 = torch.sigmoid(h64 @ oprm.w2.double().t() + oprm.b2.double())
    got = ops.senet_forward(x.to(DEV), prm).cpu()
    assert torch.isfinite(got).all()
    # pre-activations reach the hundreds here, so fp32 itself is ~1e-5 off float64: the kernel must be as good
    err32 = (g32.double() - g64).abs().max().item()
    assert (got.double() - g64).abs().max().item() <= 3 * err32 + 2e-6
    x[7, 3] = 5000.0
    bad = ops.senet_forward(x.to(DEV), prm, wide=False).cpu()      # the fast kernel alone: loud, and it raises the flag
    assert not torch.isfinite(bad[7]).all() and torch.isfinite(bad[8:]).all()
    assert ops.head_workspace(DEV).overflowed()
    # default: the flag is polled and the launch repeated on the range-free 3xTF32 kernel - finite, like the reference
    x[9, 100] = -1.0e5
    g32, _ = O.senet_forward(oprm, x)
    h64 = torch.relu(x.double() @ oprm.w1.double().t() + oprm.b1.double())
    g64 = torch.sigmoid(h64 @ oprm.w2.double().t() + oprm.b2.double())
    fixed = ops.senet_forward(x.to(DEV), prm).cpu()
    assert torch.isfinite(fixed).all()
    err32 = (g32.double() - g64).abs().max().item()       # what fp32 itself loses against float64 on these rows
    # rows 7 and 9 now have pre-activations in the thousands: the 3xTF32 tensor-core accumulation (truncating) stays
    # within a small multiple of what fp32 itself loses there
    assert (fixed.double() - g64).abs().max().item() <= 6 * err32 + 5e-6


@pytest.mark.parametrize("c,big", [(2, 5000.0), (2, 1.0e5), (30, 5000.0), (30, 1.0e5)])
def test_out_of_range_features_match_the_oracle(c, big):
    """Features beyond the fast kernels' range (|x| >= 4094 for the FP16x3 gate MLP, >= 65504 for the tensor-core
    scoring of wide prompt sets) are finite in the reference (fp32 matmul); the flag-checked passes of the engine and
    slide_process must re-dispatch to the range-free kernels and return the oracle's finite values."""
    from moc_b200 import RaggedBagStore, ops, slide_process
    from moc_b200.engine import MocEngine
    from moc_b200 import synthetic
    j, k = 100, 10
    w, we = synthetic.prompt_matrices(c)
    bags, labels = synthetic.make_cohort(3, [700, 900, 650], c, cohort_seed=31)
    bags[1][5, 17] = big
    bags[1][600, 300] = -big
    oprm = O.SenetParams.init(3)
    prm = ops_params(oprm)
    eng = MocEngine(w.to(DEV), we.to(DEV), j, k)
    store = RaggedBagStore.from_bags(bags, labels, DEV)
    unchecked = eng.eval_logits(store, prm)
    got = eng.eval_logits(store, prm, check_domain=True)
    assert eng.is_wide(store) and torch.isfinite(got).all()
    if c == 2:      # the FP16x3 gate kernel alone is loud about it: non-finite gates, non-finite bag logits
        assert not torch.isfinite(unchecked[1]).all()
    else:           # and it raises its flag; the tensor-core scoring of the wide prompt set has its own, for |x| >= 65504
        ws = ops.head_workspace(DEV)
        ws.clear_flag()
        eng.prompts.tc_flag.zero_()
        eng.eval_logits(store.__class__.from_bags(bags, labels, DEV), prm)
        assert ws.overflowed() or int(eng.prompts.tc_flag.item()) != 0
        assert (int(eng.prompts.tc_flag.item()) != 0) == (big >= 65504)
    for i, x in enumerate(bags):
        ref = O.slide_eval_logits(oprm, x, w, we, c, j, k)
        close(got[i:i + 1], ref, rtol=1e-3, atol=1e-5)
    zs = eng.zero_shot_logits(RaggedBagStore.from_bags(bags, labels, DEV), check_domain=True)
    for i, x in enumerate(bags):
        close(zs[i:i + 1], O.topj_pooling(x @ w, [k])[1][k], rtol=1e-3, atol=1e-5)
    # a fresh in-range store stays on the fast kernels
    ok_store = RaggedBagStore.from_bags([bags[0], bags[2]], [0, 1], DEV)
    eng.eval_logits(ok_store, prm, check_domain=True)
    assert not eng.is_wide(ok_store)
    # module-level API
    r = slide_process(bags[1], w.to(DEV), we.to(DEV), n_classes=c, topj=j)
    ref = O.slide_process(bags[1], w, we, c, j)
    assert r["selected_index"] == ref["selected_index"]
    close(r["logits_top_classifier"], ref["logits_top_classifier"], rtol=1e-3, atol=1e-4)
    close(r["logits_bottomk_irrel_classifier"], ref["logits_bottomk_irrel_classifier"], rtol=1e-3, atol=1e-4)


def ops_params(oprm):
    from moc_b200 import ops
    return ops.HeadParams(oprm.w1.to(DEV), oprm.b1.to(DEV), oprm.w2.to(DEV), oprm.b2.to(DEV))


@pytest.mark.parametrize("name", ["bank_rcc_ext", "bank_stress"])
def test_collapse_prompt_bank_golden(golden, name):
    """A prompt bank (cfg3: >= 64 prompts per class) collapses to the reference's classifier matrix, and scoring
    against the collapsed column is the (rescaled) mean of the per-prompt scores."""
    from moc_b200 import ops
    g = golden(name)
    bank = T(g["bank"]).float()
    counts = g["prompts_per_class"].tolist()
    w = ops.collapse_prompt_bank(bank.to(DEV), counts).cpu()
    close(w, g["W"], rtol=1e-5, atol=1e-7)
    close(w, O.collapse_prompt_bank(bank, counts), rtol=1e-5, atol=1e-7)
    x = torch.randn(64, 512, generator=torch.Generator().manual_seed(1))
    p0 = 0
    for c, n in enumerate(counts):
        e = torch.nn.functional.normalize(bank[p0:p0 + n], dim=-1)
        mean_score = (x @ e.t()).mean(dim=1)
        close((x @ w[:, c]) * e.mean(dim=0).norm(), mean_score, rtol=1e-4, atol=1e-5)
        p0 += n


def _bank_case(counts, n_bg, seed):
    """(bank [n_prompts,512], W collapsed [512,C], W_ext [512,C+n_bg]) with the reference's collapse (oracle)."""
    gen = torch.Generator().manual_seed(seed)
    c = len(counts)
    centre = torch.randn(c, 1, 512, generator=gen)
    rows = [centre[i] + 0.5 * torch.randn(n, 512, generator=gen) for i, n in enumerate(counts)]
    bank = torch.cat(rows) * (0.5 + torch.rand(sum(counts), 1, generator=gen))      # un-normalised, like raw embeddings
    w = O.collapse_prompt_bank(bank, counts)
    bg = torch.randn(n_bg, 512, generator=gen)
    bg = bg / bg.norm(dim=1, keepdim=True)
    return bank, w, torch.cat([w, bg.t()], dim=1).contiguous()


@pytest.mark.parametrize("counts,n_bg", [([64, 64, 64], 4), ([66, 61, 71], 4), ([1, 200, 5], 3), ([17, 15], 1),
                                         ([30] * 8, 4), ([2, 3], 4), ([100, 100], 56)])
@pytest.mark.parametrize("n_rows", [1, 127, 129, 1000, 4099])
def test_uncollapsed_prompt_bank_scoring(counts, n_bg, n_rows):
    """BASELINE configs[2] as SURVEY 8d defines it: the bank stays [512, C*P] on the scoring path, the kernel takes
    the group mean and the 1/||mean|| rescale (utils/zeroshot_utils.py:38-44) in its epilogue; keys must equal the
    oracle's on the collapsed matrix, and the fp32 kernel's on the collapsed matrix."""
    from moc_b200 import ops
    bank, w, we = _bank_case(counts, n_bg, seed=sum(counts) + n_bg)
    c = len(counts)
    gen = torch.Generator().manual_seed(n_rows)
    x = torch.randn(n_rows, 512, generator=gen)
    x = x / x.norm(dim=1, keepdim=True)
    bp = ops.BankPrompts.pack(bank.to(DEV), counts, we.to(DEV))
    keys = ops.score_keys(x.to(DEV), bp)
    assert keys.shape == (2 * c + 3, n_rows)
    close(keys.cpu().numpy(), oracle_keys(x, w, we, c), rtol=1e-3, atol=2e-6)
    ref = ops.score_keys(x.to(DEV), ops.Prompts.pack(w.to(DEV), we.to(DEV)))
    assert (keys[:c] - ref[:c]).abs().max().item() < 4e-6
    assert int(bp.tc_flag.item()) == 0
    # L2-normalise flag and the range-free fallback
    xs = x * 3.0
    close(ops.score_keys(xs.to(DEV), bp, normalize=True).cpu().numpy(), oracle_keys(x, w, we, c), rtol=1e-3, atol=3e-6)
    xs[0, 3] = 1.0e5
    got = ops.score_keys(xs.to(DEV), bp, check_domain=True)
    assert torch.isfinite(got).all()
    close(got[:c].cpu().numpy(), (xs @ w).t().numpy(), rtol=1e-3, atol=1e-4)


def test_uncollapsed_bank_golden_and_engine(golden):
    """The reference-collapsed matrix of tests/golden/bank_stress.npz (zero_shot_classifier's own output) against the
    un-collapsed kernel, and a whole evaluation pass with the bank on the scoring path."""
    from moc_b200 import RaggedBagStore, ops, synthetic
    from moc_b200.engine import MocEngine
    g = golden("bank_stress")
    bank = T(g["bank"]).float()
    counts = g["prompts_per_class"].tolist()
    c = len(counts)
    w = T(g["W"]).float()
    _, we_syn = synthetic.prompt_matrices(c)
    we = torch.cat([w, we_syn[:, c:]], dim=1).contiguous()
    bags, labels = synthetic.make_cohort(4, [3000, 2500, 4100, 777], c, cohort_seed=5, w_ext=we)
    store = RaggedBagStore.from_bags(bags, labels, DEV)
    oprm = O.SenetParams.init(2)
    prm = ops_params(oprm)
    j, k = 200, 10
    eng_bank = MocEngine(w.to(DEV), we.to(DEV), j, k, prompt_bank=(bank, counts))
    eng = MocEngine(w.to(DEV), we.to(DEV), j, k)
    keys_b = eng_bank.keys_for(store)
    for i, x in enumerate(bags):
        close(keys_b[:, store.offsets_h[i]:store.offsets_h[i + 1]].cpu().numpy(), oracle_keys(x, w, we, c), rtol=1e-3, atol=2e-6)
    lb, l0 = eng_bank.eval_logits(store, prm), eng.eval_logits(store, prm)
    close(lb, l0, rtol=1e-3, atol=1e-5)
    for i, x in enumerate(bags):
        close(lb[i:i + 1], O.slide_eval_logits(oprm, x, w, we, c, j, k), rtol=1e-3, atol=1e-5)
    close(eng_bank.zero_shot_logits(store), eng.zero_shot_logits(store), rtol=1e-3, atol=1e-5)
    with pytest.raises(Exception):
        MocEngine(torch.roll(w, 1, dims=1).to(DEV), we.to(DEV), j, k, prompt_bank=(bank, counts))
