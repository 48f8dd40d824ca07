"""Out-of-bounds guards (compute-sanitizer is closed on this pool): outputs embedded in larger sentinel-filled buffers
must leave every byte outside their extent untouched, for row counts that end inside a stage / tile / super-group."""
import pytest
import torch

from moc_b200 import ops, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"
SENTINEL = -12345.678


@pytest.mark.parametrize("n_rows", [1, 3, 31, 33, 127, 129, 4097, 20001])
@pytest.mark.parametrize("kind", ["regw", "tc", "bank"])
def test_score_kernels_write_only_their_rows(kind, n_rows):
    gen = torch.Generator().manual_seed(n_rows)
    x = torch.randn(n_rows + 7, 512, generator=gen)
    feat = x.to(DEV)[:n_rows]                       # rows past the end exist in memory: reading them would go unnoticed,
    c = {"regw": 2, "tc": 30, "bank": 3}[kind]      # writing keys for them would not
    w, we = synthetic.prompt_matrices(c, device=DEV)
    if kind == "bank":
        bank, w = synthetic.prompt_bank(c, 40, device=DEV)
        we = torch.cat([w, we[:, c:]], dim=1).contiguous()
        pr = ops.BankPrompts.pack(bank.t().contiguous(), [40] * c, we)
    else:
        pr = ops.Prompts.pack(w, we)
    planes, pad = ops.num_key_planes(c), 64
    big = torch.full((planes + 2, n_rows + pad), SENTINEL, device=DEV)
    out = big[1:planes + 1, :n_rows]
    got = ops.score_keys(feat, pr, out=out)
    assert got.data_ptr() == out.data_ptr()
    torch.cuda.synchronize()
    assert bool((big[0] == SENTINEL).all()) and bool((big[planes + 1] == SENTINEL).all())
    assert bool((big[1:planes + 1, n_rows:] == SENTINEL).all())
    assert bool(torch.isfinite(out).all()) and not bool((out == SENTINEL).any())
    ref = ops.score_keys(feat.clone(), pr)          # and the strided output equals the dense one
    assert torch.equal(ref, out)


def test_selection_and_gate_outputs_stay_inside_their_regions():
    """sel_rows / sel_local regions, final scores and gates of a ragged batch: slots past a slide's count are -1 /
    untouched, nothing is written past the capacity."""
    from moc_b200 import _lib
    c, j, k = 2, 100, 10
    w, we = synthetic.prompt_matrices(c, device=DEV)
    sizes = [130, 5, 999, 1]
    offs = [0]
    for n in sizes:
        offs.append(offs[-1] + n)
    feat = torch.randn(offs[-1], 512, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3))
    keys = ops.score_keys(feat, ops.Prompts.pack(w, we))
    offs_d = torch.tensor(offs, dtype=torch.int64, device=DEV)
    sel = ops.select_union(keys, offs_d, offs, c, j)
    counts = sel.sel_count.cpu().tolist()
    for i, n in enumerate(sizes):
        lo, hi = sel.sel_base_h[i], sel.sel_base_h[i + 1]
        assert hi - lo == min(n, j * (2 * c + 2)) and 0 < counts[i] <= hi - lo
        assert bool((sel.sel_rows[lo + counts[i]:hi] == -1).all())
        r = sel.sel_rows[lo:lo + counts[i]]
        assert int(r.min()) >= offs[i] and int(r.max()) < offs[i + 1]
    g = torch.Generator().manual_seed(0)
    prm = ops.HeadParams((torch.rand(64, 512, generator=g) - 0.5).to(DEV) * 0.08, torch.zeros(64, device=DEV),
                         (torch.rand(4, 64, generator=g) - 0.5).to(DEV) * 0.2, torch.zeros(4, device=DEV))
    out = ops.head_forward(feat, keys, c, sel, prm, _lib.CLS_ALL, k, want_gate=True)
    assert out.final.shape == (sel.capacity, c) and out.gate.shape == (sel.capacity, 4)
    for i in range(len(sizes)):
        lo = sel.sel_base_h[i]
        assert bool(torch.isfinite(out.final[lo:lo + counts[i]]).all())
        assert bool(((out.gate[lo:lo + counts[i]] > 0) & (out.gate[lo:lo + counts[i]] < 1)).all())
    assert bool(torch.isfinite(out.bag_logits).all())
    assert int(out.pool_pos.max()) < max(counts)
