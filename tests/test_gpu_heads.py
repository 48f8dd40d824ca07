"""Secondary MIL heads (Conch_CLIP_Ada, ABMIL = CLAM_SB, MIL_fc) and the tensor-core linear layer on the GPU,
against the oracle and the goldens produced by the reference's own modules."""
import numpy as np
import pytest
import torch

from oracle import moc_oracle_heads as H
from tests.helpers import close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _sd(g):
    return {k[3:]: T(v).float() for k, v in g.items() if k.startswith("sd_")}


@pytest.mark.parametrize("n,k,m", [(1, 512, 64), (127, 512, 128), (129, 128, 512), (1000, 384, 512), (4099, 512, 768),
                                   (300, 512, 2), (20000, 512, 130), (513, 32, 5)])
@pytest.mark.parametrize("act", [None, "relu", "tanh", "sigmoid"])
def test_linear_tc_matches_fp64(n, k, m, act):
    """Y = act(X W^T + b) on tcgen05 (3xTF32) against float64, ragged row and column tails.  Observed error is a few
    1e-6 relative (the operand split is exact to 2^-21; the rest is the tensor core's own accumulation), two orders
    inside the 1e-3 parity bar."""
    from moc_b200 import ops
    g = torch.Generator().manual_seed(n + k + m)
    x = torch.randn(n, k, generator=g)
    w = torch.randn(m, k, generator=g) * k ** -0.5
    b = torch.randn(m, generator=g)
    y = ops.linear(x.to(DEV), w.to(DEV), b.to(DEV), act).cpu()
    z = x.double() @ w.double().t() + b.double()
    ref = {None: z, "relu": z.clamp(min=0), "tanh": z.tanh(), "sigmoid": z.sigmoid()}[act]
    assert y.shape == (n, m)
    assert ((y.double() - ref).abs() / ref.abs().clamp(min=1.0)).max() < 2e-5
    y2 = ops.linear(x.to(DEV), w.to(DEV), None, "tanh", split=m // 2, act_tail="sigmoid").cpu()
    z2 = x.double() @ w.double().t()
    assert (y2[:, :m // 2].double() - z2[:, :m // 2].tanh()).abs().max() < 2e-5
    assert (y2[:, m // 2:].double() - z2[:, m // 2:].sigmoid()).abs().max() < 2e-5


@pytest.mark.parametrize("name", ["heads_clip_ada_c2", "heads_clip_ada_c3"])
def test_conch_clip_ada_golden(golden, name):
    import moc_b200
    g = golden(name)
    sd, cl = _sd(g), T(g["classifier"])
    m = moc_b200.Conch_CLIP_Ada(512, 4, int(g["C"]), cl.to(DEV), float(g["clip_ratio"]), int(g["topj"])).to(DEV)
    m.load_state_dict(sd)
    for i in range(int(g["n_bags"])):
        x = T(g["feat_%d" % i]).float()
        close(m.forward(x.to(DEV)), g["forward_%d" % i])
        close(m.forward_disable_ada(x.to(DEV)), g["forward_disable_ada_%d" % i])
        close(m.forward(x.to(DEV)), H.clip_ada_forward(sd, cl, x, float(g["clip_ratio"]), int(g["topj"])))
        close(m.topj_pooling((x @ cl).to(DEV), 7), H.topj_mean(x @ cl, 7))


@pytest.mark.parametrize("name", ["heads_abmil_c2", "heads_abmil_c3"])
def test_abmil_golden(golden, name):
    import moc_b200
    g = golden(name)
    sd = _sd(g)
    m = moc_b200.CLAM_SB(gate=True, size_arg="conch", dropout=False, n_classes=int(g["C"]), instance_loss_fn=None).to(DEV).eval()
    m.load_state_dict(sd)
    for i in range(int(g["n_bags"])):
        x = T(g["feat_%d" % i]).float().to(DEV)
        logits, y_prob, y_hat, a_raw, res = m(x, return_features=True)
        close(logits, g["logits_%d" % i])
        close(y_prob, g["y_prob_%d" % i])
        assert y_hat.cpu().numpy().tolist() == g["y_hat_%d" % i].tolist()
        close(a_raw, g["a_raw_%d" % i], rtol=1e-3, atol=2e-6)
        close(res["features"], g["features_%d" % i], rtol=1e-3, atol=2e-6)
        close(m(x, attention_only=True), g["attention_only_%d" % i], rtol=1e-3, atol=2e-6)


def test_abmil_large_bag_vs_oracle():
    """20 000-patch bag: the partial/merge softmax pooling across many blocks against the oracle."""
    import moc_b200
    torch.manual_seed(7)
    m = moc_b200.CLAM_SB(size_arg="conch", n_classes=2, instance_loss_fn=None).eval()
    x = torch.randn(20000, 512) * 0.2
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ref = H.abmil_forward(sd, x)
    m = m.to(DEV)
    logits, y_prob, y_hat, a_raw, _ = m(x.to(DEV))
    close(logits, ref[0])
    close(y_prob, ref[1])
    close(a_raw, ref[3], rtol=1e-3, atol=2e-6)


def test_mil_fc_golden(golden):
    import moc_b200
    g = golden("heads_mil_fc")
    sd = _sd(g)
    m = moc_b200.MIL_fc(size_arg="benchmark", n_classes=2, top_k=1).to(DEV).eval()
    m.load_state_dict(sd)
    for i in range(int(g["n_bags"])):
        x = T(g["feat_%d" % i]).float().to(DEV)
        top, y_prob, y_hat, y_probs, _ = m(x)
        close(top, g["top_instance_%d" % i])
        close(y_prob, g["y_prob_%d" % i])
        assert y_hat.cpu().numpy().tolist() == g["y_hat_%d" % i].tolist()
        close(y_probs, g["y_probs_%d" % i])


def close_grad(a, b, rtol=1e-3, scale_atol=2e-5):
    """Gradients are long sums with cancellation: elementwise 1e-3 relative plus 2e-5 of the tensor's largest entry."""
    b = b.detach().cpu().double() if isinstance(b, torch.Tensor) else torch.from_numpy(np.asarray(b, dtype=np.float64))
    a = a.detach().cpu().double().reshape(b.shape)
    tol = rtol * b.abs() + scale_atol * b.abs().max() + 1e-12
    err = (a - b).abs()
    assert (err <= tol).all(), "max err %.3e, tol there %.3e, max|ref| %.3e" % (err.max(), tol.flatten()[err.argmax()], b.abs().max())


@pytest.mark.parametrize("n,m,k", [(1, 4, 4), (31, 128, 128), (33, 132, 64), (1000, 768, 512), (4099, 512, 512),
                                   (20000, 768, 512), (20000, 512, 512), (257, 8, 516), (50000, 128, 512)])
def test_linear_wgrad_matches_fp64(n, m, k):
    """dW = G^T X on tcgen05 (3xTF32, split over the bag, fixed-order reduction) against float64; ragged bag tails,
    partial output tiles, one work item per SM and several."""
    from moc_b200 import ops
    gen = torch.Generator().manual_seed(n + m + k)
    g = torch.randn(n, m, generator=gen)
    x = torch.randn(n, k, generator=gen)
    ref = g.double().t() @ x.double()
    dw = ops.linear_wgrad(g.to(DEV), x.to(DEV))
    assert dw.shape == (m, k)
    scale = float(n) ** 0.5   # standard deviation of an entry of the reference

    def relerr(got, want):
        return float(((got.cpu().double() - want).abs() / (want.abs() + scale)).max())
    assert relerr(dw, ref) < 5e-5   # observed 2.4e-5: the tensor core truncates (toward zero) when it accumulates
    dw2 = ops.linear_wgrad(g.to(DEV), x.to(DEV))
    assert torch.equal(dw, dw2), "the split-K reduction must be deterministic"
    acc = torch.ones(m, k, device=DEV)
    ops.linear_wgrad(g.to(DEV), x.to(DEV), out=acc, accumulate=True)
    assert relerr(acc - 1.0, ref) < 5e-5
    # strided operands: column slices of wider buffers
    if m % 4 == 0 and k % 4 == 0:
        gw = torch.randn(n, m + 8, generator=gen)
        xw = torch.randn(n, k + 4, generator=gen)
        dws = ops.linear_wgrad(gw.to(DEV)[:, 4:4 + m], xw.to(DEV)[:, :k])
        refs = gw[:, 4:4 + m].double().t() @ xw[:, :k].double()
        assert relerr(dws, refs) < 5e-5


@pytest.mark.parametrize("name", ["heads_abmil_c2", "heads_abmil_c3"])
def test_abmil_backward_golden(golden, name):
    """One training step as utils/core_utils.py:391-414 runs it (logits = model(data); CE; loss.backward()) against
    the gradients torch autograd produced through the reference's own CLAM_SB."""
    import moc_b200
    g, gb = golden(name), golden(name.replace("abmil", "abmil_bwd"))
    sd = _sd(g)
    m = moc_b200.CLAM_SB(gate=True, size_arg="conch", dropout=False, n_classes=int(g["C"]), instance_loss_fn=None).to(DEV).train()
    m.load_state_dict(sd)
    x = T(g["feat_%d" % int(gb["bag"])]).float().to(DEV)
    label = torch.tensor([int(gb["label"])], device=DEV)
    logits, y_prob, y_hat, _, _ = m(x)
    assert logits.requires_grad and not y_prob.requires_grad
    loss = torch.nn.CrossEntropyLoss()(logits, label)
    assert abs(float(loss.detach()) - float(gb["loss"])) < 1e-5
    loss.backward()
    oloss, ograds = H.abmil_loss_and_grads(sd, x.cpu(), int(gb["label"]))
    for k, p in m.named_parameters():
        if k.startswith("instance_classifiers"):
            assert p.grad is None
            continue
        assert p.grad is not None and p.grad.shape == p.shape, k
        if k.endswith("attention_c.bias"):
            # d(loss)/d(attention_c.bias) = sum_n dA_n is exactly zero in exact arithmetic (the bag softmax is
            # shift-invariant); the reference's value (1.5e-8) is rounding noise, and so is ours
            assert abs(float(p.grad.reshape(-1)[0])) < 1e-6 and abs(float(np.asarray(gb["grad_" + k]).reshape(-1)[0])) < 1e-6
            continue
        if "grad_" + k in gb:
            close_grad(p.grad, gb["grad_" + k])
        else:
            close_grad(p.grad.reshape(-1)[::5], gb["grad5_" + k])
            assert abs(float(p.grad.double().norm()) - float(gb["norm_" + k])) < 1e-4 * float(gb["norm_" + k])
        close_grad(p.grad, ograds[k])


def test_abmil_training_loop_vs_oracle():
    """Eight Adam steps (two epochs) over four 3 000..20 000-patch bags through the reference's loop shape (forward, CE,
    backward, optimizer.step, zero_grad); every loss - the second epoch's depend on the first epoch's updates - and the
    final parameters against the same loop on the CPU oracle."""
    import moc_b200
    torch.manual_seed(11)
    m = moc_b200.CLAM_SB(size_arg="conch", n_classes=2, instance_loss_fn=None)
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    gen = torch.Generator().manual_seed(12)
    bags = [torch.randn(n, 512, generator=gen) * 0.3 for n in (3000, 20000, 4097, 12345)]
    labels = [0, 1, 1, 0]
    # oracle loop
    names = [k for k in sd0 if not k.startswith("instance_classifiers")]
    leaf = {k: sd0[k].clone().requires_grad_(True) for k in names}
    opt_o = torch.optim.Adam([leaf[k] for k in names], lr=2e-4, weight_decay=1e-5)
    ref_losses = []
    for step in range(8):
        i = step % len(bags)
        logits = H.abmil_forward(leaf, bags[i])[0]
        loss = torch.nn.functional.cross_entropy(logits, torch.tensor([labels[i]]))
        ref_losses.append(float(loss.detach()))
        loss.backward()
        opt_o.step()
        opt_o.zero_grad()
    # ours
    m = m.to(DEV).train()
    opt = torch.optim.Adam([p for k, p in m.named_parameters() if not k.startswith("instance_classifiers")], lr=2e-4,
                           weight_decay=1e-5)
    loss_fn = torch.nn.CrossEntropyLoss()
    for step in range(8):
        i = step % len(bags)
        logits, Y_prob, Y_hat, _, _ = m(bags[i].to(DEV))
        loss = loss_fn(logits, torch.tensor([labels[i]], device=DEV))
        assert abs(float(loss.detach()) - ref_losses[step]) < 2e-4 * max(1.0, abs(ref_losses[step])), (step, float(loss), ref_losses[step])
        loss.backward()
        opt.step()
        opt.zero_grad()
    for k, p in m.named_parameters():
        if k in leaf:
            # Adam's first steps move every weight by ~lr whatever the gradient's size, so an entry whose gradient is
            # rounding noise (e.g. attention_c.bias, exactly zero in exact arithmetic) may walk up to steps * lr away
            diff = (p.detach().cpu() - leaf[k].detach()).abs()
            assert float(diff.max()) <= 8 * 2e-4 * 1.01, k
            if diff.numel() > 1000:
                assert float((diff > 2e-5).float().mean()) < 1e-3, k


@pytest.mark.parametrize("name", ["heads_clip_ada_c2", "heads_clip_ada_c3"])
def test_conch_clip_ada_backward_golden(golden, name):
    """Adapter training step (CE on the scaled pooled logits, loss.backward()) against torch autograd through the
    reference's own Conch_CLIP_Ada: bags with fewer rows than topj and with more."""
    import moc_b200
    g, gb = golden(name), golden(name.replace("clip_ada", "clip_ada_bwd"))
    sd, cl = _sd(g), T(g["classifier"])
    m = moc_b200.Conch_CLIP_Ada(512, 4, int(g["C"]), cl.to(DEV), float(g["clip_ratio"]), int(g["topj"])).to(DEV).train()
    m.load_state_dict(sd)
    for i in range(int(gb["n_bags"])):
        x = T(g["feat_%d" % i]).float().to(DEV)
        m.zero_grad()
        logits = m.forward(x)
        assert logits.requires_grad
        close(logits, g["forward_%d" % i])
        loss = torch.nn.functional.cross_entropy(logits * 56.3477, torch.tensor([int(gb["label_%d" % i])], device=DEV))
        assert abs(float(loss.detach()) - float(gb["loss_%d" % i])) < 2e-4 * max(1.0, float(gb["loss_%d" % i]))
        loss.backward()
        for k, p in m.named_parameters():
            close_grad(p.grad, gb["grad_%d_%s" % (i, k)])
    with torch.no_grad():
        assert not m.forward(x).requires_grad


def test_mil_fc_backward_golden(golden):
    import moc_b200
    g, gb = golden("heads_mil_fc"), golden("heads_mil_fc_bwd")
    m = moc_b200.MIL_fc(size_arg="benchmark", n_classes=2, top_k=1).to(DEV).train()
    m.load_state_dict(_sd(g))
    i = int(gb["bag"])
    x = T(g["feat_%d" % i]).float().to(DEV)
    top, y_prob, y_hat, y_probs, _ = m(x)
    assert top.requires_grad and not y_probs.requires_grad
    close(top, g["top_instance_%d" % i])
    loss = torch.nn.functional.cross_entropy(top, torch.tensor([int(gb["label"])], device=DEV))
    assert abs(float(loss.detach()) - float(gb["loss"])) < 1e-5
    loss.backward()
    for k, p in m.named_parameters():
        close_grad(p.grad, gb["grad_" + k])
    # three SGD steps keep following the oracle
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    opt = torch.optim.SGD(m.parameters(), lr=0.05)
    for step in range(3):
        oloss, ograds = H.mil_fc_loss_and_grads(sd, x.cpu(), 1)
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(m(x)[0], torch.tensor([1], device=DEV))
        assert abs(float(loss.detach()) - float(oloss)) < 1e-5
        loss.backward()
        opt.step()
        sd = {k: sd[k] - 0.05 * ograds[k] for k in sd}


@pytest.mark.parametrize("size_arg,n,c", [("small", 777, 3), ("big", 1300, 2), ("benchmark", 65, 2)])
def test_abmil_backward_other_sizes(size_arg, n, c):
    """The other CLAM size presets (1024- and 384-wide inputs, 256-wide attention) through forward + backward against
    torch autograd over the oracle's forward."""
    import moc_b200
    torch.manual_seed(21)
    m = moc_b200.CLAM_SB(size_arg=size_arg, n_classes=c, instance_loss_fn=None)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.copy_(0.05 * torch.randn(p.shape))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = torch.randn(n, m.size_dict[size_arg][0]) * 0.3
    oloss, ograds = H.abmil_loss_and_grads(sd, x, c - 1)
    m = m.to(DEV).train()
    logits = m(x.to(DEV))[0]
    loss = torch.nn.functional.cross_entropy(logits, torch.tensor([c - 1], device=DEV))
    assert abs(float(loss.detach()) - float(oloss)) < 1e-4 * max(1.0, abs(float(oloss)))
    loss.backward()
    for k, p in m.named_parameters():
        if k.startswith("instance_classifiers"):
            continue
        if k.endswith("attention_c.bias"):
            assert abs(float(p.grad.reshape(-1)[0])) < 1e-6
            continue
        close_grad(p.grad, ograds[k])
