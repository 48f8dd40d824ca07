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
