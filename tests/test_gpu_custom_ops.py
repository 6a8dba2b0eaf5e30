"""GPU: the torch.library custom ops (torch.ops.eavit_b200.*) -- forward and autograd against plain torch fp32."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _register():
    import eavit_b200  # noqa: F401  (registers the ops)


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


def test_linear_op_forward_backward():
    torch.manual_seed(0)
    x = torch.randn(777, 256, device="cuda").bfloat16().requires_grad_(True)
    w = (torch.randn(1024, 256, device="cuda") / 16).bfloat16().requires_grad_(True)
    b = torch.randn(1024, device="cuda", requires_grad=True)
    y = torch.ops.eavit_b200.linear(x, w, b)
    g = torch.randn_like(y)
    y.backward(g)
    xr, wr, br = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = F.linear(xr, wr, br)
    yr.backward(g.float())
    assert rel(y, yr) < 5e-3
    assert rel(x.grad, xr.grad) < 5e-3 and rel(w.grad, wr.grad) < 5e-3 and rel(b.grad, br.grad) < 1e-3


def test_layer_norm_op():
    torch.manual_seed(1)
    x = torch.randn(500, 256, device="cuda", requires_grad=True)
    g = torch.randn(256, device="cuda", requires_grad=True)
    b = torch.randn(256, device="cuda", requires_grad=True)
    y, mean, rstd = torch.ops.eavit_b200.layer_norm(x, g, b, 1e-5)
    dy = torch.randn(500, 256, device="cuda").bfloat16()
    y.backward(dy)
    xr, gr, br = (t.detach().clone().requires_grad_(True) for t in (x, g, b))
    yr = F.layer_norm(xr, (256,), gr, br, 1e-5)
    yr.backward(dy.float())
    assert rel(y, yr) < 5e-3
    assert rel(x.grad, xr.grad) < 1e-4 and rel(g.grad, gr.grad) < 1e-4 and rel(b.grad, br.grad) < 1e-4
    assert rel(mean, xr.detach().mean(1)) < 1e-5


def test_attention_op_autograd():
    torch.manual_seed(2)
    lens, H, Dh = [196, 197, 50], 8, 32
    starts = [0]
    for n in lens:
        starts.append(starts[-1] + n)
    T = starts[-1]
    qkv = (torch.randn(T, 3 * H * Dh, device="cuda") * 1.2).bfloat16().requires_grad_(True)
    ss = torch.tensor(starts, dtype=torch.int32, device="cuda")
    out, lse = torch.ops.eavit_b200.attention(qkv, ss, max(lens), H, Dh ** -0.5)
    dout = torch.randn_like(out)
    out.backward(dout)
    x = qkv.detach().float().requires_grad_(True)
    q, k, v = x.split(H * Dh, dim=1)
    outs = []
    for s0, s1 in zip(starts[:-1], starts[1:]):
        hs = []
        for h in range(H):
            sl = slice(h * Dh, (h + 1) * Dh)
            hs.append(((q[s0:s1, sl] @ k[s0:s1, sl].t()) * Dh ** -0.5).softmax(-1) @ v[s0:s1, sl])
        outs.append(torch.cat(hs, 1))
    ref = torch.cat(outs, 0)
    (ref * dout.float()).sum().backward()
    assert rel(out, ref) < 6e-3
    assert rel(qkv.grad, x.grad) < 1.2e-2


def test_numerics_ops():
    torch.manual_seed(3)
    E, T = 64, 128
    r = torch.rand(E, T, device="cuda")
    v = torch.randn(E, T + 1, device="cuda")
    ret, adv = torch.ops.eavit_b200.gae(r, v, 0.99, 0.95)
    gae = torch.zeros(E, device="cuda", dtype=torch.float64)
    ref = torch.zeros(E, T, device="cuda", dtype=torch.float64)
    for t in reversed(range(T)):
        delta = r[:, t].double() + 0.99 * v[:, t + 1].double() - v[:, t].double()
        gae = delta + 0.99 * 0.95 * gae
        ref[:, t] = gae + v[:, t].double()
    assert rel(ret.reshape(E, T), ref) < 1e-5
    assert rel(adv.reshape(E, T), ref - v[:, :-1].double()) < 1e-5
    a, b = torch.randn(100, 512, device="cuda"), torch.randn(100, 512, device="cuda")
    assert rel(torch.ops.eavit_b200.intrinsic_mse(a, b), (a - b).pow(2).mean(1)) < 1e-6


def test_ops_have_no_cpu_path():
    x = torch.randn(4, 8).bfloat16()
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.eavit_b200.linear(x, x, None)
