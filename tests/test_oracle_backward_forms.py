"""CPU: closed-form backward formulas (oracle/backward_forms.py) against autograd of the pinned forward oracle."""
import torch
import torch.nn.functional as F

from oracle import backward_forms as BF
from oracle import train_oracle as TO
from oracle.gcp_oracle import lstm_cell

torch.manual_seed(0)
D = torch.float64


def _close(a, b, tol=1e-9):
    assert float((a - b).abs().max()) <= tol * max(float(b.abs().max()), 1.0), float((a - b).abs().max())


def test_dlm_nll_grad():
    m = torch.randn(3, 5, 3, 8, 8, dtype=D, requires_grad=True)
    s = (torch.randn(3, 5, 3, 8, 8, dtype=D) * 0.5 - 2.0).requires_grad_(True)
    x = torch.rand(3, 3, 8, 8, dtype=D) * 2 - 1
    x[0, 0, 0, :4] = -1.0            # both open-ended edge bins occur
    x[1, 1, 1, :4] = 1.0
    TO.dlm_nll(torch.sigmoid(m), s, x).sum().backward()
    dm, ds = BF.dlm_nll_grad(m.detach(), s.detach(), x)
    _close(dm, m.grad)
    _close(ds, s.grad)


def test_kl_grad():
    t = [torch.randn(4, 7, 16, dtype=D, requires_grad=True) for _ in range(4)]
    TO.kl_gauss(*t).sum().backward()
    for got, ref in zip(BF.kl_gauss_grad(*[a.detach() for a in t]), t):
        _close(got, ref.grad)


def test_group_norm_grad():
    x = torch.randn(9, 128, dtype=D, requires_grad=True)
    gamma = torch.randn(128, dtype=D, requires_grad=True)
    beta = torch.randn(128, dtype=D, requires_grad=True)
    dy = torch.randn(9, 128, dtype=D)
    (F.group_norm(x, 8, gamma, beta, 1e-5) * dy).sum().backward()
    dx, dg, db = BF.group_norm_rows_grad(x.detach(), gamma.detach(), beta.detach(), dy)
    _close(dx, x.grad)
    _close(dg, gamma.grad)
    _close(db, beta.grad)


def test_lstm_cell_grad():
    H = 12
    x = torch.randn(5, 20, dtype=D)
    h0 = torch.randn(5, H, dtype=D)
    c0 = torch.randn(5, H, dtype=D, requires_grad=True)
    w_ih = torch.randn(4 * H, 20, dtype=D)
    w_hh = torch.randn(4 * H, H, dtype=D)
    b = torch.randn(4 * H, dtype=D)
    gates = (F.linear(x, w_ih, b) + F.linear(h0, w_hh)).requires_grad_(True)
    # the oracle's cell on the same pre-activations: identity input weights, zero hidden weights
    h, c = lstm_cell(gates, h0, c0, torch.eye(4 * H, dtype=D), torch.zeros(4 * H, H, dtype=D), torch.zeros(4 * H, dtype=D),
                     torch.zeros(4 * H, dtype=D))
    dh, dc = torch.randn(5, H, dtype=D), torch.randn(5, H, dtype=D)
    ((h * dh).sum() + (c * dc).sum()).backward()
    d_gates, d_c0 = BF.lstm_cell_grad(gates.detach(), c0.detach(), dh, dc)
    _close(d_gates, gates.grad)
    _close(d_c0, c0.grad)
