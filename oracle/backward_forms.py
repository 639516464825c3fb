"""Closed-form backward formulas of the element-wise stages of the training step, each checked against autograd of the
pinned forward oracle (tests/test_oracle_backward_forms.py).  TEST INFRASTRUCTURE ONLY (same rules as gcp_oracle.py).

These are the formulas a device backward pass evaluates in its epilogues / element-wise kernels (DESIGN.md section 8,
"Backward pass: plan of record"); writing them down against the reference's forward definitions first keeps the later
CUDA a transcription.  Forward definitions: blox/torch/dist.py:87-130,178-197 (discretised logistic mixture),
dist.py:249-252 (Gaussian KL), torch.nn.GroupNorm, torch.nn.LSTMCell (blox/torch/recurrent_modules.py:195-223).
"""
import torch


def dlm_nll_grad(m, s, x):
    """d nll / d (mean logits m, log-scales s) of train_oracle.dlm_nll applied to mu = sigmoid(m), log_sigma = s.
    m, s [..., 5, C, H, W]; x [..., C, H, W] in [-1, 1].  Returns (dm, ds) with the shapes of m, s."""
    x01 = ((x + 1) / 2).unsqueeze(-4).expand_as(m)
    xb = torch.floor(x01 * 256.0) / 256.0
    mu = torch.sigmoid(m)
    inv = torch.exp(-s)
    d = xb - mu
    hi = torch.sigmoid(d * inv + inv / 256.0)
    lo = torch.sigmoid(d * inv)
    e0, e1 = x01 == 0, x01 == 1
    pr = torch.where(e1, 1 - lo, torch.where(e0, hi, hi - lo))
    pm = pr.mean(-4, keepdim=True)
    dpr = -(1.0 / m.shape[-4]) / (pm + 1e-7)                       # d nll / d pr_k
    ghi = torch.where(e1, torch.zeros_like(hi), hi * (1 - hi))     # the open-ended edge bins drop one of the two cdfs
    glo = torch.where(e0, torch.zeros_like(lo), lo * (1 - lo))
    dxs = ghi - glo                                                # d pr / d xs   (xs = (xb - mu) / scale)
    dinv = ghi * (d + 1.0 / 256.0) - glo * d                       # d pr / d (1 / scale)
    dm = dpr * dxs * (-inv) * mu * (1 - mu)
    ds = dpr * dinv * (-inv)
    return dm, ds


def kl_gauss_grad(q_mu, q_ls, p_mu, p_ls):
    """Gradients of train_oracle.kl_gauss (summed) w.r.t. its four inputs."""
    ip2 = torch.exp(-2 * p_ls)
    diff = q_mu - p_mu
    dq_mu = diff * ip2
    dq_ls = -1 + torch.exp(2 * q_ls) * ip2
    dp_ls = 1 - (torch.exp(2 * q_ls) + diff ** 2) * ip2
    return dq_mu, dq_ls, -dq_mu, dp_ls


def group_norm_rows_grad(x, gamma, beta, dy, groups=8, eps=1e-5):
    """GroupNorm over the channels of each row (the row-MLP layers: x [rows, C], one spatial position).
    Returns (dx, dgamma, dbeta)."""
    R, C = x.shape
    xg = x.reshape(R, groups, C // groups)
    mean = xg.mean(2, keepdim=True)
    rstd = torch.rsqrt(xg.var(2, unbiased=False, keepdim=True) + eps)
    xhat = (xg - mean) * rstd
    g = (dy * gamma).reshape(R, groups, C // groups)
    dx = rstd * (g - g.mean(2, keepdim=True) - xhat * (g * xhat).mean(2, keepdim=True))
    return dx.reshape(R, C), (dy * xhat.reshape(R, C)).sum(0), dy.sum(0)


def lstm_cell_grad(gates, c_prev, dh, dc_next):
    """torch.nn.LSTMCell from its pre-activation gates [rows, 4H] (order i, f, g, o): given d loss / d h and the
    gradient arriving at c from the next consumer, returns (d gates, d c_prev)."""
    i, f, g, o = gates.chunk(4, 1)
    i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
    c = f * c_prev + i * g
    tc = torch.tanh(c)
    dc = dc_next + dh * o * (1 - tc * tc)
    d_gates = torch.cat([dc * g * i * (1 - i), dc * c_prev * f * (1 - f), dc * i * (1 - g * g), dh * tc * o * (1 - o)], 1)
    return d_gates, dc * f
