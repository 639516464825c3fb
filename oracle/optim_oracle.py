"""CPU oracle of the optimiser step of the reference's training loop (train.py:162).

TEST INFRASTRUCTURE ONLY (same rules as oracle/gcp_oracle.py).  Parity status: PINNED -- oracle/make_golden_optim.py runs
the UNMODIFIED reference optimiser (`get_clipped_optimizer(optimizer_type=RAdam | Adam, gradient_clip=...)`,
blox/torch/training.py:146-161, blox/torch/radam.py:8-80) on seeded parameters / gradients and stores the trajectories in
tests/golden/optim_steps.npz; tests/test_oracle_optim.py checks the functions below against them.

numpy float32 arithmetic in the operation order of the reference; per-step scalars in Python floats (double), rounded to
float32 where torch rounds a Python scalar entering a float32 op.
"""
import math

import numpy as np

f32 = np.float32


def clip_scale(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ (L2): scale = min(1, max_norm / (total_norm + 1e-6)); None -> no clipping."""
    if max_norm is None:
        return f32(1.0)
    total = math.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads))
    coef = max_norm / (total + 1e-6)
    return f32(min(coef, 1.0))


def radam_scalars(step, lr, beta1, beta2):
    """blox/torch/radam.py:56-68 -> (rectified?, step_size * lr)."""
    beta2_t = beta2 ** step
    n_sma_max = 2 / (1 - beta2) - 1
    n_sma = n_sma_max - 2 * step * beta2_t / (1 - beta2_t)
    if n_sma >= 5:
        step_size = math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_sma_max - 4) * (n_sma - 2) / n_sma * n_sma_max /
                              (n_sma_max - 2)) / (1 - beta1 ** step)
    else:
        step_size = 1.0 / (1 - beta1 ** step)
    return n_sma >= 5, step_size * lr


def radam_step(p, g, m, v, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """One RAdam update (blox/torch/radam.py:44-80).  float32 arrays in, new (p, m, v) out."""
    beta1, beta2 = betas
    v = v * f32(beta2) + f32(1 - beta2) * g * g
    m = m * f32(beta1) + f32(1 - beta1) * g
    rect, ss = radam_scalars(step, lr, beta1, beta2)
    if weight_decay != 0:
        p = p + f32(-weight_decay * lr) * p
    if rect:
        p = p + f32(-ss) * (m / (np.sqrt(v) + f32(eps)))
    else:
        p = p + f32(-ss) * m
    return p.astype(f32), m.astype(f32), v.astype(f32)


def adam_step(p, g, m, v, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """One torch.optim.Adam update (single-tensor path, amsgrad off)."""
    beta1, beta2 = betas
    if weight_decay != 0:
        g = g + f32(weight_decay) * p
    m = m + f32(1 - beta1) * (g - m)
    v = v * f32(beta2) + f32(1 - beta2) * g * g
    step_size = lr / (1 - beta1 ** step)
    sqrt_bc2 = math.sqrt(1 - beta2 ** step)
    p = p + f32(-step_size) * (m / (np.sqrt(v) / f32(sqrt_bc2) + f32(eps)))
    return p.astype(f32), m.astype(f32), v.astype(f32)


def run(kind, params, grads_per_step, lr, betas, eps, weight_decay, gradient_clip):
    """Trajectory of a list of float32 parameter arrays over len(grads_per_step) steps; returns the final params."""
    step_fn = radam_step if kind == "radam" else adam_step
    ps = [p.astype(f32) for p in params]
    ms = [np.zeros_like(p) for p in ps]
    vs = [np.zeros_like(p) for p in ps]
    for t, grads in enumerate(grads_per_step, 1):
        s = clip_scale(grads, gradient_clip)
        for i, g in enumerate(grads):
            ps[i], ms[i], vs[i] = step_fn(ps[i], (g.astype(f32) * s).astype(f32), ms[i], vs[i], t, lr, betas, eps, weight_decay)
    return ps, ms, vs
