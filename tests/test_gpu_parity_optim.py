"""GPU parity of the optimiser step (gcpb200_sq_norm / gcpb200_optim_step through video_gcp_b200.optim) against the
trajectories of the unmodified reference optimisers (tests/golden/optim_steps.npz) and the CPU oracle; fp32 elementwise
arithmetic: 5e-6 relative (FMA contraction on the device vs separate roundings in torch)."""
import os

import numpy as np
import pytest
import torch

from oracle import optim_oracle as OO
from oracle.make_golden_optim_data import CASES, data

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    assert torch.cuda.is_available(), "GPU tests need the B200 box"
    from video_gcp_b200.engine import Engine
    eng = Engine(torch.device("cuda:0"), max_candidates=128)
    yield eng
    eng.close()


@pytest.mark.parametrize("name", sorted(CASES))
def test_trajectory_matches_reference(engine, golden_dir, name):
    from video_gcp_b200.optim import get_clipped_optimizer
    g = np.load(os.path.join(golden_dir, "optim_steps.npz"))
    c = CASES[name]
    params, grads = data()
    dev = engine.device
    ps = [torch.nn.Parameter(torch.tensor(p, device=dev)) for p in params]
    opt = get_clipped_optimizer(ps, engine, optimizer_type=c["kind"], lr=c["lr"], betas=c["betas"],
                                weight_decay=c["weight_decay"], gradient_clip=c["clip"])
    for step_grads in grads:
        for p, gr in zip(ps, step_grads):
            p.grad = torch.tensor(gr, device=dev)
        opt.step()
    torch.cuda.synchronize()
    for i, p in enumerate(ps):
        st = opt.state[p]
        for tag, got in (("p", p.data), ("m", st["exp_avg"]), ("v", st["exp_avg_sq"])):
            ref = g["%s_%s%d" % (name, tag, i)]
            err = np.abs(got.cpu().numpy() - ref).max() / max(np.abs(ref).max(), 1e-12)
            assert err < 5e-6, (name, tag, i, err)
    # checkpoint round trip in torch.optim's format, then one more step equals the oracle's 13th step
    sd = opt.state_dict()
    opt2 = get_clipped_optimizer(ps, engine, optimizer_type=c["kind"], gradient_clip=c["clip"])
    opt2.load_state_dict(sd)
    assert opt2.param_groups[0]["lr"] == c["lr"] and opt2.state[ps[2]]["step"] == len(grads)
    before = [p.detach().cpu().numpy().copy() for p in ps]
    ms = [opt2.state[p]["exp_avg"].cpu().numpy().copy() for p in ps]
    vs = [opt2.state[p]["exp_avg_sq"].cpu().numpy().copy() for p in ps]
    for p, gr in zip(ps, grads[0]):
        p.grad = torch.tensor(gr, device=dev)
    opt2.step()
    torch.cuda.synchronize()
    s = OO.clip_scale(grads[0], c["clip"])
    fn = OO.radam_step if c["kind"] == "radam" else OO.adam_step
    for i, p in enumerate(ps):
        want, _, _ = fn(before[i], (grads[0][i] * s).astype(np.float32), ms[i], vs[i], len(grads) + 1, c["lr"], c["betas"], 1e-8,
                        c["weight_decay"])
        assert np.abs(p.detach().cpu().numpy() - want).max() / max(np.abs(want).max(), 1e-12) < 5e-6


def test_full_model_step_bandwidth(engine):
    """One RAdam step over a 124.24 M-parameter flat array (the 25-room GCP-tree): algorithmic 28 B per parameter."""
    from video_gcp_b200.optim import get_clipped_optimizer
    dev = engine.device
    n = 124_240_000
    p = torch.nn.Parameter(torch.randn(n, device=dev))
    opt = get_clipped_optimizer([p], engine, optimizer_type="radam", gradient_clip=1.0)
    p.grad = torch.randn(n, device=dev)
    for _ in range(3):
        opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gbs = n * (28 + 4) / ms / 1e6          # + 4 B: the gradient is read once more for the norm
    print("RAdam + clipping over %d parameters: %.3f ms per step = %.0f GB/s algorithmic" % (n, ms, gbs))
    assert torch.isfinite(p).all() and gbs > 2000


def test_step_then_forward_uses_new_weights(sd):
    """the engine computes with its own packed copy of the weights, so a forward after `optimizer.step()` must
    see the update.  `ClippedOptimizer(model=...)` marks the model dirty; the next `model.engine` access repacks.  The
    rollout after the step equals the rollout of a fresh model loaded with the updated state dict, and differs from the
    rollout before it."""
    from video_gcp_b200 import hparams
    from video_gcp_b200.model import TreeModel
    from video_gcp_b200.optim import get_clipped_optimizer
    from video_gcp_b200.synthetic import synthetic_rollout_inputs
    dev = torch.device("cuda:0")
    cfg = hparams.gcp_tree_25room_config(batch_size=1, attach_cost_mdl=True)

    def build(state):
        m = TreeModel(cfg, None, max_candidates=128)
        m.load_state_dict(state, strict=True)
        m.to(dev)
        m.device = dev
        m.eval()
        return m

    inp = synthetic_rollout_inputs(8, seed=3, shared_images=True)
    I0, Ig, z, ei = inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["z"].to(dev), inp["end_ind"].to(dev)
    run = lambda m: {k: v.clone() for k, v in m.engine.rollout(I0, Ig, z, end_ind=ei, images_shared=True).items()
                     if k in ("e_df", "images_df", "actions")}
    model = build(sd)
    before = run(model)
    names = ["decoder.net.gen_head.conv.weight", "tree_module.tree_modules.0.subgoal_pred.lstm.0.weight_ih", "encoder.net.net.input.conv.weight"]
    params = dict(model.named_parameters())
    ps = [params[n] for n in names]
    opt = get_clipped_optimizer(ps, model.engine, optimizer_type="adam", lr=5e-2, gradient_clip=None, model=model)
    g = torch.Generator(device="cpu").manual_seed(0)
    for p in ps:
        p.grad = torch.randn(p.shape, generator=g).to(dev)
    opt.step()
    torch.cuda.synchronize()
    assert model._dirty
    after = run(model)
    assert not model._dirty
    fresh = run(build({k: v.detach().cpu() for k, v in model.state_dict().items()}))
    for k in after:
        assert torch.equal(after[k], fresh[k]), k
        assert not torch.equal(after[k], before[k]), k
