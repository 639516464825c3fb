"""GPU tests of the product's flat CEM loop (video_gcp_b200/planning/cem_planner.py; reference
gcp/planning/cem/cem_planner.py:55-135, sampler.py:33-46): the planner class itself -- not a re-implementation -- against
the CPU oracle, at the config-2 size for the elite-set criterion of `north_star`, chunked, and sharded over NCCL.

Tolerances: L2 cost of a rollout 2e-3 relative to the largest cost (decoded frames differ by <= 5e-3 max-abs, bf16
tensor-core operands; observed ~2e-4); elite sets must be identical wherever the oracle's cost gap to the k-th cost
exceeds twice the observed cost error; refitted mean / std 1e-5 absolute when the elite sets agree.
"""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import gcp_oracle as O

pytestmark = pytest.mark.gpu

COST_TOL = 2e-3


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need the B200 box"
    return torch.device("cuda:0")


def _model(sd, dev, max_candidates):
    from video_gcp_b200 import hparams
    from video_gcp_b200.model import TreeModel
    model = TreeModel(hparams.gcp_tree_25room_config(batch_size=1, attach_cost_mdl=True), None, max_candidates=max_candidates)
    model.load_state_dict(sd, strict=True)
    model.device = dev
    model.eval()
    return model


def _planner(model, N, elite_frac, n_iters=1, max_rollout_bs=1024, seed=3):
    from video_gcp_b200.planning import GCPImageSimulator, ImageCEMPlanner, L2ImageCost, SimpleTreeCEMSampler
    from functools import partial
    sim = GCPImageSimulator(model, append_latent=False)
    return ImageCEMPlanner(dict(batch_size=N, n_iters=n_iters, elite_frac=elite_frac, cost_fcn=L2ImageCost, dense_cost=True,
                                final_step_cost_weight=1.0, sampler=partial(SimpleTreeCEMSampler, n_level_hierarchy=8),
                                max_seq_len=200, action_dim=256, initial_std=0.3, max_rollout_bs=max_rollout_bs, seed=seed), sim)


def _images(rng):
    return (rng.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32), rng.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32))


def _oracle_costs(sd, state, goal, samples, end):
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        ro = O.simulator_rollout(sd, state, goal, samples, end)
    imgs = [p[:, :3072].reshape(-1, 3, 32, 32) for p in ro["predictions"]]
    return O.l2_image_cost(imgs, goal, True, 1.0), ro


def _check_elites(cost_dev, idx_dev, cost_ref, k):
    """north_star: identical elite index sets wherever the cost gaps exceed the tolerance.  No silent skip: the cost
    error itself is asserted, then every candidate outside the 2*err band around the oracle's k-th cost is checked."""
    cost_dev, cost_ref = np.asarray(cost_dev, np.float64), np.asarray(cost_ref, np.float64)
    err = np.abs(cost_dev - cost_ref).max()
    assert err <= COST_TOL * np.abs(cost_ref).max(), (err, np.abs(cost_ref).max())
    order = np.argsort(cost_ref, kind="stable")
    kth, nxt = cost_ref[order[k - 1]], cost_ref[order[k]] if k < len(order) else np.inf
    must_in = np.nonzero(cost_ref < kth - 2 * err)[0]
    must_out = np.nonzero(cost_ref > nxt + 2 * err)[0]
    got = set(int(i) for i in idx_dev)
    assert len(got) == k
    assert set(must_in.tolist()) <= got, sorted(set(must_in.tolist()) - got)
    assert not (set(must_out.tolist()) & got), sorted(set(must_out.tolist()) & got)
    # the device's own order is ascending in its own costs, ties by index
    d = cost_dev[np.asarray(idx_dev)]
    assert (np.diff(d) >= 0).all()
    return err, len(must_in), len(must_out)


def test_planner_two_iterations_match_oracle(dev, sd):
    """ImageCEMPlanner.cem_iteration twice, with this iteration's candidates injected (host samples, as the reference
    simulator receives them) and the rollout lengths injected: costs, elites and the refitted distribution against
    O.simulator_rollout + O.l2_image_cost + O.elites + O.refit; the second iteration's candidates are drawn from the
    ORACLE's refit, so both sides see the same inputs."""
    N, frac = 12, 0.25
    k = int(N * frac)
    rng = np.random.default_rng(11)
    state, goal = _images(rng)
    model = _model(sd, dev, 128)
    pl = _planner(model, N, frac)
    pl._sampler.init()
    mean, std = np.zeros((255, 256)), 0.3 * np.ones((255, 256))
    for it in range(2):
        samples = (mean + std * rng.standard_normal(size=(N, 255, 256))).astype(np.float32)
        end = rng.integers(2, 200, size=N)
        model.inject_end_ind = torch.as_tensor(end)
        cost, idx, val, packed = pl.cem_iteration(state, goal, samples=torch.from_numpy(samples).pin_memory())
        torch.cuda.synchronize()
        c_ref, _ = _oracle_costs(sd, state, goal, samples, end)
        err, n_in, n_out = _check_elites(cost.cpu().numpy(), idx.cpu().numpy(), c_ref, k)
        el = O.elites(c_ref, N, frac)
        assert np.allclose(val.cpu().numpy(), c_ref[el], rtol=COST_TOL, atol=0)
        assert torch.equal(pl._elite_samples(packed, idx).cpu(), torch.from_numpy(samples)[idx.cpu().long()])
        if set(el.tolist()) == set(idx.cpu().tolist()):
            m_ref, s_ref = O.refit(samples.astype(np.float64), idx.cpu().numpy())
            pl._sampler.sync_host()
            d = pl._sampler.get_dists()
            assert np.abs(d.mean - m_ref).max() < 1e-5 and np.abs(d.std - s_ref).max() < 1e-5
            mean, std = m_ref, s_ref
        print("iteration %d: cost err %.2e (max cost %.1f), %d candidates certainly in / %d certainly out, elites %s"
              % (it, err, np.abs(c_ref).max(), n_in, n_out, idx.cpu().tolist()))
    model.engine.close()


def test_elite_set_1024_candidates_against_oracle(dev, sd):
    """BASELINE config 2 size: 1024 candidates, k = 102, dense L2 cost, through ImageCEMPlanner.cem_iteration; the CPU
    oracle rolls out the same 1024 candidates (the expensive half of this test: ~1 min on 16 cores)."""
    N, frac = 1024, 0.1
    k = int(N * frac)
    rng = np.random.default_rng(5)
    state, goal = _images(rng)
    model = _model(sd, dev, N)
    pl = _planner(model, N, frac)
    pl._sampler.init()
    samples = (0.3 * rng.standard_normal(size=(N, 255, 256))).astype(np.float32)
    end = rng.integers(2, 200, size=N)
    model.inject_end_ind = torch.as_tensor(end)
    cost, idx, val, packed = pl.cem_iteration(state, goal, samples=torch.from_numpy(samples).to(dev))
    torch.cuda.synchronize()
    c_ref, _ = _oracle_costs(sd, state, goal, samples, end)
    err, n_in, n_out = _check_elites(cost.cpu().numpy(), idx.cpu().numpy(), c_ref, k)
    same = len(set(O.elites(c_ref, N, frac).tolist()) & set(idx.cpu().tolist()))
    print("1024 candidates: cost err %.3e of max %.1f; %d certainly in, %d certainly out, %d / %d elites shared with the oracle"
          % (err, np.abs(c_ref).max(), n_in, n_out, same, k))
    assert n_in + n_out >= N - 16            # the band around the k-th cost is narrow: the check is not vacuous
    # refit on the device's elites against numpy on the same rows
    m_ref, s_ref = O.refit(samples.astype(np.float64), idx.cpu().numpy())
    pl._sampler.sync_host()
    d = pl._sampler.get_dists()
    assert np.abs(d.mean - m_ref).max() < 1e-6 and np.abs(d.std - s_ref).max() < 1e-6
    model.engine.close()


def test_chunked_planner_equals_single_rollout(dev, sd):
    """batch_size = 3 * max_rollout_bs: every chunk's costs land in their own slice (round-1 ADVICE: equal-sized chunks
    aliased one engine buffer), so costs / elites / refit are bit-identical to one 384-candidate rollout; a full planner
    call returns the plan of the best elite."""
    N = 384
    rng = np.random.default_rng(2)
    state, goal = _images(rng)
    model = _model(sd, dev, N)
    end = torch.as_tensor(rng.integers(2, 200, size=N))
    res = []
    for bs in (N, 128):
        pl = _planner(model, N, 0.1, max_rollout_bs=bs, seed=9)
        pl._sampler.init()
        outs = []
        for it in range(2):
            # lengths are injected per chunk position, so give every chunk of a call the lengths of its own candidates
            model.inject_end_ind = None
            cost = torch.empty(N, device=dev)
            z = pl._sampler.sample_device(N)
            for s, e in pl._chunks(N):
                model.inject_end_ind = end[s:e]
                ro = pl._simulator.rollout_device(state, goal, z[s:e], 200)
                pl._cost_fcn.device_cost(ro, out=cost[s:e])
            idx, val = model.engine.topk(cost, 38)
            pl._sampler.fit_device(z, idx)
            outs.append((cost.clone(), idx.clone(), pl._sampler._mean_d.clone(), pl._sampler._std_d.clone()))
        res.append(outs)
    for a, b in zip(*res):
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    assert len(set(res[0][0][0].tolist())) > N // 2         # costs are not one chunk repeated
    # the planner's own loop, chunked, sampled lengths: plan = best elite's pruned frames
    model.inject_end_ind = None
    pl = _planner(model, N, 0.1, n_iters=2, max_rollout_bs=128, seed=4)
    frames, actions, latents, score = pl(state, goal)
    assert frames.shape[1:] == (3072,) and latents.shape == (frames.shape[0], 128) and actions.shape[1] == 2
    assert np.isfinite(score) and np.isfinite(frames).all()
    d1 = pl._sampler.get_dists()
    assert d1.std.mean() < 0.3                              # two refits on 10 % elites contract the distribution
    model.engine.close()


def test_topk_and_refit_against_numpy(dev, sd):
    """gcpb200_topk = `scores.argsort()[:k]` with numpy's order rule (ascending, NaN last), ties by index; gcpb200_refit =
    np.mean / np.std over the elites (cem_planner.py:129-130, sampler.py:44-46), at the config-5 size (65 536 candidates)."""
    from video_gcp_b200.engine import Engine
    eng = Engine(dev, max_candidates=128)
    rng = np.random.default_rng(0)
    for N, k in ((1, 1), (7, 7), (1000, 100), (65536, 6553), (65536, 1)):
        c = rng.standard_normal(N).astype(np.float32)
        if N >= 1000:
            c[rng.integers(0, N, 50)] = np.nan                    # NaN sorts last
            c[rng.integers(0, N, 200)] = c[rng.integers(0, N, 200)]    # exact ties -> lower index first
            c[:3] = (-np.inf, np.inf, -0.0)
            c[3] = 0.0
        idx, val = eng.topk(torch.from_numpy(c).to(dev), k)
        want = np.lexsort((np.arange(N), np.where(np.isnan(c), np.inf, c), np.isnan(c)))[:k]
        assert np.array_equal(idx.cpu().numpy(), want), (N, k)
        assert np.array_equal(val.cpu().numpy(), c[want], equal_nan=True)
    c = np.full(300, np.nan, np.float32)
    c[[17, 250]] = (2.0, 1.0)
    assert eng.topk(torch.from_numpy(c).to(dev), 5)[0].tolist() == [250, 17, 0, 1, 2]
    # refit: k from 1 to 6553 rows gathered out of 8192 resident candidates
    z = torch.randn(8192, 255, 256, device=dev) * 0.3 + 0.05
    for k in (1, 3, 102, 6553):
        el = torch.from_numpy(rng.permutation(8192)[:k].astype(np.int32)).to(dev)
        mean, std = eng.refit(z, el)
        rows = z[el.long()].double()
        assert float((mean.double() - rows.mean(0)).abs().max()) < 1e-7
        assert float((std.double() - rows.std(0, unbiased=False)).abs().max()) < 1e-7
    same = z[:1].expand(64, 255, 256).contiguous()
    mean, std = eng.refit(same, torch.arange(64, dtype=torch.int32, device=dev))
    assert float(std.abs().max()) == 0.0 and torch.equal(mean, same[0])
    eng.close()


def test_sampled_rollout_length_distribution(dev, sd):
    """The sampled branch (end_ind = NULL, what the planner and bench.py run): OneHotCategorical(logits).sample() ->
    argmax -> clamp(min=2) (base_gcp.py:215-229) as Gumbel-max with a counter-based hash.  With shared start / goal images
    every candidate has the same logits, so 16 x 1024 draws are i.i.d.: chi-square against softmax(logits) with the
    mass of lengths 0, 1 moved to 2 (the clamp), and no draw below 2."""
    from video_gcp_b200.engine import Engine
    from video_gcp_b200.synthetic import synthetic_rollout_inputs
    eng = Engine(dev, max_candidates=1024, attach_cost_mdl=True)
    eng.load_weights(sd)
    inp = synthetic_rollout_inputs(1024, seed=3, shared_images=True)
    I0, Ig, z = inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["z"].to(dev)
    draws, logits = [], None
    for seed in range(16):
        out = eng.rollout(I0, Ig, z, end_ind=None, seed=seed, images_shared=True, want_images=False, want_aux=False,
                          want_existence=False)
        draws.append(out["end_ind"].cpu().numpy().copy())
        logits = out["seq_len_logits"][0].double().cpu()
    draws = np.concatenate(draws)
    assert draws.min() >= 2 and draws.max() <= 199
    p = torch.softmax(logits, 0).numpy()
    p[2] += p[0] + p[1]
    p[:2] = 0
    exp = p * len(draws)
    obs = np.bincount(draws, minlength=200).astype(np.float64)
    # merge bins with small expectation (left to right) so every chi-square cell expects >= 8 draws
    cells_o, cells_e, o, e = [], [], 0.0, 0.0
    for b in range(2, 200):
        o, e = o + obs[b], e + exp[b]
        if e >= 8:
            cells_o.append(o), cells_e.append(e)
            o = e = 0.0
    if e > 0:
        cells_o[-1] += o
        cells_e[-1] += e
    co, ce = np.array(cells_o), np.array(cells_e)
    chi2 = float(((co - ce) ** 2 / ce).sum())
    dof = len(ce) - 1
    print("length sampling: %d draws, %d cells, chi2 %.1f (dof %d); clamp cell observed %d expected %.1f"
          % (len(draws), len(ce), chi2, dof, int(obs[2]), exp[2]))
    assert chi2 < dof + 5.0 * np.sqrt(2.0 * dof)
    assert len(np.unique(draws)) > 20          # not a degenerate sampler
    # different seeds give different draws, the same seed the same
    a = eng.rollout(I0, Ig, z, end_ind=None, seed=123, images_shared=True, want_images=False, want_aux=False, want_existence=False)["end_ind"].clone()
    b = eng.rollout(I0, Ig, z, end_ind=None, seed=123, images_shared=True, want_images=False, want_aux=False, want_existence=False)["end_ind"].clone()
    assert torch.equal(a, b) and not np.array_equal(a.cpu().numpy(), draws[:1024])
    eng.close()


# ---- (e) sharded over NCCL: needs 2 GPUs (skips on a single-GPU box; run with `gpurun --gpus 2`) --------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, N, q):
    import torch.distributed as dist
    from video_gcp_b200 import hparams
    from video_gcp_b200.synthetic import synthetic_state_dict
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=dev)
    sd = synthetic_state_dict(hparams.build_hparams(hparams.gcp_tree_25room_config(batch_size=1, attach_cost_mdl=True)), 1)
    model = _model(sd, dev, N // world)
    pl = _planner(model, N, 0.1, seed=21)
    pl._sampler.init()
    rng = np.random.default_rng(4)
    state, goal = _images(rng)
    trace = []
    for it in range(2):
        cost, idx, val, packed = pl.cem_iteration(state, goal)
        trace.append((cost.cpu(), idx.cpu(), pl._elite_samples(packed, idx).cpu(), pl._sampler._mean_d.cpu(), pl._sampler._std_d.cpu()))
    q.put((rank, trace))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_planner_nccl_two_ranks(dev, sd):
    """Two ranks over NCCL run ImageCEMPlanner.cem_iteration twice on 256 candidates (128 per rank): every rank ends with
    identical costs / elite ids / elite samples / mean / std, and these equal ONE GPU rolling out the same 256 global
    candidate ids (the counter-based noise stream is keyed by global id; sampled lengths are keyed by the rollout seed and
    the local candidate index, so the single-GPU run rolls the two halves out as two chunks with the ranks' seeds)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    N = 256
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, N, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
    for a, b in zip(res[0], res[1]):
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    # one GPU, same global ids: chunks of 128 = the ranks' slices; every chunk is a rollout call with the same model seed
    # sequence as a rank saw (seed 0 for iteration 0, 1 for iteration 1)
    model = _model(sd, dev, 128)
    pl = _planner(model, N, 0.1, max_rollout_bs=128, seed=21)
    pl._sampler.init()
    rng = np.random.default_rng(4)
    state, goal = _images(rng)
    for it in range(2):
        z = pl._sampler.sample_device(N)
        cost = torch.empty(N, device=dev)
        for s, e in pl._chunks(N):
            model.seed = it
            ro = pl._simulator.rollout_device(state, goal, z[s:e], 200, **pl._planner_mode(images=False, l2_out=cost[s:e]))
            pl._cost_fcn.device_cost(ro, out=cost[s:e])
        idx, val = model.engine.topk(cost, int(N * 0.1))
        pl._sampler.fit_device(z, idx)
        assert torch.equal(cost.cpu(), res[0][it][0]) and torch.equal(idx.cpu(), res[0][it][1])
        assert torch.equal(z[idx.long()].cpu(), res[0][it][2])
        assert torch.equal(pl._sampler._mean_d.cpu(), res[0][it][3]) and torch.equal(pl._sampler._std_d.cpu(), res[0][it][4])
    model.engine.close()


# ---- planner mode: decode only what balanced pruning keeps, L2 cost folded into the decoder tail ------------------------
@pytest.mark.parametrize("shared", [True, False])
def test_prune_before_decode_is_bit_identical(dev, sd, shared):
    """gcpb200_rollout_io.decode_kept_only / l2_cost against the full decode (all 255 nodes): the pruned frames, the
    fused cost and everything else are the same BITS -- a row of the decoder GEMMs and an image of the tail kernel depend
    only on their own latent row, wherever that row sits -- and the fused cost equals the image-reading cost kernel to
    summation order.  300 candidates (not a tile multiple), lengths from 1 to 199, shared and per-candidate start/goal."""
    from video_gcp_b200.engine import Engine
    from video_gcp_b200.synthetic import synthetic_rollout_inputs
    B = 300
    eng = Engine(dev, max_candidates=384, attach_cost_mdl=True)
    eng.load_weights(sd)
    inp = synthetic_rollout_inputs(B, seed=8, shared_images=shared)
    inp["end_ind"][:6] = torch.tensor([1, 2, 199, 3, 127, 128])
    nimg = 1 if shared else B
    I0, Ig, z, ei = inp["I_0"][:nimg].to(dev), inp["I_g"][:nimg].to(dev), inp["z"].to(dev), inp["end_ind"].to(dev)
    goal = Ig[0]
    kw = dict(end_ind=ei, images_shared=shared, fresh=True)
    full = eng.rollout(I0, Ig, z, **kw)
    cost_kernel = eng.cost_l2(full["images_df"], full["end_ind"], goal, True, 0.7).clone()
    frames_full = eng.prune_gather(full["images_df"], full["end_ind"])
    full_fused = eng.rollout(I0, Ig, z, l2_goal=goal, l2_dense=True, l2_final_step_weight=0.7, **kw)
    assert torch.equal(full_fused["images_df"], full["images_df"])
    # kept-only, images requested: the persistent image buffer keeps its sentinel wherever a node is pruned away
    sentinel = eng._buf("images_df", (B, 255, 3, 32, 32)).fill_(7.0)
    kept = eng.rollout(I0, Ig, z, end_ind=ei, images_shared=shared, decode_kept_only=True, l2_goal=goal, l2_dense=True,
                       l2_final_step_weight=0.7)
    assert kept["images_df"].data_ptr() == sentinel.data_ptr()
    frames_kept = eng.prune_gather(kept["images_df"], kept["end_ind"])
    assert torch.equal(frames_kept, frames_full)
    from video_gcp_b200.pruning import frame_nodes
    for c in (0, 1, 2, 4, 5, 17, B - 1):
        keep = torch.zeros(255, dtype=torch.bool)
        keep[torch.as_tensor(frame_nodes(int(inp["end_ind"][c])))] = True
        img = kept["images_df"][c].cpu()
        assert torch.equal(img[keep], full["images_df"][c].cpu()[keep])
        assert bool((img[~keep] == 7.0).all())
    cost_kept = kept["l2_cost"].clone()
    # kept-only without images: nothing but the cost leaves the decoder
    nocopy = eng.rollout(I0, Ig, z, end_ind=ei, images_shared=shared, decode_kept_only=True, want_images=False, l2_goal=goal,
                         l2_dense=True, l2_final_step_weight=0.7, fresh=True)
    assert "images_df" not in nocopy
    assert torch.equal(nocopy["l2_cost"], cost_kept) and torch.equal(full_fused["l2_cost"], cost_kept)
    assert float(((cost_kept - cost_kernel).abs() / cost_kernel.abs()).max()) < 2e-6
    last = eng.rollout(I0, Ig, z, end_ind=ei, images_shared=shared, decode_kept_only=True, want_images=False, l2_goal=goal,
                       l2_dense=False, l2_final_step_weight=1.0, fresh=True)["l2_cost"]
    assert float(((last - eng.cost_l2(full["images_df"], full["end_ind"], goal, False, 1.0)).abs() / last.abs()).max()) < 2e-6
    for k in ("e_df", "existence", "actions", "regressed_state", "model_enc_seq", "seq_len_logits"):
        assert torch.equal(nocopy[k], full[k]), k
    eng.close()


def test_planner_pruned_mode_equals_full_decode(dev, sd):
    """ImageCEMPlanner with prune_before_decode True / False: same candidates and rollout seeds -> identical elite ids,
    refit and plan (frames, actions, latents), two iterations + the final elite rollout; sampled lengths."""
    N = 256
    rng = np.random.default_rng(6)
    state, goal = _images(rng)
    model = _model(sd, dev, N)
    res = []
    for pruned in (False, True):
        from functools import partial
        from video_gcp_b200.planning import GCPImageSimulator, ImageCEMPlanner, L2ImageCost, SimpleTreeCEMSampler
        pl = ImageCEMPlanner(dict(batch_size=N, n_iters=2, elite_frac=0.1, cost_fcn=L2ImageCost, dense_cost=True,
                                  final_step_cost_weight=1.0, sampler=partial(SimpleTreeCEMSampler, n_level_hierarchy=8),
                                  max_seq_len=200, action_dim=256, initial_std=0.3, max_rollout_bs=128, seed=13,
                                  prune_before_decode=pruned, sort_lengths=True), GCPImageSimulator(model, append_latent=True))
        model.seed = 50
        frames, actions, latents, score = pl(state, goal)
        res.append((frames.copy(), actions.copy(), latents.copy(), score, pl._sampler.get_dists()))
    a, b = res
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert abs(a[3] - b[3]) <= 2e-6 * abs(a[3])
    assert np.array_equal(a[4].mean, b[4].mean) and np.array_equal(a[4].std, b[4].std)
    model.engine.close()


@pytest.mark.parametrize("order", ["sorted", "random"])
def test_tree_pruning_is_bit_identical(dev, sd, order):
    """gcpb200_rollout_io.tree_kept_only: the recursion visits only the (node, candidate tile) pairs some candidate keeps.
    Everything the planner reads -- fused cost, pruned frames, pruned latent sequence, actions, states -- is the same bits
    as with the whole tree computed, and the latents of kept nodes are identical; with lengths in descending order the deep
    levels' work lists are short (checked through the launch-independent row counts the library keeps), with random
    lengths the lists are long but the result is the same."""
    from video_gcp_b200.engine import Engine
    from video_gcp_b200.pruning import frame_nodes
    from video_gcp_b200.synthetic import synthetic_rollout_inputs
    B = 300
    eng = Engine(dev, max_candidates=384, attach_cost_mdl=True)
    eng.load_weights(sd)
    inp = synthetic_rollout_inputs(B, seed=9, shared_images=True)
    end = np.random.default_rng(4).integers(1, 200, size=B)
    end[:3] = (199, 128, 127)
    if order == "sorted":
        end = np.sort(end)[::-1].copy()
    ei = torch.as_tensor(end).to(dev)
    I0, Ig, z = inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["z"].to(dev)
    kw = dict(end_ind=ei, images_shared=True, decode_kept_only=True, l2_goal=Ig[0], want_existence=False, fresh=True)
    a = eng.rollout(I0, Ig, z, **kw)
    b = eng.rollout(I0, Ig, z, tree_kept_only=True, **kw)
    assert torch.equal(a["l2_cost"], b["l2_cost"])
    for k in ("model_enc_seq", "actions", "regressed_state", "seq_len_logits"):
        assert torch.equal(a[k], b[k]), k
    assert torch.equal(eng.prune_gather(a["images_df"], a["end_ind"]), eng.prune_gather(b["images_df"], b["end_ind"]))
    for c in (0, 1, 2, 150, B - 1):
        nodes = torch.as_tensor(frame_nodes(int(end[c])))
        assert torch.equal(a["e_df"][c].cpu()[nodes], b["e_df"][c].cpu()[nodes])
    eng.close()


def test_sorted_sampled_lengths(dev, sd):
    """sort_sampled_lengths: the same draws as without it (same seed), handed out in descending order."""
    from video_gcp_b200.engine import Engine
    from video_gcp_b200.synthetic import synthetic_rollout_inputs
    eng = Engine(dev, max_candidates=1024, attach_cost_mdl=True)
    eng.load_weights(sd)
    inp = synthetic_rollout_inputs(1000, seed=3, shared_images=True)
    I0, Ig, z = inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["z"].to(dev)
    kw = dict(end_ind=None, seed=77, images_shared=True, want_images=False, want_aux=False, want_existence=False)
    plain = eng.rollout(I0, Ig, z, **kw)["end_ind"].cpu().numpy().copy()
    srt = eng.rollout(I0, Ig, z, sort_sampled_lengths=True, **kw)["end_ind"].cpu().numpy().copy()
    assert np.array_equal(srt, np.sort(plain)[::-1]) and (np.diff(srt) <= 0).all() and srt.min() >= 2
    eng.close()
