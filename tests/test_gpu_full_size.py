"""Full-size property tests (BASELINE.json configs 2 and 4: 1024 candidates on the balanced tree, 4096 on the
adaptive-binding tree), run with -m gpu on the B200.  The CPU oracle takes ~50 ms per candidate, so at these sizes the
checks are the size-independent properties of the path: a candidate's rollout does not depend on its batch (bit-exact
against a small batch that the oracle-parity tests cover), pinned-host noise == device noise (bit-exact; this is the
size at which the level-ordered decoder and the throttled upload actually overlap), repeatability, the cost of every
candidate recomputed from the returned frames, top-k order with ties, the refit moments, and the sortedness / length of
every pruned sequence.  torch on the device is only the checker here."""
import numpy as np
import pytest
import torch

from video_gcp_b200 import hparams
from video_gcp_b200.synthetic import synthetic_rollout_inputs, synthetic_state_dict

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need the B200 box"
    return torch.device("cuda:0")


def test_tree_rollout_1024_candidates(dev, sd):
    from video_gcp_b200.engine import Engine
    B, n_small = 1024, 24
    eng = Engine(dev, max_candidates=B, attach_cost_mdl=True)
    eng.load_weights(sd)
    inp = synthetic_rollout_inputs(B, seed=101, shared_images=True)
    inp["end_ind"][:4] = torch.tensor([2, 199, 3, 198])                  # shortest / longest rollouts are in the batch
    I0, Ig, ei = inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["end_ind"].to(dev)
    zd = inp["z"].to(dev)
    keys = ("e_df", "images_df", "actions", "regressed_state", "existence", "model_enc_seq")
    a = eng.rollout(I0, Ig, zd, end_ind=ei, images_shared=True, fresh=True)
    # ---- repeatability and host-noise == device-noise, bit for bit
    b = eng.rollout(I0, Ig, zd, end_ind=ei, images_shared=True, fresh=True)
    h = eng.rollout(I0, Ig, inp["z"].pin_memory(), end_ind=ei, images_shared=True, fresh=True)
    torch.cuda.synchronize()
    for k in keys:
        assert torch.equal(a[k], b[k]), "repeat " + k
        assert torch.equal(a[k], h[k]), "host noise " + k
    assert torch.equal(h["z"], zd)
    assert all(bool(torch.isfinite(a[k]).all()) for k in keys)
    # ---- candidate independence: the first candidates rolled out alone (non-shared images, batch of 24)
    s = eng.rollout(inp["I_0"][:n_small].to(dev), inp["I_g"][:n_small].to(dev), zd[:n_small].contiguous(),
                    end_ind=ei[:n_small].contiguous(), fresh=True)
    torch.cuda.synchronize()
    for k in keys:
        assert torch.equal(a[k][:n_small], s[k]), "independence " + k
    # ---- pruned sequences: frame t of candidate c is an in-order node, strictly increasing, exactly end_ind + 1 frames
    seq = eng.prune_gather(a["e_df"], a["end_ind"])
    lens = (seq.abs().sum(-1) > 0).sum(1)
    assert torch.equal(lens.cpu(), inp["end_ind"] + 1)
    # every kept latent is one of the candidate's node latents, in depth-first order
    c = 1                                                                 # end_ind 199: 200 kept frames of 255 nodes
    node_of = [int((a["e_df"][c] == seq[c, t]).all(-1).nonzero()[0]) for t in range(int(lens[c]))]
    assert all(x < y for x, y in zip(node_of, node_of[1:])) and len(node_of) == 200
    # ---- dense L2 cost of every candidate recomputed from the returned frames (cost_fcn.py:9-22,65-72)
    cost = eng.cost_l2(a["images_df"], a["end_ind"], Ig[0], True, 1.0)
    frames = eng.prune_gather(a["images_df"], a["end_ind"]).reshape(B, 200, 3 * 32 * 32)
    d = (frames - Ig[0].reshape(1, 1, -1)).double().pow(2).sum(-1).sqrt()
    mask = torch.arange(200, device=dev)[None] <= a["end_ind"][:, None]
    want = (d * mask).sum(1)
    assert float(((cost.double() - want).abs() / want).max()) < 1e-5
    # ---- elites: k lowest, ascending, ties by index; refit = moments of the elite noise (sampler.py:44-46)
    k = 102
    cost[7] = cost[3]                                                     # force a tie
    idx, val = eng.topk(cost, k)
    order = torch.argsort(cost, stable=True)[:k]
    assert idx.tolist() == order.tolist() and torch.equal(val, cost[order])
    mean, std = eng.refit(zd, idx)
    ze = zd[idx.long()].double()
    assert float((mean.double() - ze.mean(0)).abs().max()) < 1e-5
    assert float((std.double() - ze.std(0, unbiased=False)).abs().max() / ze.std(0, unbiased=False).max()) < 1e-5
    eng.close()


def test_adaptive_rollout_4096_candidates(dev):
    from video_gcp_b200.engine import Engine
    B, n_small = 4096, 16
    hp = hparams.build_hparams(hparams.gcp_adaptive_25room_config(batch_size=1))
    eng = Engine(dev, max_candidates=B, model="tree_adaptive")
    eng.load_weights(synthetic_state_dict(hp, 3))
    inp = synthetic_rollout_inputs(B, seed=103, shared_images=True)
    I0, Ig, ei = inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["end_ind"].to(dev)
    zd = inp["z"].to(dev)
    a = eng.rollout(I0, Ig, zd, end_ind=ei, images_shared=True, fresh=True)
    s = eng.rollout(inp["I_0"][:n_small].to(dev), inp["I_g"][:n_small].to(dev), zd[:n_small].contiguous(),
                    end_ind=ei[:n_small].contiguous(), fresh=True)
    torch.cuda.synchronize()
    for k in ("e_df", "images_df", "distances", "pruned_len"):
        assert torch.equal(a[k][:n_small], s[k]), "independence " + k
    live = torch.arange(255, device=dev)[None] < s["pruned_len"].long()[:, None]      # entries past pruned_len are unspecified
    assert torch.equal(a["pruned_nodes"][:n_small][live], s["pruned_nodes"][live])
    assert bool(torch.isfinite(a["images_df"]).all()) and float(a["images_df"].abs().max()) <= 1.0 + 1e-3
    n = a["pruned_len"].long()
    assert int(n.min()) >= 1 and int(n.max()) <= 255
    nodes = a["pruned_nodes"].long()
    valid = torch.arange(255, device=dev)[None] < n[:, None]
    inc = (nodes[:, 1:] > nodes[:, :-1]) | ~valid[:, 1:]
    assert bool(inc.all()) and bool((nodes[:, 0] == 0).all())              # kept nodes in order, node 0 always kept
    # the kept list is exactly what the device's own distance logits imply (adaptive.py:62-77)
    keep = torch.cat([torch.ones(B, 1, dtype=torch.bool, device=dev), ~(a["distances"] > 0.0)], 1)
    assert torch.equal(keep.sum(1), n)
    eng.close()


def test_fused_row_mlp_is_bit_identical_to_the_layerwise_path(dev):
    """mlp_fused_kernel keeps the rounding points of the four-launch row-MLP body (bf16 activations between layers), so a
    rollout with GCPB200_NO_FUSED_MLP=1 (read once per process, hence the subprocesses) must give the same bits; the same
    holds for programmatic dependent launch (GCPB200_NO_PDL=1), which only changes when kernels start.  The switches exist
    in the verification build of the library only (tests/verify_lib.py); the shipped library's result must equal them."""
    import hashlib
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, hashlib, torch\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')\n"
        "from video_gcp_b200 import hparams\n"
        "from video_gcp_b200.engine import Engine\n"
        "from verify_lib import verify_engine\n"
        "from video_gcp_b200.synthetic import synthetic_rollout_inputs, synthetic_state_dict\n"
        "dev = torch.device('cuda:0')\n"
        "hp = hparams.build_hparams(hparams.gcp_tree_25room_config(batch_size=1, attach_cost_mdl=True))\n"
        "eng = Engine(dev, max_candidates=384, attach_cost_mdl=True) if sys.argv[1] == 'shipped' else \\\n"
        "    verify_engine(dev, simt=False, max_candidates=384, attach_cost_mdl=True)\n"
        "eng.load_weights(synthetic_state_dict(hp, 1))\n"
        "inp = synthetic_rollout_inputs(300, seed=5, shared_images=True)\n"
        "out = eng.rollout(inp['I_0'][:1].to(dev), inp['I_g'][:1].to(dev), inp['z'].to(dev), end_ind=inp['end_ind'].to(dev),\n"
        "                  images_shared=True, want_prior=True)\n"
        "torch.cuda.synchronize()\n"
        "h = hashlib.sha256()\n"
        "for k in ('e_df', 'mu_df', 'log_sigma_df', 'images_df', 'existence', 'actions', 'regressed_state', 'seq_len_logits'):\n"
        "    h.update(out[k].cpu().numpy().tobytes())\n"
        "print('HASH', h.hexdigest(), eng.launch_count())\n" % (root, root))
    res = {}
    for name, env in (("shipped", {}), ("fused", {}), ("layerwise", {"GCPB200_NO_FUSED_MLP": "1"}), ("no_pdl", {"GCPB200_NO_PDL": "1"})):
        p = subprocess.run([sys.executable, "-c", code, name], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        line = [l for l in p.stdout.splitlines() if l.startswith("HASH")][0].split()
        res[name] = (line[1], int(line[2]))
    assert res["shipped"][0] == res["fused"][0] == res["layerwise"][0] == res["no_pdl"][0], res
    assert res["fused"][1] < res["layerwise"][1]            # and it really is the fused path that ran
