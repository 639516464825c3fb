"""Generate the golden fixtures in tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden
TEST INFRASTRUCTURE.  Output: small .npz / .json files, committed, which pin `oracle/gcp_oracle.py`
and the product kernels to the reference's own behaviour (the reference ships no tests for this path).

Determinism: weights = video_gcp_b200.synthetic.synthetic_state_dict(seed) loaded with
load_state_dict(strict=True) into the reference model; inputs / noise seeded; the sampled rollout
length (base_gcp.py:223) is replaced by an injected `end_ind`.
"""
import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

refshim.install()
import torch  # noqa: E402
from blox import AttrDict as RefAttrDict  # noqa: E402
from gcp.prediction.models.tree.tree import TreeModel as RefTreeModel  # noqa: E402
from gcp.planning.cem.cem_simulator import GCPImageSimulator as RefSimulator  # noqa: E402
from gcp.planning.cem import cost_fcn as ref_cost  # noqa: E402
from gcp.planning.cem.sampler import FlatCEMSampler as RefSampler  # noqa: E402
from gcp.prediction.models.tree.frame_binding import BalancedBinding  # noqa: E402

from video_gcp_b200 import hparams as my_hparams  # noqa: E402
from video_gcp_b200 import spec as my_spec  # noqa: E402
from video_gcp_b200.synthetic import synthetic_state_dict, synthetic_rollout_inputs  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
WEIGHT_SEED = 1


def ref_config(**extra):
    from experiments.prediction.base_configs import gcp_tree as base_conf
    h = RefAttrDict(base_conf.model_config)
    h.update({
        'state_dim': 2, 'ngf': 16, 'max_seq_len': 200, 'hierarchy_levels': 8, 'nz_mid_lstm': 512,
        'n_lstm_layers': 3, 'nz_mid': 128, 'nz_enc': 128, 'nz_vae': 256, 'regress_length': True,
        'attach_state_regressor': True, 'attach_inv_mdl': True,
        'inv_mdl_params': RefAttrDict(n_actions=2, use_convs=False, build_encoder=False),
        'untied_layers': True, 'decoder_distribution': 'discrete_logistic_mixture', 'batch_size': 1,
    })
    h.pop("add_weighted_pixel_copy")
    h.update(extra)
    return h


def build_ref_model(**extra):
    with contextlib.redirect_stdout(io.StringIO()):
        m = RefTreeModel(ref_config(**extra), None)
    m.device = torch.device('cpu')
    m._hp.device = m.device
    m.eval()
    return m


def inject_end_ind(model, end_ind):
    orig = model.get_end_ind

    def patched(inputs, outputs):
        orig(inputs, outputs)                      # still runs the length predictor (seq_len_logits)
        outputs.end_ind = end_ind.clone()
        return outputs.end_ind

    model.get_end_ind = patched


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.manual_seed(0)
    np.random.seed(0)
    torch.set_num_threads(os.cpu_count())

    # ---------------- manifest of the reference state dicts (planner + training config)
    ref = build_ref_model()
    manifest = {k: list(v.shape) for k, v in ref.state_dict().items()}
    ref_train = build_ref_model(attach_cost_mdl=True,
                                cost_mdl_params=RefAttrDict(cost_fcn=ref_cost.EuclideanPathLength))
    manifest_train = {k: list(v.shape) for k, v in ref_train.state_dict().items()}
    del ref_train
    with open(os.path.join(GOLDEN, "state_dict_manifest.json"), "w") as f:
        json.dump({"planner": manifest, "training": manifest_train}, f, indent=0, sort_keys=True)

    # ---------------- load the synthetic weights into the reference
    hp = my_hparams.build_hparams(my_hparams.gcp_tree_25room_config(batch_size=1, attach_cost_mdl=True))
    sd_full = synthetic_state_dict(hp, WEIGHT_SEED)
    sd_planner = {k: v for k, v in sd_full.items() if not k.startswith("cost_mdl.")}
    ref.load_state_dict(sd_planner, strict=True)

    # ---------------- case A: model forward, B=2, distinct start/goal per candidate
    B = 2
    inp = synthetic_rollout_inputs(B, seed=3, shared_images=False)
    end_ind = torch.tensor([24, 199])
    inject_end_ind(ref, end_ind)
    inputs = RefAttrDict(I_0=inp["I_0"].clone(), I_g=inp["I_g"].clone(),
                         z=inp["z"].clone()[..., None, None],
                         start_ind=torch.zeros(B, dtype=torch.long),
                         end_ind=torch.full((B,), 199, dtype=torch.long))
    with torch.no_grad(), ref.val_mode():
        out = ref(inputs)
    tree = out.tree
    df = lambda name: tree.df[name]
    e_df = df("e_g_prime")[..., 0, 0]
    hid_df = df("hidden_state")
    pz = tree.df["p_z"] if False else None
    mu_df = torch.stack([n.subgoal.p_z.mu for n in tree.depth_first_iter()], 1)[..., 0, 0]
    ls_df = torch.stack([n.subgoal.p_z.log_sigma for n in tree.depth_first_iter()], 1)[..., 0, 0]
    images_df = df("images")
    feat_df = df("feat")
    hid_nodes = [127, 63, 191, 0, 254, 100]
    img_nodes = [0, 1, 63, 127, 128, 200, 254]
    sims = [ref.dense_rec.get_sample_with_len(i, int(end_ind[i]) + 1, out, inputs, 'basic')[0] for i in range(B)]
    np.savez_compressed(
        os.path.join(GOLDEN, "tree_forward_B2.npz"),
        weight_seed=WEIGHT_SEED, input_seed=3, end_ind=end_ind.numpy(),
        e0=inputs.e_0[..., 0, 0].numpy(), eg=inputs.e_g[..., 0, 0].numpy(),
        skip0=inputs.skips[0].numpy(), skip2=inputs.skips[2].numpy(),
        seq_len_logits=out.seq_len_logits.numpy(),
        e_df=e_df.numpy(), mu_df=mu_df.numpy(), log_sigma_df=ls_df.numpy(),
        hid_nodes=np.array(hid_nodes), hidden_sel=hid_df[:, hid_nodes].numpy(),
        img_nodes=np.array(img_nodes), images_sel=images_df[:, img_nodes].numpy(),
        images_sum=images_df.double().sum((2, 3, 4)).numpy(),
        images_abs=images_df.double().abs().sum((2, 3, 4)).numpy(),
        images_f16=images_df.numpy().astype(np.float16),
        feat_root=feat_df[:, 127].numpy(),
        existence=out.existence_predictor.existence.numpy(),
        actions=out.actions.numpy(), regressed_state=out.regressed_state.numpy(),
        model_enc_seq=inputs.model_enc_seq.numpy(),
        pruned_len=np.array([s.shape[0] for s in sims]),
        pruned0=sims[0].numpy(),
    )
    print("case A done; pruned lens", [s.shape[0] for s in sims])

    # ---------------- case B: balanced pruning masks for every end_ind (integer path)
    keep = np.zeros((200, 255), dtype=bool)
    tsteps = np.zeros((200, 255), dtype=np.int64)
    binding = ref.tree_module.tree_modules[0].binding
    for e in range(1, 200):
        l, r = torch.tensor([-1]), torch.tensor([e + 1])

        def rec(l, r, lvl, j):
            if lvl == 8:
                return
            o = BalancedBinding.__call__(binding, {}, None, RefAttrDict(timesteps=l), RefAttrDict(timesteps=r))
            i = (2 * j + 1) * 2 ** (7 - lvl) - 1
            keep[e, i] = bool(o.c_n_prime.bool().any())
            tsteps[e, i] = int(o.timesteps)
            rec(l, o.timesteps, lvl + 1, 2 * j)
            rec(o.timesteps, r, lvl + 1, 2 * j + 1)

        rec(l, r, 0, 0)
    np.savez_compressed(os.path.join(GOLDEN, "balanced_pruning.npz"), keep=keep, timesteps=tsteps)
    print("case B done; kept counts ok:", all(keep[e].sum() == e + 1 for e in range(1, 200)))

    # ---------------- case C: simulator + costs + elites + refit, N=12 candidates, shared start/goal
    N = 12
    r = np.random.default_rng(7)
    state = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    goal = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    samples = r.normal(0, 0.3, size=(N, 255, 256))
    end_c = torch.tensor(r.integers(2, 200, size=N))
    ref.get_end_ind = ref.__class__.get_end_ind.__get__(ref)
    inject_end_ind(ref, end_c)
    sim = RefSimulator(ref, append_latent=True)
    with torch.no_grad():
        ro = sim.rollout(state, goal, samples, 200)
    l2 = ref_cost.L2ImageCost(True, 1.0)
    l2_dense = l2(ro.predictions, goal)
    l2_last = ref_cost.L2ImageCost(False, 2.0)(ro.predictions, goal)
    # learned cost through the reference's own TestTimeCostModel (needs a checkpoint on disk)
    tmp = tempfile.mkdtemp()
    torch.save({'epoch': 0, 'global_step': 0, 'state_dict': sd_full, 'optimizer': {}},
               os.path.join(tmp, "weights_ep0.pth"))
    with contextlib.redirect_stdout(io.StringIO()):
        lc = ref_cost.ImageWrappedLearnedCostFcn(RefAttrDict(checkpt_path=tmp))
        with torch.no_grad():
            learned = lc(ro.predictions, None)
    order = l2_dense.argsort()
    k = int(N * 0.25)
    elite = order[:k]
    smp = RefSampler(float("inf"), 255, 256, 0.3)
    smp.fit(samples[elite], l2_dense[elite])
    np.savez_compressed(
        os.path.join(GOLDEN, "cem_N12.npz"),
        weight_seed=WEIGHT_SEED, rng_seed=7, state=state, goal=goal,
        end_ind=end_c.numpy(), l2_dense=l2_dense, l2_last_w2=l2_last, learned=learned,
        elite_idx=elite, fit_mean=smp.mean, fit_std=smp.std,
        pred_len=np.array([p.shape[0] for p in ro.predictions]),
        pred_sum=np.array([p.astype(np.float64).sum() for p in ro.predictions]),
        pred3=ro.predictions[3], act3=ro.actions[3], state3=ro.states[3], lat3=ro.latents[3],
    )
    print("case C done", l2_dense[:4], learned[:4])


if __name__ == "__main__":
    main()
