"""Golden fixture of the 9-room GCP-tree planner model (experiments/control/9room/gcp_tree/mod_hyper.py:33-54: 7 levels = 127
nodes, max_seq_len 100, `untied_layers` at its default False, i.e. ONE TreeModule for all levels), by RUNNING THE UNMODIFIED
REFERENCE.  Run in the build container only (needs /root/reference):   python -m oracle.make_golden_9room
TEST INFRASTRUCTURE.  Output: tests/golden/tree9room.npz (model forward B=2 with distinct images; simulator + L2 cost + elites
+ refit on 8 candidates), pinning oracle/gcp_oracle.py and the device path at the second tree shape.
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

refshim.install()
import torch  # noqa: E402
from blox import AttrDict as RefAttrDict  # noqa: E402
from gcp.planning.cem import cost_fcn as ref_cost  # noqa: E402
from gcp.planning.cem.cem_simulator import GCPImageSimulator as RefSimulator  # noqa: E402
from gcp.planning.cem.sampler import FlatCEMSampler as RefSampler  # noqa: E402
from gcp.prediction.models.tree.tree import TreeModel as RefTreeModel  # noqa: E402

from video_gcp_b200 import hparams as my_hparams  # noqa: E402
from video_gcp_b200.synthetic import synthetic_state_dict  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
WEIGHT_SEED = 2
DEPTH, N_NODES, MAX_LEN = 7, 127, 100


def ref_config():
    from experiments.prediction.base_configs import gcp_tree as base_conf
    h = RefAttrDict(base_conf.model_config)
    h.update({
        'state_dim': 2, 'ngf': 16, 'max_seq_len': MAX_LEN, 'hierarchy_levels': DEPTH, 'nz_mid_lstm': 512,
        'n_lstm_layers': 3, 'nz_mid': 128, 'nz_enc': 128, 'nz_vae': 256, 'regress_length': True,
        'attach_state_regressor': True, 'attach_inv_mdl': True,
        'inv_mdl_params': RefAttrDict(n_actions=2, use_convs=False, build_encoder=False),
        'decoder_distribution': 'discrete_logistic_mixture', 'batch_size': 1,
    })
    h.pop("add_weighted_pixel_copy")
    return h


def inject_end_ind(model, end_ind):
    orig = model.__class__.get_end_ind.__get__(model)

    def patched(inputs, outputs):
        orig(inputs, outputs)
        outputs.end_ind = end_ind.clone()
        return outputs.end_ind

    model.get_end_ind = patched


def inputs_9room(B, seed):
    r = np.random.default_rng([int(seed), 999])
    I_0 = r.uniform(-1, 1, size=(B, 3, 32, 32)).astype(np.float32)
    I_g = r.uniform(-1, 1, size=(B, 3, 32, 32)).astype(np.float32)
    z = r.standard_normal(size=(B, N_NODES, 256)).astype(np.float32)
    return torch.from_numpy(I_0), torch.from_numpy(I_g), torch.from_numpy(z)


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    torch.set_num_threads(os.cpu_count())
    with contextlib.redirect_stdout(io.StringIO()):
        ref = RefTreeModel(ref_config(), None)
    ref.device = torch.device('cpu')
    ref._hp.device = ref.device
    ref.eval()
    hp = my_hparams.build_hparams(my_hparams.gcp_tree_9room_config(batch_size=1))
    sd = synthetic_state_dict(hp, WEIGHT_SEED)
    assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    ref.load_state_dict(sd, strict=True)

    # ---- model forward, B = 2, distinct start / goal images, lengths 37 and 99
    B = 2
    I_0, I_g, z = inputs_9room(B, 5)
    end_ind = torch.tensor([37, 99])
    inject_end_ind(ref, end_ind)
    inputs = RefAttrDict(I_0=I_0.clone(), I_g=I_g.clone(), z=z.clone()[..., None, None],
                         start_ind=torch.zeros(B, dtype=torch.long), end_ind=torch.full((B,), MAX_LEN - 1, dtype=torch.long))
    with torch.no_grad(), ref.val_mode():
        out = ref(inputs)
    tree = out.tree
    e_df = tree.df["e_g_prime"][..., 0, 0]
    mu_df = torch.stack([n.subgoal.p_z.mu for n in tree.depth_first_iter()], 1)[..., 0, 0]
    images_df = tree.df["images"]
    img_nodes = [0, 1, 31, 63, 64, 100, 126]
    sims = [ref.dense_rec.get_sample_with_len(i, int(end_ind[i]) + 1, out, inputs, 'basic')[0] for i in range(B)]

    # ---- simulator + L2 cost + elites + refit, 8 candidates, shared start / goal
    N = 8
    r = np.random.default_rng(17)
    state = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    goal = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    samples = r.normal(0, 0.3, size=(N, N_NODES, 256))
    end_c = torch.tensor(r.integers(2, MAX_LEN, size=N))
    inject_end_ind(ref, end_c)
    sim = RefSimulator(ref, append_latent=True)
    with torch.no_grad():
        ro = sim.rollout(state, goal, samples, MAX_LEN)
    l2 = ref_cost.L2ImageCost(True, 1.0)(ro.predictions, goal)
    elite = l2.argsort()[:2]
    smp = RefSampler(float("inf"), N_NODES, 256, 0.3)
    smp.fit(samples[elite], l2[elite])
    np.savez_compressed(
        os.path.join(GOLDEN, "tree9room.npz"),
        weight_seed=WEIGHT_SEED, input_seed=5, end_ind=end_ind.numpy(),
        e0=inputs.e_0[..., 0, 0].numpy(), seq_len_logits=out.seq_len_logits.numpy(),
        e_df=e_df.numpy(), mu_df=mu_df.numpy(), img_nodes=np.array(img_nodes), images_sel=images_df[:, img_nodes].numpy(),
        images_f16=images_df.numpy().astype(np.float16), existence=out.existence_predictor.existence.numpy(),
        actions=out.actions.numpy(), regressed_state=out.regressed_state.numpy(), model_enc_seq=inputs.model_enc_seq.numpy(),
        pruned_len=np.array([s.shape[0] for s in sims]), pruned0=sims[0].numpy(),
        rng_seed=17, cem_end_ind=end_c.numpy(), l2_dense=l2, elite_idx=elite, fit_mean=smp.mean, fit_std=smp.std,
        pred_len=np.array([p.shape[0] for p in ro.predictions]), pred2=ro.predictions[2], act2=ro.actions[2], lat2=ro.latents[2],
    )
    print("9-room fixture written; pruned lens", [s.shape[0] for s in sims], "l2", l2[:3])


if __name__ == "__main__":
    main()
