"""Golden fixtures for the sequential GCP rollout (config 3), made by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden_seq
TEST INFRASTRUCTURE.  Output: tests/golden/seq_forward_B2.npz, seq_sim_N6.npz and the `sequential` entry of
state_dict_manifest.json.  Same determinism rules as oracle/make_golden.py: seeded synthetic weights loaded with
load_state_dict(strict=True) into the reference SequentialModel, seeded inputs / noise, injected rollout length.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

refshim.install()
import torch  # noqa: E402
from blox import AttrDict as RefAttrDict  # noqa: E402
from gcp.prediction.models.sequential import SequentialModel as RefSequentialModel  # noqa: E402
from gcp.planning.cem.cem_simulator import GCPImageSimulator as RefSimulator  # noqa: E402
from gcp.planning.cem import cost_fcn as ref_cost  # noqa: E402

from video_gcp_b200 import hparams as my_hparams  # noqa: E402
from video_gcp_b200.synthetic import synthetic_state_dict, synthetic_seq_inputs  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
WEIGHT_SEED = 2


def ref_config(**extra):
    from experiments.prediction.base_configs import gcp_sequential as base_conf
    h = RefAttrDict(base_conf.model_config)
    h.update({
        'state_dim': 2, 'ngf': 16, 'max_seq_len': 200, 'nz_mid_lstm': 1024, 'n_lstm_layers': 3, 'nz_mid': 128,
        'nz_enc': 128, 'nz_vae': 256, 'regress_length': True, 'attach_state_regressor': True, 'attach_inv_mdl': True,
        'inv_mdl_params': RefAttrDict(n_actions=2, use_convs=False, build_encoder=False),
        'decoder_distribution': 'discrete_logistic_mixture', 'batch_size': 1,
    })
    h.pop("add_weighted_pixel_copy")
    h.update(extra)
    return h


def inject_end_ind(model, end_ind):
    orig = model.get_end_ind

    def patched(inputs, outputs):
        orig(inputs, outputs)
        outputs.end_ind = end_ind.clone()
        return outputs.end_ind

    model.get_end_ind = patched


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    torch.set_num_threads(os.cpu_count())
    with contextlib.redirect_stdout(io.StringIO()):
        ref = RefSequentialModel(ref_config(), None)
    ref.device = torch.device('cpu')
    ref._hp.device = ref.device
    ref.eval()
    mpath = os.path.join(GOLDEN, "state_dict_manifest.json")
    with open(mpath) as f:
        man = json.load(f)
    man["sequential"] = {k: list(v.shape) for k, v in ref.state_dict().items()}
    with open(mpath, "w") as f:
        json.dump(man, f, indent=0, sort_keys=True)

    hp = my_hparams.build_hparams(my_hparams.gcp_sequential_25room_config(batch_size=1))
    sd = synthetic_state_dict(hp, WEIGHT_SEED)
    ref.load_state_dict(sd, strict=True)

    # ---------------- case A: model forward, B=2, distinct start/goal per candidate
    B = 2
    inp = synthetic_seq_inputs(B, seed=5, shared_images=False)
    end_pred = torch.tensor([37, 199])
    inject_end_ind(ref, end_pred)
    inputs = RefAttrDict(I_0=inp["I_0"].clone(), I_g=inp["I_g"].clone(), z=inp["z"].clone()[..., None, None],
                         start_ind=torch.zeros(B, dtype=torch.long), end_ind=torch.full((B,), 199, dtype=torch.long))
    with torch.no_grad(), ref.val_mode():
        out = ref(inputs)
    dr = out.dense_rec
    img_t = [0, 1, 2, 50, 120, 199]
    np.savez_compressed(
        os.path.join(GOLDEN, "seq_forward_B2.npz"),
        weight_seed=WEIGHT_SEED, input_seed=5, end_pred=end_pred.numpy(),
        e0=inputs.e_0[..., 0, 0].numpy(), eg=inputs.e_g[..., 0, 0].numpy(),
        seq_len_logits=out.seq_len_logits.numpy(),
        encodings=dr.encodings[..., 0, 0].numpy(),
        mu=dr.p_z.mu[..., 0, 0].numpy(), log_sigma=dr.p_z.log_sigma[..., 0, 0].numpy(),
        img_t=np.array(img_t), images_sel=dr.images[:, img_t].numpy(),
        images_f16=dr.images.numpy().astype(np.float16),
        images_sum=dr.images.double().sum((2, 3, 4)).numpy(),
        actions=out.actions.numpy(), regressed_state=out.regressed_state.numpy(),
        model_enc_seq=inputs.model_enc_seq.numpy(),
    )
    print("case A done", dr.encodings.abs().max().item(), dr.images.shape)

    # ---------------- case B: simulator + L2 cost, N=6 candidates, shared start/goal
    N = 6
    r = np.random.default_rng(11)
    state = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    goal = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    samples = r.normal(0, 1.0, size=(N, 199, 256))
    end_c = torch.tensor(r.integers(2, 200, size=N))
    ref.get_end_ind = ref.__class__.get_end_ind.__get__(ref)
    inject_end_ind(ref, end_c)
    sim = RefSimulator(ref, append_latent=True)
    with torch.no_grad():
        ro = sim.rollout(state, goal, samples, 200)
    l2_dense = ref_cost.L2ImageCost(True, 1.0)(ro.predictions, goal)
    np.savez_compressed(
        os.path.join(GOLDEN, "seq_sim_N6.npz"),
        weight_seed=WEIGHT_SEED, rng_seed=11, state=state, goal=goal, end_ind=end_c.numpy(), l2_dense=l2_dense,
        pred_len=np.array([p.shape[0] for p in ro.predictions]),
        pred_sum=np.array([p.astype(np.float64).sum() for p in ro.predictions]),
        pred2=ro.predictions[2].astype(np.float16), lat2=ro.latents[2], act2=ro.actions[2], state2=ro.states[2],
    )
    print("case B done", l2_dense)


if __name__ == "__main__":
    main()
