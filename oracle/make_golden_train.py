"""Golden fixtures for the training-phase forward + loss (BASELINE config 1), produced by
RUNNING THE UNMODIFIED REFERENCE:  python -m oracle.make_golden_train     (build container only)

TEST INFRASTRUCTURE.  The reference model (`TreeModel` with the 25-room prediction config,
experiments/prediction/25room/gcp_tree/conf.py:20-44) runs `model(inputs)` + `model.loss` +
`model.get_total_loss` exactly as `train.py:155-157` / `train.py:204-206` do, in `.train()` mode (batch-statistic
BatchNorm).  Three sources of randomness are pinned, none changes arithmetic:
  * posterior samples `q_z.sample()` (blox/torch/dist.py:246-247) draw their N(0,1) noise from an injected
    depth-first tensor eps [B,255,256] instead of `torch.randn_like`;
  * `np.random.randint` (inverse-model offsets inverse_mdl.py:88-98, cost-model pair cost_mdl.py:105-107) is
    seeded and its draws are recorded, so the oracle / device path can be given the same indices;
  * the cost target of the cost model comes from the reference's own `EuclideanPathLength` (recorded).
Weights are `synthetic_state_dict(seed 1)` loaded with strict=True.
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
import blox.torch.dist as ref_dist  # noqa: E402
from gcp.prediction.models.tree.tree import TreeModel as RefTreeModel  # noqa: E402
from gcp.planning.cem import cost_fcn as ref_cost  # noqa: E402

from oracle.make_golden import ref_config, GOLDEN, WEIGHT_SEED  # noqa: E402
from oracle.gcp_oracle import df_index, DEPTH  # noqa: E402
from video_gcp_b200 import hparams as my_hparams  # noqa: E402
from video_gcp_b200.synthetic import synthetic_state_dict, synthetic_train_batch  # noqa: E402


def build_train_model(B):
    with contextlib.redirect_stdout(io.StringIO()):
        m = RefTreeModel(ref_config(attach_cost_mdl=True,
                                    cost_mdl_params=RefAttrDict(cost_fcn=ref_cost.EuclideanPathLength),
                                    batch_size=B, n_actions=2), None)
    m.device = torch.device('cpu')
    m._hp.device = m.device
    m.train()                      # train.py:140; val() leaves it in train mode too (train.py:195)
    return m


class EpsQueue:
    """Feeds Gaussian.sample level by level from a depth-first eps tensor [B,255,256]."""

    def __init__(self, eps):
        self.eps, self.level = eps, 0

    def next(self, shape):
        B = self.eps.shape[0]
        n = 2 ** self.level
        assert shape[0] == B * n, (shape, self.level)
        idx = [df_index(self.level, j) for j in range(n)]
        self.level += 1
        return self.eps[:, idx].reshape(B * n, 256, 1, 1)


def run_reference(model, batch):
    """One training-phase forward + loss of the reference.  Returns (inputs, outputs, losses, draws)."""
    q = EpsQueue(batch["eps"])
    orig_sample = ref_dist.Gaussian.sample
    ref_dist.Gaussian.sample = lambda self: self.mu + self.sigma * q.next(self.mu.shape)
    draws = []
    orig_randint = np.random.randint

    def logged_randint(*a, **k):
        r = orig_randint(*a, **k)
        draws.append(np.asarray(r).reshape(-1).copy())
        return r

    np.random.randint = logged_randint
    np.random.seed(int(batch["np_seed"]))
    try:
        B = batch["traj_seq"].shape[0]
        inputs = RefAttrDict(
            traj_seq=batch["traj_seq"].clone(), traj_seq_images=batch["traj_seq"].clone(),
            pad_mask=batch["pad_mask"].clone(), end_ind=batch["end_ind"].clone(),
            start_ind=torch.zeros(B, dtype=torch.long), traj_seq_states=batch["states"].clone(),
            actions=batch["actions"].clone(), I_0=batch["I_0"].clone(), I_g=batch["I_g"].clone())
        with torch.no_grad():
            out = model(inputs)
            losses = model.loss(inputs, out)
            losses.total = model.get_total_loss(inputs, losses)
    finally:
        ref_dist.Gaussian.sample = orig_sample
        np.random.randint = orig_randint
    assert q.level == DEPTH
    return inputs, out, losses, draws


def parse_draws(draws, B):
    """np.random.randint call order inside run_auxilliary_models (base_gcp.py:250-260): inverse model first
    (B scalar t0 draws, then one vector of B delta_t, inverse_mdl.py:94-96), then the cost model (per sequence:
    start_idx, end_idx, cost_mdl.py:105-106)."""
    assert len(draws) == B + 1 + 2 * B, len(draws)
    t0 = np.array([int(d[0]) for d in draws[:B]])
    t1 = t0 + draws[B]
    cs = np.array([int(draws[B + 1 + 2 * b][0]) for b in range(B)])
    ce = np.array([int(draws[B + 2 + 2 * b][0]) for b in range(B)])
    return t0, t1, cs, ce


def loss_dict(losses):
    return {k: float(v.value) for k, v in losses.items()}


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    hp = my_hparams.build_hparams(my_hparams.gcp_tree_25room_config(batch_size=1, attach_cost_mdl=True))
    sd = synthetic_state_dict(hp, WEIGHT_SEED)

    # ---------------- case A: B = 2, full intermediates
    B = 2
    model = build_train_model(B)
    model.load_state_dict(sd, strict=True)
    batch = synthetic_train_batch(B, seed=5, end_ind=[61, 198])
    inputs, out, losses, draws = run_reference(model, batch)
    t0, t1, cs, ce = parse_draws(draws, B)
    tree = out.tree
    # depth_first_iter yields the LAYER object with `.subgoal` set to the current node: evaluate lazily
    stack = lambda f: torch.stack([f(n) for n in tree.depth_first_iter()], 1)
    L = loss_dict(losses)
    print("case A losses", L)
    img_nodes = [0, 63, 127, 128, 254]
    dmu = stack(lambda n: n.subgoal.distr.mu)          # [B,255,5,3,32,32]
    dls = stack(lambda n: n.subgoal.distr.log_sigma)
    np.savez_compressed(
        os.path.join(GOLDEN, "train_forward_B2.npz"),
        weight_seed=WEIGHT_SEED, batch_seed=5, end_ind=batch["end_ind"].numpy(), np_seed=int(batch["np_seed"]),
        inv_t0=t0, inv_t1=t1, cost_start=cs, cost_end=ce, cost_target=out.cost_target.numpy(),
        loss_names=np.array(sorted(L.keys())), loss_values=np.array([L[k] for k in sorted(L.keys())]),
        e0=inputs.e_0[..., 0, 0].numpy(), eg=inputs.e_g[..., 0, 0].numpy(),
        skip0=inputs.skips[0].numpy(), skip2=inputs.skips[2].numpy(),
        enc_traj_seq=inputs.enc_traj_seq[..., 0, 0].numpy(), inf_enc_seq=inputs.inf_enc_seq[..., 0, 0].numpy(),
        seq_len_logits=out.seq_len_logits.numpy(),
        e_df=tree.df.e_g_prime[..., 0, 0].numpy(),
        p_mu=stack(lambda n: n.subgoal.p_z.mu)[..., 0, 0].numpy(),
        p_log_sigma=stack(lambda n: n.subgoal.p_z.log_sigma)[..., 0, 0].numpy(),
        q_mu=stack(lambda n: n.subgoal.q_z.mu)[..., 0, 0].numpy(),
        q_log_sigma=stack(lambda n: n.subgoal.q_z.log_sigma)[..., 0, 0].numpy(),
        match_timesteps=stack(lambda n: n.subgoal.match_timesteps)[..., 0].numpy(),
        match_node=tree.bf.match_dist.argmax(1).numpy(),          # [B,T] breadth-first node index per frame
        kl_per_node=losses.kl.error_mat.sum((2, 3, 4)).numpy() if losses.kl.error_mat.dim() == 5
        else losses.kl.error_mat.reshape(B, 255, -1).sum(2).numpy(),                 # breadth-first
        nll_per_frame=losses.dense_img_rec.error_mat.sum((2, 3, 4)).numpy(),         # [B,T]
        img_nodes=np.array(img_nodes), images_sel=tree.df.images[:, img_nodes].numpy(),
        distr_mu_sel=dmu[:, img_nodes[:2]].numpy(), distr_ls_sel=dls[:, img_nodes[:2]].numpy(),
        images_sum=tree.df.images.double().sum((2, 3, 4)).numpy(),
        existence=out.existence_predictor.existence.numpy(),
        model_enc_seq=inputs.model_enc_seq.numpy(), regressed_state=out.regressed_state.numpy(),
        inv_actions=out.actions.numpy(), cost_pred=out.cost.numpy(),
    )

    # ---------------- case B: BASELINE config 1 (B = 16): losses + per-sequence reductions only
    B = 16
    model = build_train_model(B)
    model.load_state_dict(sd, strict=True)
    batch = synthetic_train_batch(B, seed=0)
    import time
    t = time.time()
    inputs, out, losses, draws = run_reference(model, batch)
    dt = time.time() - t
    t0, t1, cs, ce = parse_draws(draws, B)
    L = loss_dict(losses)
    print("case B (config 1, B=16) losses", L, "reference CPU forward+loss %.2f s on %d threads" % (dt, torch.get_num_threads()))
    np.savez_compressed(
        os.path.join(GOLDEN, "train_losses_B16.npz"),
        weight_seed=WEIGHT_SEED, batch_seed=0, end_ind=batch["end_ind"].numpy(), np_seed=int(batch["np_seed"]),
        inv_t0=t0, inv_t1=t1, cost_start=cs, cost_end=ce, cost_target=out.cost_target.numpy(),
        loss_names=np.array(sorted(L.keys())), loss_values=np.array([L[k] for k in sorted(L.keys())]),
        kl_per_seq=losses.kl.error_mat.reshape(B, -1).sum(1).numpy(),
        nll_per_seq=losses.dense_img_rec.error_mat.reshape(B, -1).sum(1).numpy(),
        e_df_abs=out.tree.df.e_g_prime[..., 0, 0].abs().mean((1, 2)).numpy(),
        ref_cpu_seconds=dt, ref_cpu_threads=torch.get_num_threads(),
    )

    # ---------------------------------------------------------------- case C: the extreme sequence lengths (end_ind = 1 is the
    # shortest the inverse / cost models accept, 199 the longest the data format holds): losses + per-sequence reductions
    B = 3
    model = build_train_model(B)
    model.load_state_dict(sd, strict=True)
    batch = synthetic_train_batch(B, seed=9, end_ind=[1, 199, 100])
    inputs, out, losses, draws = run_reference(model, batch)
    t0, t1, cs, ce = parse_draws(draws, B)
    L = loss_dict(losses)
    print("case C (end_ind 1 / 199 / 100) losses", L)
    np.savez_compressed(
        os.path.join(GOLDEN, "train_losses_edge_B3.npz"),
        weight_seed=WEIGHT_SEED, batch_seed=9, end_ind=batch["end_ind"].numpy(), np_seed=int(batch["np_seed"]),
        inv_t0=t0, inv_t1=t1, cost_start=cs, cost_end=ce, cost_target=out.cost_target.numpy(),
        loss_names=np.array(sorted(L.keys())), loss_values=np.array([L[k] for k in sorted(L.keys())]),
        kl_per_seq=losses.kl.error_mat.reshape(B, -1).sum(1).numpy(),
        nll_per_frame=losses.dense_img_rec.error_mat.sum((2, 3, 4)).numpy(),
        match_node=out.tree.bf.match_dist.argmax(1).numpy(),
    )


if __name__ == "__main__":
    main()
