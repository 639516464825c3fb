"""Golden gradients of the training step (groundwork for the backward pass), produced by RUNNING THE UNMODIFIED
REFERENCE:  python -m oracle.make_golden_train_grad     (build container only)

TEST INFRASTRUCTURE.  Same setup as oracle/make_golden_train.py case A (B = 2, seeded weights / batch / posterior noise /
np.random draws), but with autograd on: `losses.total.value.backward()` as train.py:155-160 does, then for every
parameter the gradient's L2 norm, sum and first 8 entries go to tests/golden/train_grads_B2.npz (a few kB).
"""
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

from oracle.make_golden import GOLDEN, WEIGHT_SEED  # noqa: E402
from oracle.make_golden_train import EpsQueue, build_train_model, parse_draws  # noqa: E402
from video_gcp_b200 import hparams as my_hparams  # noqa: E402
from video_gcp_b200.synthetic import synthetic_state_dict, synthetic_train_batch  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count())
    hp = my_hparams.build_hparams(my_hparams.gcp_tree_25room_config(batch_size=1, attach_cost_mdl=True))
    sd = synthetic_state_dict(hp, WEIGHT_SEED)
    B = 2
    model = build_train_model(B)
    model.load_state_dict(sd, strict=True)
    batch = synthetic_train_batch(B, seed=5, end_ind=[61, 198])
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
        inputs = RefAttrDict(
            traj_seq=batch["traj_seq"].clone(), traj_seq_images=batch["traj_seq"].clone(),
            pad_mask=batch["pad_mask"].clone(), end_ind=batch["end_ind"].clone(),
            start_ind=torch.zeros(B, dtype=torch.long), traj_seq_states=batch["states"].clone(),
            actions=batch["actions"].clone(), I_0=batch["I_0"].clone(), I_g=batch["I_g"].clone())
        model.zero_grad()
        out = model(inputs)
        losses = model.loss(inputs, out)
        losses.total = model.get_total_loss(inputs, losses)
        losses.total.value.backward()
    finally:
        ref_dist.Gaussian.sample = orig_sample
        np.random.randint = orig_randint
    t0, t1, cs, ce = parse_draws(draws, B)
    names, norm, gsum, head = [], [], [], []
    seen = set()
    for k, p in model.named_parameters():
        if p.grad is None or id(p) in seen:
            continue
        seen.add(id(p))
        g = p.grad.detach().double().reshape(-1)
        names.append(k)
        norm.append(float(g.norm()))
        gsum.append(float(g.sum()))
        h = np.zeros(8)
        h[:min(8, g.numel())] = g[:8].numpy()
        head.append(h)
    print("parameters with a gradient:", len(names), " total loss", float(losses.total.value))
    np.savez_compressed(os.path.join(GOLDEN, "train_grads_B2.npz"), names=np.array(names), norm=np.array(norm),
                        sum=np.array(gsum), head=np.array(head), total=float(losses.total.value),
                        inv_t0=t0, inv_t1=t1, cost_start=cs, cost_end=ce, cost_target=out.cost_target.detach().numpy(),
                        end_ind=batch["end_ind"].numpy(), batch_seed=5, weight_seed=WEIGHT_SEED)


if __name__ == "__main__":
    main()
