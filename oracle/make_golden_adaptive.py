"""Golden fixtures for the adaptive-binding GCP-tree rollout (config 4), made by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden_adaptive
TEST INFRASTRUCTURE.  Output: tests/golden/adaptive_forward_B2.npz and the `adaptive` entry of
state_dict_manifest.json.  Seeded synthetic weights loaded with load_state_dict(strict=True) into the reference
TreeModel built from base_configs/gcp_adaptive.py with the 25-room sizes; seeded inputs / noise.
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
from gcp.prediction.models.tree.tree import TreeModel as RefTreeModel  # noqa: E402

from video_gcp_b200 import hparams as my_hparams  # noqa: E402
from video_gcp_b200.synthetic import synthetic_state_dict, synthetic_rollout_inputs  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
WEIGHT_SEED = 3


def ref_config(**extra):
    from experiments.prediction.base_configs import gcp_adaptive as base_conf
    h = RefAttrDict(base_conf.model_config)
    h.update({'ngf': 16, 'max_seq_len': 200, 'hierarchy_levels': 8, 'nz_mid_lstm': 512, 'n_lstm_layers': 3,
              'nz_mid': 128, 'nz_enc': 128, 'nz_vae': 256, 'regress_length': True, 'untied_layers': True,
              'batch_size': 1})
    h.update(extra)
    return h


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    torch.set_num_threads(os.cpu_count())
    with contextlib.redirect_stdout(io.StringIO()):
        ref = RefTreeModel(ref_config(), None)
    ref.device = torch.device('cpu')
    ref._hp.device = ref.device
    ref.eval()
    mpath = os.path.join(GOLDEN, "state_dict_manifest.json")
    with open(mpath) as f:
        man = json.load(f)
    man["adaptive"] = {k: list(v.shape) for k, v in ref.state_dict().items()}
    with open(mpath, "w") as f:
        json.dump(man, f, indent=0, sort_keys=True)

    hp = my_hparams.build_hparams(my_hparams.gcp_adaptive_25room_config(batch_size=1))
    sd = synthetic_state_dict(hp, WEIGHT_SEED)
    ref.load_state_dict(sd, strict=True)

    B = 2
    inp = synthetic_rollout_inputs(B, seed=9, shared_images=False)
    inputs = RefAttrDict(I_0=inp["I_0"].clone(), I_g=inp["I_g"].clone(), z=inp["z"].clone()[..., None, None],
                         start_ind=torch.zeros(B, dtype=torch.long), end_ind=torch.full((B,), 199, dtype=torch.long))
    with torch.no_grad(), ref.val_mode():
        out = ref(inputs)
    tree = out.tree
    images_df = tree.df.images
    img_nodes = [0, 1, 63, 127, 128, 200, 254]
    np.savez_compressed(
        os.path.join(GOLDEN, "adaptive_forward_B2.npz"),
        weight_seed=WEIGHT_SEED, input_seed=9,
        e_df=tree.df.e_g_prime[..., 0, 0].numpy(),
        img_nodes=np.array(img_nodes), images_sel=images_df[:, img_nodes].numpy(),
        mask_sel=tree.df.pixel_copy_mask[:, img_nodes].numpy(),
        images_f16=images_df.numpy().astype(np.float16),
        images_sum=images_df.double().sum((2, 3, 4)).numpy(),
        distances=out.distance_predictor.distances.numpy(),
        pruned_len=np.array([p.shape[0] for p in out.pruned_prediction]),
        pruned_sum=np.array([float(p.double().sum()) for p in out.pruned_prediction]),
        pruned0=out.pruned_prediction[0].numpy().astype(np.float16),
        seq_len_logits=out.seq_len_logits.numpy(),
    )
    print("adaptive golden done; pruned lens", [p.shape[0] for p in out.pruned_prediction],
          "min |distance|", out.distance_predictor.distances.abs().min().item())


if __name__ == "__main__":
    main()
