"""Back-to-back rollout timing without host syncs (measurement aid)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_gcp_b200 import hparams
from video_gcp_b200.engine import Engine
from video_gcp_b200.synthetic import synthetic_rollout_inputs, synthetic_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
hp = hparams.build_hparams(hparams.gcp_tree_25room_config(batch_size=1))
eng = Engine(dev, max_candidates=B); eng.load_weights(synthetic_state_dict(hp, 1))
inp = synthetic_rollout_inputs(B, seed=5, shared_images=True)
I0, Ig, z, ei = inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["z"].to(dev), inp["end_ind"].to(dev)
for mode in ("injected", "sampled"):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
    eng.rollout(I0, Ig, z, end_ind=ei if mode == "injected" else None, images_shared=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    evs[0].record()
    for i in range(8):
        eng.rollout(I0, Ig, z, end_ind=ei if mode == "injected" else None, seed=i, images_shared=True)
        evs[i + 1].record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(mode, "per-iter ms:", [round(evs[i].elapsed_time(evs[i + 1]), 2) for i in range(8)], "cpu enqueue ms/iter %.2f" % ((t1 - t0) * 1e3 / 8))
eng.profile_enable(True)
for i in range(4): eng.rollout(I0, Ig, z, end_ind=ei, images_shared=True)
torch.cuda.synchronize()
print({k: (round(v / 4, 3) if isinstance(v, float) else v) for k, v in eng.profile_read().items()})
