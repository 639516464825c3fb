"""Measurement aid: where the host-noise (e2e) rollout loses time against the device-resident one (B=1024)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_gcp_b200 import hparams
from video_gcp_b200.engine import Engine
from video_gcp_b200.synthetic import synthetic_rollout_inputs, synthetic_state_dict

B = 1024
dev = torch.device("cuda:0")
hp = hparams.build_hparams(hparams.gcp_tree_25room_config(batch_size=1))
eng = Engine(dev, max_candidates=B); eng.load_weights(synthetic_state_dict(hp, 1))
inp = synthetic_rollout_inputs(B, seed=5, shared_images=True)
I0, Ig, ei = inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["end_ind"].to(dev)
zh = inp["z"].pin_memory(); zd = inp["z"].to(dev)

def t(fn, n=8, sync_each=False):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
        if sync_each: torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for name, z in (("dev z ", zd), ("host z", zh)):
    print("rollout, %s              : %.2f ms" % (name, t(lambda: eng.rollout(I0, Ig, z, end_ind=ei, images_shared=True))))
    print("rollout, %s, sync each   : %.2f ms" % (name, t(lambda: eng.rollout(I0, Ig, z, end_ind=ei, images_shared=True), sync_each=True)))
    eng.profile_enable(True)
    for _ in range(4): eng.rollout(I0, Ig, z, end_ind=ei, images_shared=True)
    torch.cuda.synchronize()
    print("   phases", {k: (round(v / 4, 3) if isinstance(v, float) else v) for k, v in eng.profile_read().items()})
    eng.profile_enable(False)
print("copy only                    : %.2f ms" % t(lambda: zh.to(dev, non_blocking=True)))
print("tree only, dev z             : %.2f ms" % t(lambda: eng.rollout(I0, Ig, zd, end_ind=ei, images_shared=True, want_images=False, want_aux=False, want_existence=False)))
print("tree only, host z            : %.2f ms" % t(lambda: eng.rollout(I0, Ig, zh, end_ind=ei, images_shared=True, want_images=False, want_aux=False, want_existence=False)))
