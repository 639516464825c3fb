"""Measurement aid: overlapped host-noise upload vs explicit copy (B=1024)."""
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

def t(fn, n=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

print("copy only        : %.2f ms" % t(lambda: zh.to(dev, non_blocking=True)))
print("rollout, dev z   : %.2f ms" % t(lambda: eng.rollout(I0, Ig, zd, end_ind=ei, images_shared=True)))
print("rollout, host z  : %.2f ms" % t(lambda: eng.rollout(I0, Ig, zh, end_ind=ei, images_shared=True)))
print("copy + rollout   : %.2f ms" % t(lambda: eng.rollout(I0, Ig, zh.to(dev, non_blocking=True), end_ind=ei, images_shared=True)))
print("tree only, dev z : %.2f ms" % t(lambda: eng.rollout(I0, Ig, zd, end_ind=ei, images_shared=True, want_images=False, want_aux=False, want_existence=False)))
print("tree only, host z: %.2f ms" % t(lambda: eng.rollout(I0, Ig, zh, end_ind=ei, images_shared=True, want_images=False, want_aux=False, want_existence=False)))
