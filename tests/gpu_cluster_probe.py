"""A/B probe (not a test): GEMM cluster size (TMA multicast width) on the tree and the sequential rollout, through the
verification build's GCPB200_GEMM_CLUSTER switch.  One subprocess per setting (the switch is read at context creation)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from video_gcp_b200 import hparams
from video_gcp_b200.synthetic import synthetic_state_dict, synthetic_rollout_inputs, synthetic_seq_inputs
from verify_lib import verify_engine
dev = torch.device("cuda:0")
B = 1024
def run(kind):
    if kind == "tree":
        hp = hparams.build_hparams(hparams.gcp_tree_25room_config(batch_size=1))
        eng = verify_engine(dev, simt=False, max_candidates=B)
        inp = synthetic_rollout_inputs(B, seed=1, shared_images=True)
        fn = lambda: eng.rollout(I0, Ig, z, images_shared=True)
    else:
        hp = hparams.build_hparams(hparams.gcp_sequential_25room_config(batch_size=1))
        eng = verify_engine(dev, simt=False, max_candidates=B, model="sequential")
        inp = synthetic_seq_inputs(B, seed=1, shared_images=True)
        fn = lambda: eng.seq_rollout(I0, Ig, z, images_shared=True)
    eng.load_weights(synthetic_state_dict(hp, 1))
    I0, Ig, z = inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["z"].to(dev)
    for _ in range(3): fn()
    eng.profile_enable(True)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    p = eng.profile_read()
    print(kind, "cluster", sys.argv[1], {k: round(p[k] / 3, 3) for k in eng.PHASES})
    eng.close()
run("tree"); run("seq")
''' % (ROOT, ROOT)
for c in ("1", "2", "4", "8"):
    p = subprocess.run([sys.executable, "-c", CODE, c], env=dict(os.environ, GCPB200_GEMM_CLUSTER=c), capture_output=True, text=True)
    print(p.stdout.strip() or p.stderr[-1500:])
