"""Throughput of the adaptive-binding rollout (BASELINE config 4: 4096 candidates on one B200), CPU oracle timed beside
it.  Measurement aid, not a test:   python tests/gpu_adaptive_timing.py > profiles/<round>_adaptive_rollout.txt"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gcp_oracle as O  # noqa: E402
from video_gcp_b200 import hparams  # noqa: E402
from video_gcp_b200.engine import Engine  # noqa: E402
from video_gcp_b200.synthetic import synthetic_rollout_inputs, synthetic_state_dict  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    hp = hparams.build_hparams(hparams.gcp_adaptive_25room_config(batch_size=1))
    sd = synthetic_state_dict(hp, 3)
    for B in (1024, 4096):
        eng = Engine(dev, max_candidates=B, model="tree_adaptive")
        eng.load_weights(sd)
        inp = synthetic_rollout_inputs(B, seed=103, shared_images=True)
        I0, Ig, ei, z = inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["end_ind"].to(dev), inp["z"].to(dev)

        def step():
            out = eng.rollout(I0, Ig, z, end_ind=ei, images_shared=True)
            cost = eng.cost_l2_nodes(out["images_df"], out["pruned_nodes"], out["pruned_len"], Ig[0], True, 1.0)
            return out, cost
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        n0 = eng.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out, cost = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        eng.profile_enable(True)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        pr = eng.profile_read()
        print("B=%4d  adaptive rollout + node-list L2 cost: %.2f ms/step (%.0f rollouts/s, %d launches), kept nodes per candidate "
              "%.1f of 255" % (B, ms, B / ms * 1e3, (eng.launch_count() - n0) // 5, float(out["pruned_len"].float().mean())))
        print("        phases (ms):", {k: round(v / 2, 3) for k, v in pr.items() if isinstance(v, float)})
        eng.close()
        del eng, out, cost
        torch.cuda.empty_cache()
    torch.set_num_threads(os.cpu_count())
    inp = synthetic_rollout_inputs(8, seed=103, shared_images=False)
    with torch.no_grad():
        O.adaptive_rollout(sd, inp["I_0"][:2], inp["I_g"][:2], inp["z"][:2])
        t0 = time.perf_counter()
        O.adaptive_rollout(sd, inp["I_0"], inp["I_g"], inp["z"])
        cpu = time.perf_counter() - t0
    print("CPU oracle port (%d threads): %.0f ms for 8 candidates = %.1f rollouts/s" % (os.cpu_count(), cpu * 1e3, 8 / cpu))


if __name__ == "__main__":
    main()
