"""Throughput of the training-phase forward + loss (BASELINE config 1: batch 16, T = 200) on the B200, with the CPU
oracle port timed beside it.  Measurement aid, not a test:
    python tests/gpu_train_timing.py > profiles/<round>_train_forward.txt"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import train_oracle as TO  # noqa: E402
from video_gcp_b200 import hparams  # noqa: E402
from video_gcp_b200.engine import Engine  # noqa: E402
from video_gcp_b200.synthetic import synthetic_state_dict, synthetic_train_batch  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    hp = hparams.build_hparams(hparams.gcp_tree_25room_config(batch_size=1, attach_cost_mdl=True))
    sd = synthetic_state_dict(hp, 1)
    eng = Engine(dev, max_candidates=128, attach_cost_mdl=True)
    eng.load_weights(sd)
    for B in (16, 64, 128):
        batch = synthetic_train_batch(B, seed=7)
        r = np.random.default_rng(3)
        ei = batch["end_ind"].numpy()
        aux = dict(inv_t0=r.integers(0, ei), inv_t1=r.integers(0, ei) + 1, cost_start=np.zeros(B, np.int64), cost_end=ei)
        aux["inv_t1"] = np.minimum(aux["inv_t0"] + 1, ei)
        d = {k: v.to(dev) for k, v in batch.items() if isinstance(v, torch.Tensor)}
        pin = {k: v.pin_memory() for k, v in batch.items() if isinstance(v, torch.Tensor)}

        def step(src):
            return eng.forward_loss(src["traj_seq"], src["pad_mask"], src["end_ind"], src["states"], src["actions"],
                                    src["eps"], aux["inv_t0"], aux["inv_t1"], aux["cost_start"], aux["cost_end"],
                                    I_0=src["I_0"], I_g=src["I_g"])
        for _ in range(3):
            step(d)
        torch.cuda.synchronize()
        n0 = eng.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            out = step(d)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        launches = (eng.launch_count() - n0) // 10
        # end to end: pinned host batch -> device -> losses back on the host
        t0 = time.perf_counter()
        for _ in range(5):
            out = step({k: v.to(dev, non_blocking=True) for k, v in pin.items()})
            losses = out["losses"].cpu()
        e2e = (time.perf_counter() - t0) / 5 * 1e3
        print("B=%3d  device %.2f ms/step (%.0f sequences/s, %d launches)   e2e %.2f ms/step (%.0f sequences/s)   total loss %.6f"
              % (B, ms, B / ms * 1e3, launches, e2e, B / e2e * 1e3, float(losses[-1])))
        if B == 16:
            torch.set_num_threads(os.cpu_count())
            aux_o = dict(aux, cost_target=np.ones((B, 1), np.float32))
            with torch.no_grad():
                TO.forward_loss(sd, batch, aux_o)
                t0 = time.perf_counter()
                TO.forward_loss(sd, batch, aux_o)
                cpu = time.perf_counter() - t0
            print("B= 16  CPU oracle port (%d threads): %.0f ms/step (%.1f sequences/s)" % (os.cpu_count(), cpu * 1e3, B / cpu))
    eng.close()


if __name__ == "__main__":
    main()
