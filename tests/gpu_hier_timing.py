"""Latency of the hierarchical planning call (10 -> 10 -> 5 candidates + final rollout) and of the
closed-loop execution step (8f-3) on the B200, with the CPU oracle timed beside them.  Measurement aid, not a test:
    python tests/gpu_hier_timing.py  > profiles/<round>_hier_latency.txt"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gcp_oracle as O  # noqa: E402
from oracle import hier_oracle as H  # noqa: E402
from video_gcp_b200 import hparams  # noqa: E402
from video_gcp_b200.model import TreeModel  # noqa: E402
from video_gcp_b200.planning import (GCPImageSimulator, HierarchicalImageCEMPlanner, ImageHierarchicalTreeCEMSampler,  # noqa: E402
                                     ImageLearnedCostEstimate)
from video_gcp_b200.synthetic import synthetic_state_dict  # noqa: E402
from video_gcp_b200.types import AttrDict  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    cfg = hparams.gcp_tree_25room_config(batch_size=1, attach_cost_mdl=True)
    model = TreeModel(cfg, None, max_candidates=128)
    sd = synthetic_state_dict(model._hp, 1)
    model.load_state_dict(sd, strict=True)
    model.device = dev
    model.eval()
    r = np.random.default_rng(0)
    state = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    goal = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    for rng in ("numpy", "device"):
        sim = GCPImageSimulator(model, append_latent=True)
        p = HierarchicalImageCEMPlanner(AttrDict(
            action_dim=256, n_iters=3, batch_size=10, n_level_hierarchy=8, sampler=ImageHierarchicalTreeCEMSampler,
            sampling_rates_per_layer=[10, 10], cost_fcn=ImageLearnedCostEstimate, cost_config=AttrDict(), max_seq_len=200,
            sampler_rng=rng), sim)
        for _ in range(3):
            p(state, goal)
        torch.cuda.synchronize()
        l0 = model.engine.launch_count()
        t0 = time.perf_counter()
        n = 10
        for _ in range(n):
            p(state, goal)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        print("hierarchical plan (3 iterations 10/10/5 + final rollout), rng=%-6s: %8.2f ms per planning call, %d kernel launches"
              % (rng, dt * 1e3, (model.engine.launch_count() - l0) // n))
    # closed-loop step
    eng = model.engine
    img = torch.as_tensor(r.uniform(-1, 1, size=(1, 3, 32, 32)).astype(np.float32)).to(dev)
    tgt = torch.as_tensor(r.standard_normal((1, 128)).astype(np.float32)).to(dev)
    for _ in range(10):
        eng.infer_action(img, tgt)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        a = eng.infer_action(img, tgt).cpu()
    dt = (time.perf_counter() - t0) / 200
    print("closed-loop step (encoder + inverse model, action read back): %8.1f us per call" % (dt * 1e6))
    # CPU oracle beside it
    torch.set_num_threads(os.cpu_count())
    calls = []

    def rollout_fn(samples):
        end = H.injected_end_ind(len(calls), samples.shape[0])
        calls.append(1)
        with torch.no_grad():
            return O.simulator_rollout(sd, state, goal, samples, end)["predictions"]

    np.random.seed(0)
    t0 = time.perf_counter()
    H.plan(sd, rollout_fn, goal)
    print("CPU oracle (port of the reference planner, %d threads): %8.2f ms per planning call" % (os.cpu_count(), (time.perf_counter() - t0) * 1e3))
    img_hwc = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    with torch.no_grad():
        O.infer_action(sd, img_hwc, tgt[0].cpu().numpy())
        t0 = time.perf_counter()
        for _ in range(50):
            O.infer_action(sd, img_hwc, tgt[0].cpu().numpy())
    print("CPU oracle closed-loop step: %8.1f us per call" % ((time.perf_counter() - t0) / 50 * 1e6))


if __name__ == "__main__":
    main()
