#!/usr/bin/env python
"""bench.py -- tree-GCP CEM rollouts/sec on B200 (BASELINE.json metric, config 2).

A "step" is one pass of the CEM hot path over one batch of synthetic candidates of the 25-room shape:
batched GCP-tree rollout (encoder, sampled length, 8 TreeLSTM levels, all 255 nodes decoded, pruning,
inverse model / state regressor / existence heads) + dense L2 image cost + (N>1: one NCCL all-gather of the
costs) + elite top-k + refit.  `value` times it with the noise already resident in HBM; `e2e` goes through
the reference-facing simulator call with HOST (pinned) noise and start/goal images, copies inside the
timed region, and reads costs + elite indices back.

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference      (CPU restatement of the reference on the host cores)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tree-GCP CEM rollouts/sec"
UNIT = "rollouts/s"
FLOP_PER_ROLLOUT = 16.754e9          # canonical work, BASELINE.md section 3 (8.377 GMAC)
TAIL_FLOP_PER_IMAGE = 2 * (8.388608e6 + 7.8643e6)   # the two full-resolution decoder convolutions (canonical count)
# tensor-core FLOPs dec_tail3_kernel actually issues per image: 2 tiles x 2 convs x 28 tcgen05.mma of 128x64x16.  It is
# BELOW the canonical count because the encoder-skip half of conv 32->16 is a per-candidate constant computed once,
# and only the 15 mixture-mean channels of conv 16->30 reach the image (dec_tail3.cuh header) -- so the canonical
# `achieved` can exceed the tensor peak; `executed` is the hardware-side figure.
TAIL_EXECUTED_FLOP_PER_IMAGE = 2 * 2 * 28 * (2 * 128 * 64 * 16)
ELITE_FRAC = 0.1
# DRAM traffic of one decoder-tail launch at 1024 candidates: dram__bytes_read.sum + dram__bytes_write.sum of one
# `ncu --set full` capture, profiles/r1k_dec_tail3_ncu_full.txt.  With level-ordered decoding the four launches of a step
# decode 63, 64, 64 and 64 slots x 1024 candidates; the figure is their mean (63-slot launch 529.0 + 747.0 MB, 64-slot
# launches 537.4 + 758.7 MB).  Algorithmic I/O of a 64-slot launch is 536.9 MB read (bf16 layer-3 maps) + 805.3 MB written
# (fp32 images; the last 6 % drain from L2 after the kernel window): no re-reads.
TAIL_TRAFFIC_BYTES_B1024 = (528.97e6 + 746.97e6 + 3 * (537.36e6 + 758.71e6)) / 4


def workload(cands):
    return ("25-room gcp_tree CEM planning rollout, one start/goal pair, %d candidates per GPU, depth-8 tree, "
            "255 nodes decoded per candidate, sampled rollout length, dense L2 image cost, elite_frac 0.1" % cands)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.rows, self.proc, self.index = [], None, index

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:  # noqa: BLE001
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        busy = sorted(sm)[len(sm) // 4:] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return (d.get("bf16_tflops_sustained", 1371.4), d.get("hbm_gbs", 6553.3), "measured (MEASURED_PEAKS.json, sustained)",
                d.get("bf16_tflops", 1650.0))
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)", 1650.0


def oracle_step(sd, O, state, goal, n, seed):
    """One CEM iteration of the CPU restatement on n candidates; returns seconds."""
    r = np.random.default_rng(seed)
    samples = r.normal(0, 0.3, size=(n, 255, 256))
    end = r.integers(2, 200, size=n)
    t0 = time.perf_counter()
    with torch.no_grad():
        ro = O.simulator_rollout(sd, state, goal, samples, end)
    imgs = [p[:, :3072].reshape(-1, 3, 32, 32) for p in ro["predictions"]]
    cost = O.l2_image_cost(imgs, goal, True, 1.0)
    el = O.elites(cost, n, max(ELITE_FRAC, 1.0 / n))
    O.refit(samples, el)
    return time.perf_counter() - t0


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores (oracle port; the Python reference
    itself cannot travel to the GPU box).  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import gcp_oracle as O
    from video_gcp_b200 import hparams
    from video_gcp_b200.synthetic import synthetic_state_dict
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    hp = hparams.build_hparams(hparams.gcp_tree_25room_config(batch_size=1))
    sd = synthetic_state_dict(hp, 1)
    r = np.random.default_rng(0)
    state = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    goal = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    n = args.ref_sample
    for w in range(args.warmup):
        oracle_step(sd, O, state, goal, n, 100 + w)
    t = sum(oracle_step(sd, O, state, goal, n, 200 + s) for s in range(args.steps))
    v = n * args.steps / t
    sample = "%d of %d candidates per step (same per-candidate work)" % (n, args.candidates)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload(args.candidates), "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--candidates", type=int, default=1024, help="candidates per GPU")
    ap.add_argument("--ref-sample", type=int, default=16, help="candidates per CPU-reference step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from video_gcp_b200 import hparams
    from video_gcp_b200.model import TreeModel
    from video_gcp_b200.planning import GCPImageSimulator, L2ImageCost, SimpleTreeCEMSampler
    from video_gcp_b200.synthetic import synthetic_state_dict

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.candidates
    N = B * world
    k = max(int(N * ELITE_FRAC), 1)

    hp_cfg = hparams.gcp_tree_25room_config(batch_size=1)
    model = TreeModel(hp_cfg, None, max_candidates=B)
    model.load_state_dict(synthetic_state_dict(model._hp, 1), strict=True)
    model.device = dev
    model.eval()
    eng = model.engine
    sim = GCPImageSimulator(model, append_latent=False)
    cost_fcn = L2ImageCost(True, 1.0)
    sampler = SimpleTreeCEMSampler(float("inf"), 200, 256, 0.3, n_level_hierarchy=8).attach(eng, seed=7)

    r = np.random.default_rng(0)
    state = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    goal = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    # resident inputs for `value`; pinned host inputs for `e2e`
    z_dev = sampler.sample_device(B, first_id=rank * B)
    z_host = z_dev.cpu().pin_memory()
    state_t = torch.as_tensor(state).pin_memory()
    goal_t = torch.as_tensor(goal).pin_memory()

    def cem_tail(ro, z):
        cost_loc = cost_fcn.device_cost(ro)
        if world > 1:
            cost = torch.empty(N, device=dev, dtype=torch.float32)
            dist.all_gather_into_tensor(cost, cost_loc)
        else:
            cost = cost_loc
        idx, val = eng.topk(cost, k)
        if world > 1:
            # elites of other ranks are regenerated from the shared counter-based RNG (no payload)
            z_elite = sampler.regenerate(idx)
            mean, std = eng.refit(z_elite, torch.arange(k, device=dev, dtype=torch.int32))
        else:
            mean, std = eng.refit(z, idx)
        return cost, idx, val, mean

    def step_resident():
        ro = sim.rollout_device(state_t, goal_t, z_dev, 200)
        return cem_tail(ro, z_dev)

    def step_e2e():
        # the pinned host noise goes straight into the simulator call: the library uploads it level by level on
        # its copy stream (all 267 MB inside this step), overlapped with the encoder and the upper tree levels
        ro = sim.rollout_device(state_t, goal_t, z_host, 200)
        cost, idx, val, mean = cem_tail(ro, ro.z)
        return cost.cpu(), idx.cpu()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    l0 = eng.launch_count()
    ms = timed(step_resident, args.steps, args.warmup)
    launches = (eng.launch_count() - l0) * args.steps // (args.steps + args.warmup)
    ms_e2e = timed(step_e2e, args.steps, max(args.warmup, 3))
    clk = clocks.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel (decoder tail conv), timed live with CUDA events on its stream
    eng.profile_enable(True)
    for _ in range(2):
        step_resident()
    torch.cuda.synchronize()
    prof = eng.profile_read()
    eng.profile_enable(False)
    peak_tf, peak_hbm, peak_src, peak_burst = measured_peaks()
    tail_ms = prof["decoder_tail"] / max(prof["tail_launches"], 1)
    tail_imgs = prof["tail_images"] / max(prof["tail_launches"], 1)
    tail_tf = TAIL_FLOP_PER_IMAGE * tail_imgs / (tail_ms * 1e-3) / 1e12 if tail_ms > 0 else 0.0

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = N * args.steps / (ms * 1e-3)
    e2e_value = N * args.steps / (ms_e2e * 1e-3)
    phase_ms = {p: round(prof[p] / 2, 3) for p in eng.PHASES}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload(B), "candidates_total": N, "elites": k,
                   "l2": "inputs larger than L2 (noise z = %.0f MB read per step; images written %.1f GB)"
                         % (B * 255 * 256 * 4 / 1e6, B * 255 * 3072 * 4 / 1e9),
                   "parallelism": "candidates sharded over %d rank(s), cost all-gather" % world},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(z_host.numel() * 4 + 2 * 3072 * 4),
                "d2h_bytes_per_step": int(N * 4 + k * 4), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "dec_tail3_kernel", "achieved": tail_tf, "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": tail_tf / peak_tf,
                     "executed": {"achieved": tail_tf * TAIL_EXECUTED_FLOP_PER_IMAGE / TAIL_FLOP_PER_IMAGE,
                                  "frac": tail_tf * TAIL_EXECUTED_FLOP_PER_IMAGE / TAIL_FLOP_PER_IMAGE / peak_tf,
                                  "frac_of_burst": tail_tf * TAIL_EXECUTED_FLOP_PER_IMAGE / TAIL_FLOP_PER_IMAGE / peak_burst,
                                  "note": "tcgen05 FLOPs actually issued (29.36 MFLOP/image vs 32.51 canonical: shared "
                                          "skip half + unused mixture-scale channels are not computed)"},
                     "traffic": TAIL_TRAFFIC_BYTES_B1024 if B == 1024 else None,
                     "traffic_note": "bytes per launch, ncu dram read+write (profiles/r1k_dec_tail3_ncu_full.txt)",
                     "peak_source": peak_src,
                     "ms_per_launch": tail_ms, "images_per_launch": tail_imgs},
        "roofline_step": {"bound": "tensor", "achieved": value / world * FLOP_PER_ROLLOUT / 1e12, "peak": peak_tf,
                          "unit": "TFLOP/s per GPU (canonical 16.75 GFLOP/rollout)",
                          "frac": value / world * FLOP_PER_ROLLOUT / 1e12 / peak_tf},
        "phase_ms_per_step": phase_ms,
    }
    if world == 1 and not args.no_cpu_baseline:
        from oracle import gcp_oracle as O
        cores = os.cpu_count()
        torch.set_num_threads(cores)
        sd = synthetic_state_dict(model._hp, 1)
        n = args.ref_sample
        oracle_step(sd, O, state, goal, 4, 1)
        t = sum(oracle_step(sd, O, state, goal, n, 10 + s) for s in range(3))
        out["cpu_baseline"] = {"value": 3 * n / t, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": "3 steps of %d candidates (of %d), same per-candidate work" % (n, B)}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
