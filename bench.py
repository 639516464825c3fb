#!/usr/bin/env python
"""bench.py -- tree-GCP CEM rollouts/sec on B200 (BASELINE.json metric, config 2).

A "step" is ONE ITERATION OF THE PRODUCT'S FLAT CEM LOOP, `ImageCEMPlanner.cem_iteration`
(video_gcp_b200/planning/cem_planner.py; reference gcp/planning/cem/cem_planner.py:58-69) over one batch of synthetic
candidates of the 25-room shape: draw the candidates (device Philox, keyed by global candidate id) -> batched GCP-tree
rollout (encoder, sampled length, 8 TreeLSTM levels, all 255 nodes decoded, pruning, inverse model / state regressor /
existence heads) -> dense L2 image cost -> (N>1: one NCCL all-gather of the costs) -> elite top-k -> refit.

  value            the step with everything resident in HBM (the two 12 KB start / goal images are the call's arguments)
  e2e              the same planner call from HOST start / goal images (pinned), costs + elite ids read back to the host
  e2e_host_noise   the reference simulator's contract: this step's 267 MB of candidates come from pinned HOST memory
                   (`simulator.rollout(state, goal, samples, ...)`, cem_simulator.py:14), copied inside the timed region
  e2e_planner      a whole `planner(state, goal)` call: n_iters iterations + final rollout of the elites + plan to host
  value_pruned     planner mode "decode only what the cost reads" (kept nodes only, L2 cost folded into the decoder
                   tail); identical costs / elites, fewer executed FLOPs -- a second figure, `value` stays canonical

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --gpus N --candidates-total 65536        (strong scaling: config 5's sweep, sharded over N ranks)
  python bench.py --config seq                             (BASELINE config 3: sequential GCP rollout)
  python bench.py --impl reference      (the UNMODIFIED reference staged under baseline/_ref, on the host cores)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from functools import partial

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tree-GCP CEM rollouts/sec"
UNIT = "rollouts/s"
FLOP_PER_ROLLOUT = 16.754e9          # canonical work, BASELINE.md section 3 (8.377 GMAC)
SEQ_FLOP_PER_ROLLOUT = 2 * 9.76e9    # sequential GCP: 199 steps x (prior MLP + 3 x 1024-wide LSTM) + decoder
TAIL_FLOP_PER_IMAGE = 2 * (8.388608e6 + 7.8643e6)   # the two full-resolution decoder convolutions (canonical count)
# tensor-core FLOPs dec_tail3_kernel actually issues per image: 2 tiles x 2 convs x 28 tcgen05.mma of 128x64x16.  It is
# BELOW the canonical count because the encoder-skip half of conv 32->16 is a per-candidate constant computed once,
# and only the 15 mixture-mean channels of conv 16->30 reach the image (dec_tail3.cuh header) -- so the canonical
# `achieved` can exceed the tensor peak; `executed` is the hardware-side figure.
TAIL_EXECUTED_FLOP_PER_IMAGE = 2 * 2 * 28 * (2 * 128 * 64 * 16)
ELITE_FRAC = 0.1
# DRAM traffic of one decoder-tail launch at 1024 candidates: dram__bytes_read.sum + dram__bytes_write.sum of one
# `ncu --set full` capture, profiles/r2o_dec_tail3_ncu_full.txt (round 2, tail kernel with the fused L2 epilogue).  With
# level-ordered decoding the four launches of a step decode 63, 64, 64 and 64 slots x 1024 candidates; the figure is their
# mean (63-slot launch 529.6 + 746.8 MB, 64-slot launches 538.1 + 757.8 MB).  Algorithmic I/O of a 64-slot launch is 536.9 MB
# read (bf16 layer-3 maps) + 805.3 MB written (fp32 images; the last 6 % drain from L2 after the kernel window): no re-reads.
TAIL_TRAFFIC_BYTES_B1024 = (529.59e6 + 746.79e6 + 3 * (538.09e6 + 757.79e6)) / 4


def workload(cands, config="tree"):
    if config == "seq":
        return ("25-room sequential GCP (blox vrnn) rollout, one start/goal pair, %d candidates per GPU, 199 recurrent steps, "
                "200 frames decoded per candidate, sampled rollout length, dense L2 image cost, elite_frac 0.1" % cands)
    return ("25-room gcp_tree CEM planning rollout, one start/goal pair, %d candidates per GPU, depth-8 tree, "
            "255 nodes decoded per candidate, sampled rollout length, dense L2 image cost, elite_frac 0.1" % cands)


class ClockSampler(threading.Thread):
    """SM clock / power / throttle reasons during the timed region (B200_PROFILING.md recipe).  Read in-process through
    NVML (the library nvidia-smi itself queries), initialised BEFORE the warm-up: spawning `nvidia-smi -lms` next to the
    timed loop attaches a second client to the GPU, and on a run where its start-up landed inside the K timed steps the
    launches of one leg stalled (r2y: value 24.5 ms against e2e 14.1 ms for the same step).  The nvidia-smi loop stays as
    the fall-back when NVML cannot be loaded, and then start() waits for its first row before anything is timed."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, period=0.1):
        super().__init__(daemon=True)
        self.rows, self.proc, self.index, self.period = [], None, index, period
        self._halt = threading.Event()
        self._first = threading.Event()
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # torch's device index counts CUDA_VISIBLE_DEVICES entries; NVML counts physical devices: go through the PCI id
            bus = torch.cuda.get_device_properties(index)
            bdf = "%08x:%02x:%02x.0" % (bus.pci_domain_id, bus.pci_bus_id, bus.pci_device_id)
            self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(bdf.encode())
            self.nvml = pynvml
            self.bits = [pynvml.nvmlClocksEventReasonHwSlowdown, pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                         pynvml.nvmlClocksEventReasonSwThermalSlowdown, pynvml.nvmlClocksEventReasonSwPowerCap]
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.nvml = None

    def _run_nvml(self):
        n = self.nvml
        reasons_fn = n.nvmlDeviceGetCurrentClocksEventReasons
        while not self._halt.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(reasons_fn(self.handle))
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                self.rows.append([str(sm), str(self.max_mhz), "%.1f" % pw] + ["Active" if mask & b else "Not Active" for b in self.bits])
            except Exception:  # noqa: BLE001
                pass
            self._first.set()
            self._halt.wait(self.period)

    def run(self):
        if self.nvml is not None:
            return self._run_nvml()
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                self._first.set()
        except Exception:  # noqa: BLE001
            pass
        self._first.set()

    def start_and_wait(self):
        """Starts sampling and returns once the first sample is in (at most 5 s), so no client start-up overlaps a timed step."""
        self.start()
        self._first.wait(5.0)

    def stop(self):
        self._halt.set()
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        reasons = [n for i, n in enumerate(self.NAMES)
                   if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        busy = sorted(sm)[len(sm) // 4:] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "sm_mhz_min": min(sm) if sm else None, "power_w_max": max(pw) if pw else None,
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return (d.get("bf16_tflops_sustained", 1371.4), d.get("hbm_gbs", 6553.3), "measured (MEASURED_PEAKS.json, sustained)",
                d.get("bf16_tflops", 1650.0))
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)", 1650.0


def pick_peak(clk, sustained, burst, src):
    """Roofline denominator for a kernel timed inside the step.  MEASURED_PEAKS.json holds two tensor figures: `burst`
    (best single matmul, SM clock at its maximum) and `sustained` (matmuls back to back for seconds, by which time the power
    cap has pulled the SM clock to ~1.25 GHz).  The timed region here is K steps of ~14 ms: when the sampled SM clock stayed
    within 10 % of its maximum the GPU was in the burst regime and the burst figure is the honest (larger) denominator;
    otherwise the sustained one.  Both fractions are reported either way."""
    if clk and clk.get("sm_mhz") and clk.get("sm_max_mhz") and clk["sm_mhz"] >= 0.9 * clk["sm_max_mhz"]:
        return burst, src.replace("sustained", "burst: SM clock %.0f of %.0f MHz during the timed region" % (clk["sm_mhz"], clk["sm_max_mhz"]))
    return sustained, src


_ORIG_AFFINITY = None


def bind_to_gpu_numa(local):
    """Pin this rank to the CPUs that are local to its GPU (sysfs local_cpulist of the GPU's PCI function) BEFORE any pinned
    host buffer is allocated: first-touch then places the staging memory on the GPU's own NUMA node, so 8 ranks pulling
    267 MB of host noise per step do not all cross the same socket interconnect (round-1 finding: e2e efficiency 0.84 at
    N=8 with unplaced buffers).  Best effort; returns a short description for the JSON line."""
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            global _ORIG_AFFINITY
            _ORIG_AFFINITY = os.sched_getaffinity(0)
            os.sched_setaffinity(0, cpus)
            return "%s -> %d local cpus" % (bdf, len(cpus))
    except Exception as e:  # noqa: BLE001
        return "unbound (%s)" % type(e).__name__
    return "unbound"


def port_step(sd, O, state, goal, n, seed):
    """One CEM iteration of the CPU restatement (oracle port) on n candidates; returns seconds."""
    r = np.random.default_rng(seed)
    end = r.integers(2, 200, size=n)
    t0 = time.perf_counter()
    samples = r.normal(0, 0.3, size=(n, 255, 256))
    with torch.no_grad():
        ro = O.simulator_rollout(sd, state, goal, samples, end)
    imgs = [p[:, :3072].reshape(-1, 3, 32, 32) for p in ro["predictions"]]
    cost = O.l2_image_cost(imgs, goal, True, 1.0)
    el = O.elites(cost, n, max(ELITE_FRAC, 1.0 / n))
    O.refit(samples, el)
    return time.perf_counter() - t0


def cpu_reference_arm(sd, state, goal, n, steps, warmup):
    """Times the reference's CPU path on the host cores: the UNMODIFIED reference staged under baseline/_ref when it is
    there (kind "reference"), else the oracle port.  Returns (rollouts/s, seconds per step, kind, cores)."""
    from oracle import ref_arm
    if _ORIG_AFFINITY is not None:
        os.sched_setaffinity(0, _ORIG_AFFINITY)       # the CPU arm may use every host core again
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    if ref_arm.available():
        R = ref_arm.ReferenceCEM(sd, cores)
        fn, kind = (lambda s: R.step(state, goal, n, max(ELITE_FRAC, 1.0 / n), s)), "reference"
    else:
        from oracle import gcp_oracle as O
        fn, kind = (lambda s: port_step(sd, O, state, goal, n, s)), "port"
    for w in range(warmup):
        fn(100 + w)
    t = sum(fn(200 + s) for s in range(steps))
    return n * steps / t, t / steps, kind, cores


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the step on the host cores.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from video_gcp_b200 import hparams
    from video_gcp_b200.synthetic import synthetic_state_dict
    hp = hparams.build_hparams(hparams.gcp_tree_25room_config(batch_size=1))
    sd = synthetic_state_dict(hp, 1)
    r = np.random.default_rng(0)
    state = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    goal = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    n = args.ref_sample
    v, sec, kind, cores = cpu_reference_arm(sd, state, goal, n, args.steps, min(args.warmup, 1))
    sample = "%d of %d candidates per step (same per-candidate work)" % (n, args.candidates)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload(args.candidates), "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


LAST_STEP_MS = []


def timed(fn, steps, warmup, world, dev, dist):
    import gc
    for _ in range(warmup):
        fn()
    gc.collect()
    gc.disable()              # a generation-2 collection inside the K timed steps would be charged to the step
    try:
        return _timed(fn, steps, world, dev, dist)
    finally:
        gc.enable()


def _timed(fn, steps, world, dev, dist):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    LAST_STEP_MS[:] = [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(steps)]    # this rank's steps (diagnostic)
    ms = torch.tensor([ev[0].elapsed_time(ev[steps])], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="tree", choices=["tree", "seq"])
    ap.add_argument("--candidates", type=int, default=1024, help="candidates per GPU (weak scaling)")
    ap.add_argument("--candidates-total", type=int, default=0,
                    help="total candidates, sharded over the ranks (strong scaling; BASELINE config 5)")
    ap.add_argument("--ref-sample", type=int, default=8, help="candidates per CPU-reference step")
    ap.add_argument("--planner-iters", type=int, default=2, help="n_iters of the e2e_planner call")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only value / e2e (skip e2e_host_noise, e2e_planner, value_pruned)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "seq":
        return main_seq(args)

    import torch.distributed as dist
    from video_gcp_b200 import hparams
    from video_gcp_b200.model import TreeModel
    from video_gcp_b200.planning import GCPImageSimulator, ImageCEMPlanner, L2ImageCost, SimpleTreeCEMSampler
    from video_gcp_b200.synthetic import synthetic_state_dict

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    strong = args.candidates_total > 0
    if strong:
        assert args.candidates_total % world == 0, "--candidates-total must divide over the ranks"
        B = args.candidates_total // world
    else:
        B = args.candidates
    N = B * world
    k = max(int(N * ELITE_FRAC), 1)
    chunk = min(B, 8192)                      # rollout chunk = engine capacity (3.5 GB of workspace per 1024 candidates)
    assert B % chunk == 0

    hp_cfg = hparams.gcp_tree_25room_config(batch_size=1)
    model = TreeModel(hp_cfg, None, max_candidates=chunk)
    model.load_state_dict(synthetic_state_dict(model._hp, 1), strict=True)
    model.device = dev
    model.eval()
    eng = model.engine
    sim = GCPImageSimulator(model, append_latent=False)

    def make_planner(n_iters=1, seed=7, pruned=False, sort_lengths=None):
        return ImageCEMPlanner(dict(batch_size=N, n_iters=n_iters, elite_frac=ELITE_FRAC, cost_fcn=L2ImageCost, dense_cost=True,
                                    final_step_cost_weight=1.0, sampler=partial(SimpleTreeCEMSampler, n_level_hierarchy=8),
                                    max_seq_len=200, action_dim=256, initial_std=0.3, max_rollout_bs=chunk, seed=seed,
                                    prune_before_decode=pruned, sort_lengths=sort_lengths), sim)

    # `value` / `e2e`: the canonical workload, all 255 nodes of every candidate decoded (what the reference computes)
    planner = make_planner()
    planner._sampler.init()

    r = np.random.default_rng(0)
    state = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    goal = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    state_t = torch.as_tensor(state).pin_memory()
    goal_t = torch.as_tensor(goal).pin_memory()

    def step_value():
        return planner.cem_iteration(state_t, goal_t)

    trace = os.environ.get("BENCH_TRACE") == "1"      # per-call host wall times on stderr (diagnostic)

    def step_e2e():
        t0 = time.perf_counter()
        cost, idx, val, _ = planner.cem_iteration(state_t, goal_t)
        t1 = time.perf_counter()
        out = cost.cpu(), idx.cpu()
        if trace:
            print("[e2e] enqueue %.2f ms, read-back wait %.2f ms" % ((t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3), file=sys.stderr)
        return out

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start_and_wait()
    l0 = eng.launch_count()
    ms = timed(step_value, args.steps, args.warmup, world, dev, dist)
    step_ms = list(LAST_STEP_MS)
    launches = (eng.launch_count() - l0) // (args.steps + args.warmup)
    ms_e2e = timed(step_e2e, args.steps, max(args.warmup, 3), world, dev, dist)
    step_ms_e2e = list(LAST_STEP_MS)
    clk = clocks.stop() if rank == 0 else None

    extras = {}
    if not args.no_extras:
        # ---- the reference simulator's contract: candidates in pinned HOST memory (this rank's slice), copied in the call
        n_host = min(B, chunk)
        z_host = planner._sampler.sample_device(n_host, first_id=rank * B).cpu().pin_memory()
        cost_fcn = planner._cost_fcn

        def step_host_noise():
            ro = sim.rollout_device(state_t, goal_t, z_host, 200)
            c = cost_fcn.device_cost(ro)
            if world > 1:
                full = torch.empty(n_host * world, device=dev, dtype=torch.float32)
                dist.all_gather_into_tensor(full, c)
                c = full
            idx, val = eng.topk(c, max(int(n_host * world * ELITE_FRAC), 1))
            if world == 1:
                eng.refit(ro.z, idx)
            return c.cpu(), idx.cpu()

        ms_hn = timed(step_host_noise, args.steps, 3, world, dev, dist)
        extras["e2e_host_noise"] = {
            "value": n_host * world * args.steps / (ms_hn * 1e-3), "unit": UNIT, "ms_per_step": ms_hn / args.steps,
            "h2d_bytes_per_step": int(z_host.numel() * 4 + 2 * 3072 * 4), "d2h_bytes_per_step": int(n_host * world * 4 + k * 4),
            "candidates_per_gpu": n_host, "numa": numa,
            "note": "simulator.rollout(state, goal, samples) with this rank's samples in pinned host memory + cost + "
                    "all-gather + top-k" + (" + refit" if world == 1 else " (no refit: injected samples are not regenerable)")}
        del z_host

        # ---- planner mode: decode only the nodes the planner reads, L2 cost folded into the decoder tail (no image writes)
        plp = make_planner(pruned=True)
        plp._sampler.init()
        ms_p = timed(lambda: plp.cem_iteration(state_t, goal_t), args.steps, args.warmup, world, dev, dist)
        pruned_phase = None
        if B == chunk:
            eng.profile_enable(True)
            for _ in range(2):
                plp.cem_iteration(state_t, goal_t)
            torch.cuda.synchronize()
            pp = eng.profile_read()
            eng.profile_enable(False)
            pruned_phase = {ph: round(pp[ph] / 2, 3) for ph in eng.PHASES}
        # same candidates, same rollout seeds -> the two modes must agree bit for bit (checked outside the timed region)
        # (the full-decode side also takes the sampled lengths in descending order, so both see the same lengths)
        pa, pb = make_planner(seed=31, sort_lengths=True), make_planner(seed=31, pruned=True)
        pa._sampler.init(), pb._sampler.init()
        model.seed = 1000
        ca, ia, va, _ = pa.cem_iteration(state_t, goal_t)
        kept = float((sim._model.engine._bufs[("end_ind", (chunk,), torch.int64)].double() + 1).mean()) if B == chunk else None
        model.seed = 1000
        cb, ib, vb, _ = pb.cem_iteration(state_t, goal_t)
        same = bool(torch.equal(ia, ib) and torch.equal(pa._sampler._mean_d, pb._sampler._mean_d)
                    and torch.equal(pa._sampler._std_d, pb._sampler._std_d))
        same = same and bool(torch.equal(ca, cb))
        # decoder = 69 %, tree recursion = 31 % of the canonical work, both per node; the tree is pruned
        # at (node, 128-candidate tile) granularity, so the executed figure is a lower bound
        NODE_FLOP = FLOP_PER_ROLLOUT / 255
        executed = None if kept is None else NODE_FLOP * kept
        extras["value_pruned"] = {
            "value": N * args.steps / (ms_p * 1e-3), "unit": UNIT, "ms_per_step": ms_p / args.steps,
            "speedup_vs_value": ms / ms_p, "phase_ms_per_step": pruned_phase, "kept_nodes_mean": kept, "executed_flop_per_rollout": executed,
            "same_costs_elites_refit_as_full_decode": same,
            "note": "ImageCEMPlanner(prune_before_decode=True).cem_iteration: only the end_ind+1 nodes balanced pruning keeps "
                    "are decoded AND computed by the tree recursion (sampled lengths handed out in descending order so that "
                    "128-candidate tiles share a length class), the L2 cost is reduced in the decoder-tail epilogue, no image "
                    "is written, no existence / action / state heads; `value` (all 255 nodes, canonical FLOPs) stays the headline"}

        # ---- a whole planner call: n_iters iterations + final rollout of the elites + the plan on the host
        if B <= chunk:
            pl2 = make_planner(n_iters=args.planner_iters, seed=11, pruned=True)

            def plan():
                return pl2(state_t, goal_t)

            keep = [plan(), plan()]      # warm-up; results held as a policy holds the current plan
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            import gc
            gc.collect()
            gc.disable()          # as in timed(): a generation-2 collection of this process's heap is 10-40 ms, one plan is 15
            t0 = time.perf_counter()
            n_plans, plan_ms = 3, []
            marks = []
            if trace:       # BENCH_TRACE=1: synchronised host time of every stage of the call (diagnostic, slows the call)
                def wrap(obj, name, label):
                    fn = getattr(obj, name)

                    def inner(*a, **k):
                        torch.cuda.synchronize()
                        t_in = time.perf_counter()
                        res = fn(*a, **k)
                        torch.cuda.synchronize()
                        marks.append("%s %.2f" % (label, (time.perf_counter() - t_in) * 1e3))
                        return res
                    setattr(obj, name, inner)
                from video_gcp_b200.planning import cem_simulator as _cs
                wrap(pl2, "cem_iteration", "iter"), wrap(pl2, "_elite_samples", "elite_z"), wrap(pl2, "_rollout_host", "final")
                wrap(sim, "rollout_device", "rollout_device"), wrap(_cs.DeviceRollouts, "to_host", "to_host")
            for _ in range(n_plans):
                a0 = torch.cuda.memory_stats().get("num_device_alloc", 0)
                t1 = time.perf_counter()
                frames, actions, latents, score = plan()      # returns host arrays: the call itself synchronises
                plan_ms.append(round((time.perf_counter() - t1) * 1e3, 3))
                if trace:
                    print("[plan] %.2f ms | %s | cudaMalloc +%d" % (plan_ms[-1], ", ".join(marks),
                                                                   torch.cuda.memory_stats().get("num_device_alloc", 0) - a0), file=sys.stderr)
                    marks.clear()
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device=dev)
            gc.enable()
            del keep
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dt = float(dt.item()) / n_plans
            extras["e2e_planner"] = {
                "ms_per_plan": dt * 1e3, "plan_ms": plan_ms, "n_iters": args.planner_iters, "rollouts_per_plan": args.planner_iters * N + k,
                "value": (args.planner_iters * N + k) / dt, "unit": UNIT,
                "note": "ImageCEMPlanner.__call__(state, goal): %d CEM iterations over %d candidates, final rollout of the %d "
                        "elites, the plan (best rollout: frames, actions, latents) and the elite scores copied to the host (host wall clock, synchronised)" % (args.planner_iters, N, k)}

    # ---- N > 1: every rank must hold the same elites / distribution, and rank 0 must reproduce any rank's costs
    rank_consistent = None
    if world > 1:
        seed0 = model.seed
        cost, idx, val, packed = planner.cem_iteration(state_t, goal_t)
        chk = torch.stack([idx.double().sum(), (idx.double() * torch.arange(1, k + 1, device=dev)).sum(),
                           planner._sampler._mean_d.double().sum(), planner._sampler._std_d.double().sum(), cost.double().sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(lo, hi))
        reroll = True
        if rank == 0:
            # re-roll the LAST rank's candidate ids here: same global ids (regenerated from the shared noise stream), same
            # rollout seeds as that rank used -> its slice of the gathered cost vector, bit for bit
            other = world - 1
            ids = torch.arange(other * B, (other + 1) * B, device=dev, dtype=torch.int32)
            model.seed = seed0
            c2, _ = planner._rollout_costs(state_t, goal_t, planner._sampler.regenerate(ids))
            reroll = bool(torch.equal(c2, cost[other * B:(other + 1) * B]))
        flag = torch.tensor([1.0 if (same and reroll) else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        rank_consistent = bool(flag.item() == 1.0)

    # ---- roofline of the dominant kernel (decoder tail conv), timed live with CUDA events on its stream
    eng.profile_enable(True)
    for _ in range(2):
        step_value()
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
    peak_used, peak_src_used = pick_peak(clk, peak_tf, peak_burst, peak_src)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "step_ms": step_ms, "step_ms_e2e": step_ms_e2e, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload(B), "candidates_total": N, "candidates_per_gpu": B, "rollout_chunk": chunk, "elites": k,
                   "step": "ImageCEMPlanner.cem_iteration (sample -> rollout -> L2 cost -> all-gather -> top-k -> refit)",
                   "l2": "inputs larger than L2 (noise z = %.0f MB drawn + read per step; images written %.1f GB)"
                         % (B * 255 * 256 * 4 / 1e6, B * 255 * 3072 * 4 / 1e9),
                   "parallelism": "candidates sharded over %d rank(s), cost all-gather" % world},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(2 * 3072 * 4),
                "d2h_bytes_per_step": int(N * 4 + k * 4), "ms_per_step": ms_e2e / args.steps,
                "note": "planner call from pinned host start / goal images; the candidates are drawn inside the call (as the "
                        "reference planner does with np.random); costs + elite ids copied back every step"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "dec_tail3_kernel", "achieved": tail_tf, "peak": peak_used,
                     "unit": "TFLOP/s", "frac": tail_tf / peak_used,
                     "frac_of_burst": tail_tf / peak_burst, "frac_of_sustained": tail_tf / peak_tf,
                     "executed": {"achieved": tail_tf * TAIL_EXECUTED_FLOP_PER_IMAGE / TAIL_FLOP_PER_IMAGE,
                                  "frac": tail_tf * TAIL_EXECUTED_FLOP_PER_IMAGE / TAIL_FLOP_PER_IMAGE / peak_used,
                                  "frac_of_burst": tail_tf * TAIL_EXECUTED_FLOP_PER_IMAGE / TAIL_FLOP_PER_IMAGE / peak_burst,
                                  "note": "tcgen05 FLOPs actually issued (29.36 MFLOP/image vs 32.51 canonical: shared "
                                          "skip half + unused mixture-scale channels are not computed)"},
                     "traffic": TAIL_TRAFFIC_BYTES_B1024 if chunk == 1024 else None,
                     "traffic_note": "bytes per launch, ncu dram read+write (profiles/r2o_dec_tail3_ncu_full.txt)",
                     "peak_source": peak_src_used,
                     "ms_per_launch": tail_ms, "images_per_launch": tail_imgs},
        "roofline_step": {"bound": "tensor", "achieved": value / world * FLOP_PER_ROLLOUT / 1e12, "peak": peak_used,
                          "unit": "TFLOP/s per GPU (canonical 16.75 GFLOP/rollout)",
                          "frac": value / world * FLOP_PER_ROLLOUT / 1e12 / peak_used,
                          "frac_of_burst": value / world * FLOP_PER_ROLLOUT / 1e12 / peak_burst,
                          "frac_of_sustained": value / world * FLOP_PER_ROLLOUT / 1e12 / peak_tf},
        "phase_ms_per_step": phase_ms,
    }
    out.update(extras)
    if rank_consistent is not None:
        out["rank_consistent"] = rank_consistent
    if world == 1 and not args.no_cpu_baseline:
        sd = synthetic_state_dict(model._hp, 1)
        n = args.ref_sample
        v, sec, kind, cores = cpu_reference_arm(sd, state, goal, n, 3, 1)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                               "sample": "3 steps of %d candidates (of %d), same per-candidate work" % (n, B)}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main_seq(args):
    """BASELINE config 3: the sequential GCP rollout (199-step recurrence) + dense L2 cost + elite top-k; one line, same
    schema.  Single GPU or replicas sharded by candidate like the tree path."""
    import torch.distributed as dist
    from video_gcp_b200 import hparams
    from video_gcp_b200.model import SequentialModel
    from video_gcp_b200.planning import GCPImageSimulator, L2ImageCost
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
    model = SequentialModel(hparams.gcp_sequential_25room_config(batch_size=1), None, max_candidates=B)
    model.load_state_dict(synthetic_state_dict(model._hp, 1), strict=True)
    model.device = dev
    model.eval()
    eng = model.engine
    sim = GCPImageSimulator(model, append_latent=False)
    cost_fcn = L2ImageCost(True, 1.0)
    r = np.random.default_rng(0)
    state_t = torch.as_tensor(r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)).pin_memory()
    goal_t = torch.as_tensor(r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)).pin_memory()
    z_dev = torch.randn(B, 199, 256, device=dev) * 0.3
    z_host = z_dev.cpu().pin_memory()

    def tail(ro):
        c = cost_fcn.device_cost(ro)
        if world > 1:
            full = torch.empty(N, device=dev, dtype=torch.float32)
            dist.all_gather_into_tensor(full, c)
            c = full
        idx, val = eng.topk(c, k)
        return c, idx

    def step_value():
        return tail(sim.rollout_device(state_t, goal_t, z_dev, 200))

    def step_e2e():
        c, idx = tail(sim.rollout_device(state_t, goal_t, z_host, 200))
        return c.cpu(), idx.cpu()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start_and_wait()
    l0 = eng.launch_count()
    ms = timed(step_value, args.steps, args.warmup, world, dev, dist)
    step_ms = list(LAST_STEP_MS)
    launches = (eng.launch_count() - l0) // (args.steps + args.warmup)
    ms_e2e = timed(step_e2e, args.steps, max(args.warmup, 3), world, dev, dist)
    step_ms_e2e = list(LAST_STEP_MS)
    clk = clocks.stop() if rank == 0 else None
    eng.profile_enable(True)
    for _ in range(2):
        step_value()
    torch.cuda.synchronize()
    prof = eng.profile_read()
    eng.profile_enable(False)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak_tf, peak_hbm, peak_src, peak_burst = measured_peaks()
    peak_used, peak_src_used = pick_peak(clk, peak_tf, peak_burst, peak_src)
    value = N * args.steps / (ms * 1e-3)
    rec_ms = prof["tree_recursion"] / 2
    print(json.dumps({
        "metric": "sequential-GCP CEM rollouts/sec", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "step_ms": step_ms, "step_ms_e2e": step_ms_e2e, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload(B, "seq"), "candidates_total": N, "elites": k,
                   "l2": "inputs larger than L2 (images written %.1f GB per step)" % (B * 200 * 3072 * 4 / 1e9)},
        "e2e": {"value": N * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(z_host.numel() * 4 + 2 * 3072 * 4),
                "d2h_bytes_per_step": int(N * 4 + k * 4), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "199-step recurrence (prior MLP + reparametrisation + 3 LSTM cells + output)",
                     "achieved": value / world * SEQ_FLOP_PER_ROLLOUT / 1e12, "peak": peak_used, "unit": "TFLOP/s",
                     "frac": value / world * SEQ_FLOP_PER_ROLLOUT / 1e12 / peak_used,
                     "frac_of_burst": value / world * SEQ_FLOP_PER_ROLLOUT / 1e12 / peak_burst,
                     "frac_of_sustained": value / world * SEQ_FLOP_PER_ROLLOUT / 1e12 / peak_tf, "traffic": None,
                     "peak_source": peak_src_used,
                     "note": "whole-step canonical FLOPs (2 x 9.76 GMAC per rollout) over the step time"},
        "phase_ms_per_step": {p: round(prof[p] / 2, 3) for p in eng.PHASES}, "recurrence_ms": rec_ms,
    }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
