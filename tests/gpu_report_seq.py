"""Numerical gap and timing of the sequential GCP rollout against the CPU oracle (sets the tolerances written in
tests/test_gpu_parity_seq.py).  Run on the B200 box:  python tests/gpu_report_seq.py [B_parity] [B_timing]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gcp_oracle as O  # noqa: E402
from video_gcp_b200 import hparams  # noqa: E402
from video_gcp_b200.engine import Engine  # noqa: E402
from video_gcp_b200.synthetic import synthetic_seq_inputs, synthetic_state_dict  # noqa: E402


def stats(name, got, ref):
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    d = (got - ref).abs()
    rel = d.max() / ref.abs().max().clamp_min(1e-12)
    rms = (d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt().clamp_min(1e-12))
    print("  %-26s max|d| %.3e   max|ref| %.3e   max-rel %.3e   rms-rel %.3e   nan %d"
          % (name, d.max(), ref.abs().max(), rel, rms, int(torch.isnan(got).sum())))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    Bt = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    hp = hparams.build_hparams(hparams.gcp_sequential_25room_config(batch_size=1))
    sd = synthetic_state_dict(hp, 2)
    inp = synthetic_seq_inputs(B, seed=3, shared_images=False)
    torch.set_num_threads(os.cpu_count())
    t0 = time.time()
    with torch.no_grad():
        ref = O.seq_rollout(sd, inp["I_0"], inp["I_g"], inp["z"], [199] * B)
    print("oracle (CPU, %d threads) B=%d: %.2f s" % (torch.get_num_threads(), B, time.time() - t0))
    dev = torch.device("cuda:0")
    for use_ref in (True, False):
        from verify_lib import verify_engine
        eng = verify_engine(dev, max_candidates=max(B, 128), model="sequential") if use_ref else \
            Engine(dev, max_candidates=max(B, 128), model="sequential")
        eng.load_weights(sd)
        out = eng.seq_rollout(inp["I_0"].to(dev), inp["I_g"].to(dev), inp["z"].to(dev), end_ind=inp["end_ind"].to(dev),
                              want_prior=True)
        torch.cuda.synchronize()
        print("SIMT verification kernels" if use_ref else "tcgen05 product kernels")
        stats("e_0", out["e_0"], ref["e0"])
        stats("encodings", out["encodings"], ref["encodings"])
        for t0_ in (0, 10, 50, 100, 150, 198):
            stats("  encodings[t=%d]" % t0_, out["encodings"][:, t0_], ref["encodings"][:, t0_])
        stats("mu", out["mu"], ref["mu"])
        stats("log_sigma", out["log_sigma"], ref["log_sigma"])
        stats("images", out["images"], ref["images"])
        stats("actions", out["actions"][:, :199], ref["actions"])
        stats("regressed_state", out["regressed_state"], ref["regressed_state"])
        eng.close()
    # timing at the benchmark size
    eng = Engine(dev, max_candidates=Bt, model="sequential")
    eng.load_weights(sd)
    inp = synthetic_seq_inputs(Bt, seed=5, shared_images=True)
    I0, Ig, z = inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["z"].to(dev)
    for _ in range(2):
        eng.seq_rollout(I0, Ig, z, images_shared=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.launch_count()
    t0 = time.perf_counter()
    e0.record()
    n = 4
    for i in range(n):
        eng.seq_rollout(I0, Ig, z, seed=i, images_shared=True)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("B=%d sequential rollout: %.2f ms (%.0f rollouts/s), cpu enqueue %.2f ms/iter, launches/iter %d"
          % (Bt, ms, Bt / ms * 1e3, (t1 - t0) * 1e3 / n, (eng.launch_count() - l0) // n))
    eng.profile_enable(True)
    for i in range(2):
        eng.seq_rollout(I0, Ig, z, images_shared=True)
    torch.cuda.synchronize()
    print({k: (round(v / 2, 3) if isinstance(v, float) else v) for k, v in eng.profile_read().items()})
    # 9.76 GMAC per rollout without the inference LSTM
    print("  %.1f TFLOP/s on the canonical 2 x 9.76 GMAC per rollout" % (Bt / ms * 1e3 * 2 * 9.76e9 / 1e12))


if __name__ == "__main__":
    main()
