"""Prints the numerical gap between the CUDA rollout and the CPU oracle, stage by stage, for both the
SIMT verification kernels and the tcgen05 product kernels.  Used to SET the tolerances written in
tests/test_gpu_parity.py; run on the B200 box:  python tests/gpu_report.py [B]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gcp_oracle as O  # noqa: E402
from video_gcp_b200 import hparams  # noqa: E402
from video_gcp_b200.engine import Engine  # noqa: E402
from video_gcp_b200.synthetic import synthetic_rollout_inputs, synthetic_state_dict  # noqa: E402


def stats(name, got, ref):
    got = got.detach().double().cpu() if isinstance(got, torch.Tensor) else torch.as_tensor(got).double()
    ref = ref.detach().double().cpu() if isinstance(ref, torch.Tensor) else torch.as_tensor(ref).double()
    d = (got - ref).abs()
    rel = d.max() / ref.abs().max().clamp_min(1e-12)
    rms = (d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt().clamp_min(1e-12))
    print("  %-26s max|d| %.3e   max|ref| %.3e   max-rel %.3e   rms-rel %.3e   nan %d"
          % (name, d.max(), ref.abs().max(), rel, rms, int(torch.isnan(got).sum())))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    hp = hparams.build_hparams(hparams.gcp_tree_25room_config(batch_size=1, attach_cost_mdl=True))
    sd = synthetic_state_dict(hp, 1)
    inp = synthetic_rollout_inputs(B, seed=3, shared_images=False)
    torch.set_num_threads(os.cpu_count())
    t0 = time.time()
    with torch.no_grad():
        ref = O.rollout(sd, inp["I_0"], inp["I_g"], inp["z"], inp["end_ind"].numpy())
    print("oracle (CPU, %d threads) B=%d: %.2f s" % (torch.get_num_threads(), B, time.time() - t0))
    dev = torch.device("cuda:0")
    outs = {}
    for mode in ("ref", "tc"):
        print("== %s kernels ==" % mode)
        from verify_lib import verify_engine
        eng = verify_engine(dev, max_candidates=128, attach_cost_mdl=True) if mode == "ref" else \
            Engine(dev, max_candidates=128, attach_cost_mdl=True)
        eng.load_weights(sd)
        out = eng.rollout(inp["I_0"].to(dev), inp["I_g"].to(dev), inp["z"].to(dev), end_ind=inp["end_ind"].to(dev),
                          want_prior=True)
        torch.cuda.synchronize()
        outs[mode] = out
        stats("e_0", out["e_0"], ref["e0"])
        stats("e_g", out["e_g"], ref["eg"])
        stats("seq_len_logits", out["seq_len_logits"], ref["seq_len_logits"])
        stats("mu_df", out["mu_df"], ref["tree"]["mu"])
        stats("log_sigma_df", out["log_sigma_df"], ref["tree"]["log_sigma"])
        stats("e_df (latents)", out["e_df"], ref["tree"]["e"])
        for lvl in range(8):
            idx = [O.df_index(lvl, j) for j in range(2 ** lvl)]
            stats("  e_df level %d" % lvl, out["e_df"][:, idx], ref["tree"]["e"][:, idx])
        stats("images_df", out["images_df"], ref["images_df"])
        stats("existence", out["existence"], ref["existence"])
        lmax = ref["model_enc_seq"].shape[1]
        stats("model_enc_seq", out["model_enc_seq"][:, :lmax], ref["model_enc_seq"])
        stats("actions", out["actions"][:, :lmax - 1], ref["actions"])
        stats("regressed_state", out["regressed_state"][:, :lmax], ref["regressed_state"])
        # costs
        goal = inp["I_g"][0]
        imgs = [p.numpy() for p in ref["pruned_images"]]
        l2_ref = O.l2_image_cost(imgs, ((goal.permute(1, 2, 0)[None] + 1) / 2).numpy(), True, 1.0)
        l2 = eng.cost_l2(out["images_df"], out["end_ind"], goal.to(dev), True, 1.0)
        stats("cost_l2 (own images)", l2, l2_ref)
        l2o = eng.cost_l2(ref["images_df"].to(dev).contiguous(), out["end_ind"], goal.to(dev), True, 1.0)
        stats("cost_l2 (oracle images)", l2o, l2_ref)
        lat = [p.numpy() for p in ref["pruned_latents"]]
        with torch.no_grad():
            lc_ref = O.image_wrapped_learned_cost(sd, lat)
        lc = eng.cost_learned(ref["tree"]["e"].to(dev).contiguous(), out["end_ind"], ref["pruned_latents"][-1].to(dev))
        stats("cost_learned (oracle lat)", lc, lc_ref)
        idx, val = eng.topk(l2, max(B // 2, 1))
        print("  topk idx", idx.tolist(), "oracle", O.elites(l2_ref, B, 0.5).tolist())
        z = inp["z"].to(dev)
        mean, std = eng.refit(z, idx)
        m_ref, s_ref = O.refit(inp["z"].double().numpy(), idx.cpu().numpy())
        stats("refit mean", mean, m_ref)
        stats("refit std", std, s_ref)
        print("  launches:", eng.launch_count())
        eng.close()
    print("== tc vs ref ==")
    for k in ("e_df", "images_df", "mu_df", "actions", "regressed_state", "existence", "seq_len_logits"):
        stats(k, outs["tc"][k], outs["ref"][k])

    # ---- first throughput number
    for Bt in (256, 1024):
        try:
            eng = Engine(dev, max_candidates=Bt)
            eng.load_weights(sd)
            big = synthetic_rollout_inputs(Bt, seed=5, shared_images=True)
            I0, Ig, z, ei = big["I_0"][:1].to(dev), big["I_g"][:1].to(dev), big["z"].to(dev), big["end_ind"].to(dev)
            for _ in range(2):
                out = eng.rollout(I0, Ig, z, end_ind=ei, images_shared=True, want_existence=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 3
            for _ in range(n):
                out = eng.rollout(I0, Ig, z, end_ind=ei, images_shared=True, want_existence=True)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            print("B=%d rollout: %.2f ms  -> %.0f rollouts/s, %.1f TFLOP/s canonical, nan=%d"
                  % (Bt, ms, Bt / ms * 1e3, Bt / ms * 1e3 * O.FLOP_PER_ROLLOUT / 1e12, int(torch.isnan(out["images_df"]).sum())))
            eng.close()
            del out
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            print("B=%d timing failed: %r" % (Bt, e))


if __name__ == "__main__":
    main()
