"""CPU oracle of the training-phase forward + loss of the 25-room GCP-tree (BASELINE config 1).

TEST INFRASTRUCTURE ONLY (same rules as oracle/gcp_oracle.py: imported by tests/, smoke() and bench.py's CPU legs as
the checker / reported baseline, never by the product package).

Parity status: PINNED against the reference itself.  `oracle/make_golden_train.py` runs the UNMODIFIED reference
`TreeModel` (prediction config, `.train()` mode) through `model(inputs)`, `model.loss`, `model.get_total_loss` with the
posterior noise, the auxiliary heads' sampled indices and the cost target injected / recorded, and stores losses and
intermediates in tests/golden/train_forward_B2.npz and train_losses_B16.npz; tests/test_oracle_train.py checks every
function below against them.

A plain torch-fp32 functional restatement; paths cite /root/reference.
"""
import contextlib

import numpy as np
import torch
import torch.nn.functional as F

from .gcp_oracle import (BN_EPS, DEPTH, GN_EPS, LRELU, N_NODES, _up_pad_conv, balanced_keep_mask, df_index, interleave,
                         lstm_cell, mlp)

LOSS_NAMES = ("len_pred", "action_reconst", "cost_estimation", "state_regression", "dense_img_rec", "kl",
              "existence_predictor", "entropy", "total")


@contextlib.contextmanager
def bf16_operands():
    """Error-envelope aid for the GPU parity tests: inside this context every F.linear of the oracle rounds its two
    operands to bf16 (fp32 accumulation, fp32 bias), which is the arithmetic of the device's tensor-core GEMMs.  Running
    the tree with and without it shows how much of a device-vs-fp32 difference is operand rounding amplified by the
    network (random-init posteriors reach sigma = 17.8, and z = mu + sigma * eps multiplies log-sigma errors by that)."""
    lin = F.linear
    r = lambda t: t.bfloat16().float()
    F.linear = lambda x, w, b=None: lin(r(x), r(w), b)
    try:
        yield
    finally:
        F.linear = lin


def _bn_train(sd, name, x):
    """nn.BatchNorm2d in training mode: biased batch statistics over (N, H, W) (blox/torch/layers.py:62-77)."""
    return F.batch_norm(x, None, None, sd[name + ".weight"], sd[name + ".bias"], True, 0.0, BN_EPS)


def encoder_train(sd, img):
    """ConvEncoder (blox/torch/encoder_decoder.py:31-53) with batch-statistic BN over the images of ONE call.
    img [N,3,32,32] -> e [N,128], skips s0 [N,16,16,16], s2 [N,64,4,4]."""
    p = "encoder.net.net."
    x = F.leaky_relu(F.conv2d(img, sd[p + "input.conv.weight"], sd[p + "input.conv.bias"], 2, 1), LRELU)
    s0 = x
    x = F.leaky_relu(_bn_train(sd, p + "pyramid-0.norm", F.conv2d(x, sd[p + "pyramid-0.conv.weight"], None, 2, 1)), LRELU)
    x = F.leaky_relu(_bn_train(sd, p + "pyramid-1.norm", F.conv2d(x, sd[p + "pyramid-1.conv.weight"], None, 2, 1)), LRELU)
    s2 = x
    e = F.conv2d(x, sd[p + "head.weight"], sd[p + "head.bias"])
    return e[:, :, 0, 0], (s0, s2)


def inf_encoder(sd, enc_seq):
    """ConvSeqEncodingModule (blox/torch/subnetworks.py:120-147): append the frame index as a 129th channel, then
    Conv1d(129->128,k3,p1)+LReLU -> Conv1d(128->128, no bias)+GroupNorm(8)+LReLU -> Conv1d(128->128)+bias over time.
    enc_seq [B,T,128] -> [B,T,128]."""
    B, T, _ = enc_seq.shape
    time = torch.arange(T, dtype=enc_seq.dtype)[None, :, None].repeat(B, 1, 1)
    x = torch.cat([enc_seq, time], 2).transpose(1, 2)
    p = "inf_encoder.net."
    x = F.leaky_relu(F.conv1d(x, sd[p + "input.conv.weight"], sd[p + "input.conv.bias"], padding=1), LRELU)
    x = F.conv1d(x, sd[p + "pyramid-0.conv.weight"], None, padding=1)
    x = F.leaky_relu(F.group_norm(x, 8, sd[p + "pyramid-0.norm.weight"], sd[p + "pyramid-0.norm.bias"], GN_EPS), LRELU)
    x = F.conv1d(x, sd[p + "head.conv.weight"], sd[p + "head.conv.bias"], padding=1)
    return x.transpose(1, 2)


def match_tables(end_ind, T=200):
    """BalancedBinding (gcp/prediction/models/tree/frame_binding.py:42-65) for every sequence: per depth-first node
    its matched frame index and whether it is bound (c_n_prime row non-zero); per frame the depth-first node bound
    to it (frames past end_ind: -1).  Integer recursion shared with the planner-side pruning (balanced_keep_mask)."""
    B = len(end_ind)
    keep = np.zeros((B, N_NODES), dtype=bool)
    tstep = np.zeros((B, N_NODES), dtype=np.int64)
    frame_node = -np.ones((B, T), dtype=np.int64)
    for b in range(B):
        keep[b], tstep[b] = balanced_keep_mask(int(end_ind[b]))
        for i in np.nonzero(keep[b])[0]:
            frame_node[b, tstep[b, i]] = i
    return keep, tstep, frame_node


def tree_inference(sd, e0, eg, inf_seq, tstep, eps):
    """Training-phase tree: TreeModule.produce_subgoal with the approximate posterior (tree_module.py:67-114,
    tree/inference.py:16-36): p = prior(e_l, e_r); q = q(e_l, e_r, inf_enc_seq[b, match_timestep]);
    z = mu_q + sigma_q * eps; then the TreeLSTM exactly as in the prior rollout (gcp_oracle.tree_rollout).
    e0, eg [B,128]; inf_seq [B,T,128]; tstep [B,255] (depth-first); eps [B,255,256] (depth-first)."""
    B = e0.shape[0]
    eL, eR = e0[:, None], eg[:, None]
    hL = hR = None
    out = dict(e=torch.zeros(B, N_NODES, 128), p_mu=torch.zeros(B, N_NODES, 256), p_ls=torch.zeros(B, N_NODES, 256),
               q_mu=torch.zeros(B, N_NODES, 256), q_ls=torch.zeros(B, N_NODES, 256))
    tstep = torch.as_tensor(tstep)
    for lvl in range(DEPTH):
        n = 2 ** lvl
        tm = "tree_module.tree_modules.%d." % lvl
        idx = [df_index(lvl, j) for j in range(n)]
        el, er = eL.reshape(B * n, 128), eR.reshape(B * n, 128)
        pz = mlp(sd, tm + "prior", torch.cat([el, er], 1))
        e_tilde = torch.gather(inf_seq, 1, tstep[:, idx][:, :, None].expand(B, n, 128)).reshape(B * n, 128)
        qz = mlp(sd, tm + "inference.q", torch.cat([el, er, e_tilde], 1))
        q_mu, q_ls = qz[:, :256], qz[:, 256:]
        zeta = q_mu + q_ls.exp() * eps[:, idx].reshape(B * n, 256)                   # Gaussian.sample
        if lvl == 0:
            init = mlp(sd, tm + "lstm_initializer.net", torch.cat([el, er, zeta], 1))
            hl, hr = init[:, :3072], init[:, 3072:]
        else:
            hl, hr = hL.reshape(B * n, 3072), hR.reshape(B * n, 3072)
        sp = tm + "subgoal_pred."
        s = [F.linear(torch.cat([hl[:, 512 * k:512 * (k + 1)], hr[:, 512 * k:512 * (k + 1)]], 1),
                      sd[sp + "projections.%d.weight" % k], sd[sp + "projections.%d.bias" % k]) for k in range(6)]
        x = F.linear(torch.cat([el, er, zeta, e0.repeat_interleave(n, 0), eg.repeat_interleave(n, 0)], 1),
                     sd[sp + "embed.weight"], sd[sp + "embed.bias"])
        new_state = []
        for i in range(3):
            h, c = lstm_cell(x, s[2 * i], s[2 * i + 1], sd[sp + "lstm.%d.weight_ih" % i], sd[sp + "lstm.%d.weight_hh" % i],
                             sd[sp + "lstm.%d.bias_ih" % i], sd[sp + "lstm.%d.bias_hh" % i])
            new_state += [h, c]
            x = h
        e = F.linear(x, sd[sp + "output.weight"], sd[sp + "output.bias"])
        hid = torch.cat(new_state, 1)
        out["e"][:, idx] = e.reshape(B, n, 128)
        for k, v in (("p_mu", pz[:, :256]), ("p_ls", pz[:, 256:]), ("q_mu", q_mu), ("q_ls", q_ls)):
            out[k][:, idx] = v.reshape(B, n, 256)
        e3, h3 = e.reshape(B, n, 128), hid.reshape(B, n, 3072)
        if lvl == 0:
            hL, hR = hl.reshape(B, 1, 3072), hr.reshape(B, 1, 3072)
        eL, eR = interleave(eL, e3), interleave(e3, eR)
        hL, hR = interleave(hL, h3), interleave(h3, hR)
    return out


def decoder_train(sd, lat, s0, s2):
    """DecoderModule on ALL node latents of the batch in one call, batch-statistic BN (encoder_decoder.py:56-97,
    150-218,341-372).  lat [M,128]; s0 [M,16,16,16]; s2 [M,64,4,4].  Returns the DLM parameters
    mu [M,5,3,32,32] (after HalfSigmoid), log_sigma [M,5,3,32,32] and the mean image [M,3,32,32]."""
    p = "decoder.net.net."
    x = F.conv_transpose2d(lat[:, :, None, None], sd[p + "net.conv.weight"])
    x = F.relu(_bn_train(sd, p + "net.norm", x))
    x = torch.cat([x, s2], 1)
    x = F.relu(_bn_train(sd, p + "pyramid-1.norm", _up_pad_conv(x, sd[p + "pyramid-1.conv.weight"], None)))
    x = F.relu(_bn_train(sd, p + "pyramid-0.norm", _up_pad_conv(x, sd[p + "pyramid-0.conv.weight"], None)))
    x = torch.cat([x, s0], 1)
    feat = torch.tanh(_up_pad_conv(x, sd[p + "additional_conv_layer.conv.weight"], sd[p + "additional_conv_layer.conv.bias"]))
    y = F.conv2d(F.pad(feat, (1, 2, 1, 2)), sd["decoder.net.gen_head.conv.weight"], sd["decoder.net.gen_head.conv.bias"])
    mu = torch.sigmoid(y[:, :15]).reshape(-1, 5, 3, 32, 32)
    ls = y[:, 15:].reshape(-1, 5, 3, 32, 32)
    return mu, ls, mu.mean(1) * 2 - 1


def dlm_nll(mu, log_sigma, x):
    """ImageDLM.nll -> DiscreteLogisticMixture.nll -> DiscreteLogistic.prob (blox/torch/encoder_decoder.py:150-156,
    blox/torch/dist.py:87-130,178-197).  mu, log_sigma [M,5,3,32,32]; x [M,3,32,32] in [-1,1].  Returns [M,3,32,32]."""
    x = ((x + 1) / 2)[:, None].expand_as(mu)
    binsize = 1 / 256.0
    scale = torch.exp(log_sigma)
    xs = (torch.floor(x / binsize) * binsize - mu) / scale
    hi = torch.sigmoid(xs + binsize / scale)
    lo = torch.sigmoid(xs)
    p = hi - lo
    mb, mt = (x == 0).float(), (x == 1).float()
    p = hi * mb + p * (1 - mb)
    p = (1 - lo) * mt + p * (1 - mt)
    return -(p.mean(1) + 1e-7).log()


def kl_gauss(q_mu, q_ls, p_mu, p_ls):
    """Gaussian.kl_divergence(q, p) (blox/torch/dist.py:249-252)."""
    return (p_ls - q_ls) + (torch.exp(q_ls) ** 2 + (q_mu - p_mu) ** 2) / (2 * torch.exp(p_ls) ** 2) - 0.5


def forward_loss(sd, batch, aux):
    """BaseGCPModel.forward(phase='train') + TreeModel.loss + get_total_loss (gcp/prediction/models/base_gcp.py:
    140-161,184-304; tree/tree.py:42-79; tree/tree_module.py:116-157; tree/frame_binding.py:80-100).

    batch: traj_seq [B,T,3,32,32], pad_mask [B,T], end_ind [B], states [B,T,2], actions [B,T-1,2], I_0, I_g, eps.
    aux: inv_t0, inv_t1, cost_start, cost_end [B] ints and cost_target [B,1] (see make_golden_train.py).
    Returns dict(losses={name: float tensor}, plus intermediates)."""
    traj, pad, end_ind = batch["traj_seq"], batch["pad_mask"], batch["end_ind"]
    B, T = traj.shape[:2]
    out = {}
    # ---- run_encoder (base_gcp.py:184-209): three separate encoder calls = three sets of batch statistics
    enc_seq, _ = encoder_train(sd, traj.reshape(B * T, 3, 32, 32))
    enc_seq = enc_seq.reshape(B, T, 128)
    e0, (s0, s2) = encoder_train(sd, batch["I_0"])
    eg, _ = encoder_train(sd, batch["I_g"])
    inf_seq = inf_encoder(sd, enc_seq)
    out.update(e0=e0, eg=eg, s0=s0, s2=s2, enc_traj_seq=enc_seq, inf_enc_seq=inf_seq)
    # ---- length predictor (misc.py:38-57)
    logits = mlp(sd, "length_pred.p", torch.cat([e0, eg], 1))
    out["seq_len_logits"] = logits
    # ---- tree with the approximate posterior
    keep, tstep, frame_node = match_tables(end_ind.numpy(), T)
    tree = tree_inference(sd, e0, eg, inf_seq, tstep, batch["eps"])
    out["tree"] = tree
    out.update(keep=keep, tstep=tstep, frame_node=frame_node)
    # ---- decoder over all nodes (TreeDenseRec.forward, tree_dense_rec.py:41-44)
    mu, ls, images = decoder_train(sd, tree["e"].reshape(B * N_NODES, 128), s0.repeat_interleave(N_NODES, 0),
                                   s2.repeat_interleave(N_NODES, 0))
    out["images_df"] = images.reshape(B, N_NODES, 3, 32, 32)
    out["distr_mu"] = mu.reshape(B, N_NODES, 5, 3, 32, 32)
    out["distr_ls"] = ls.reshape(B, N_NODES, 5, 3, 32, 32)
    # ---- existence predictor (frame_binding.py:67-78)
    ex = mlp(sd, "tree_module.tree_modules.0.binding.existence_predictor", tree["e"].reshape(-1, 128)).reshape(B, N_NODES)
    out["existence"] = ex
    # ---- matched pruned latents (base_gcp.py:361-374; evaluation_matching.py:192-206), aux heads (base_gcp.py:234-262)
    fn = torch.as_tensor(np.where(frame_node < 0, 0, frame_node))
    Lmax = int(end_ind.max()) + 1
    seq = torch.gather(tree["e"], 1, fn[:, :Lmax, None].expand(B, Lmax, 128)) * pad[:, :Lmax, None]
    out["model_enc_seq"] = seq
    ar = torch.arange(B)
    t0, t1 = torch.as_tensor(aux["inv_t0"]), torch.as_tensor(aux["inv_t1"])
    # the three auxiliary heads train on DETACHED latents (inverse_mdl.py:30,161-163; base_gcp.py:252-256;
    # cost_mdl.py:53,91-92): no effect on the forward values, decisive for loss_gradients() below
    inv = mlp(sd, "inv_mdl.action_pred", torch.cat([enc_seq[ar, t0].detach(), seq[ar, t1].detach()], 1), conv=False)  # inverse_mdl.py:139-170
    out["inv_actions"] = inv
    reg = mlp(sd, "state_regressor", seq.detach().reshape(-1, 128), conv=False).reshape(B, Lmax, 2)
    out["regressed_state"] = reg
    cs, ce = torch.as_tensor(aux["cost_start"]), torch.as_tensor(aux["cost_end"])
    cost = mlp(sd, "cost_mdl.cost_pred", torch.cat([seq[ar, cs].detach(), seq[ar, ce].detach()], 1), conv=False)   # cost_mdl.py:55-67
    out["cost_pred"] = cost
    # ---- losses
    L = {}
    L["len_pred"] = F.cross_entropy(logits, end_ind)                                                    # misc.py:53-57
    L["action_reconst"] = ((inv - batch["actions"][ar, t0]) ** 2).mean()                               # inverse_mdl.py:180-190
    L["cost_estimation"] = ((cost - torch.as_tensor(aux["cost_target"]).float()) ** 2).mean()           # cost_mdl.py:69-72
    L["state_regression"] = (((reg - batch["states"][:, :Lmax]) ** 2) * pad[:, :Lmax, None]).mean()     # base_gcp.py:284-288
    # reconstruction: frame t is explained by the node bound to it (frames past end_ind: bf node 0 = root, weight 0)
    root = df_index(0, 0)
    fn_all = torch.as_tensor(np.where(frame_node < 0, root, frame_node))
    sel = lambda a: torch.gather(a, 1, fn_all.reshape(B, T, 1, 1, 1, 1).expand(B, T, 5, 3, 32, 32)).reshape(B * T, 5, 3, 32, 32)
    nll = dlm_nll(sel(out["distr_mu"]), sel(out["distr_ls"]), traj.reshape(B * T, 3, 32, 32)).reshape(B, T, 3, 32, 32)
    nll = nll * pad[:, :, None, None, None]
    out["nll_per_frame"] = nll.sum((2, 3, 4))
    L["dense_img_rec"] = nll.sum((1, 2, 3, 4)).mean()                                                   # encoder_decoder.py:220-232
    kl = kl_gauss(tree["q_mu"], tree["q_ls"], tree["p_mu"], tree["p_ls"])
    out["kl_per_node"] = kl.sum(2)                                                                      # depth-first
    L["kl"] = kl.sum((1, 2)).mean()                                                                     # losses.py:75-109
    L["existence_predictor"] = F.binary_cross_entropy_with_logits(ex, torch.as_tensor(keep).float())    # frame_binding.py:80-86
    L["entropy"] = torch.zeros(())                # safe_entropy of one-hot matching = 0, weight 0 (tree_module.py:127)
    # get_total_loss (base_gcp.py:290-301): all listed losses have weight 1 except entropy (0) and nll (0)
    total = sum(L[k] for k in LOSS_NAMES[:7]) / float(np.prod(traj.shape[1:]))
    L["total"] = total
    out["losses"] = L
    return out


def loss_gradients(sd, batch, aux):
    """d total / d parameter for every floating-point tensor of the state dict the training-phase forward uses: what
    `losses.total.value.backward()` leaves in `.grad` in the reference's training step (train.py:155-160).  Groundwork
    for the backward pass: it pins this restatement's autograd graph -- in particular
    the reference's three detach points -- against gradients recorded from the unmodified reference
    (tests/golden/train_grads_B2.npz, oracle/make_golden_train_grad.py).  Returns (losses, {key: grad})."""
    leaf = {}
    uniq = {}
    for k, v in sd.items():
        if not torch.is_floating_point(v) or k.endswith(("running_mean", "running_var")):
            leaf[k] = v
            continue
        # aliased entries (the decoder is registered under three names) must share one leaf
        key = (v.data_ptr(), tuple(v.shape))
        if key not in uniq:
            uniq[key] = v.detach().clone().requires_grad_(True)
        leaf[k] = uniq[key]
    out = forward_loss(leaf, batch, aux)
    # stage targets for a device backward pass: d total / d (intermediate) of the tensors a staged implementation hands
    # from one kernel group to the next (node latents, posterior / prior parameters, DLM head outputs, encoder outputs)
    stages = dict(e_df=out["tree"]["e"], q_mu=out["tree"]["q_mu"], q_log_sigma=out["tree"]["q_ls"], p_mu=out["tree"]["p_mu"],
                  p_log_sigma=out["tree"]["p_ls"], distr_mu=out["distr_mu"], distr_log_sigma=out["distr_ls"],
                  enc_traj_seq=out["enc_traj_seq"], inf_enc_seq=out["inf_enc_seq"], e_0=out["e0"], e_g=out["eg"],
                  skip0=out["s0"], skip2=out["s2"], seq_len_logits=out["seq_len_logits"], existence=out["existence"])
    for t in stages.values():
        if t.requires_grad:
            t.retain_grad()
    out["losses"]["total"].backward()
    grads = {k: v.grad for k, v in leaf.items() if isinstance(v, torch.Tensor) and v.requires_grad and v.grad is not None}
    grads["__stages__"] = {k: (t.grad if t.grad is not None else torch.zeros_like(t)) for k, t in stages.items()}
    return out["losses"], grads
