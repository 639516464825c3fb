"""CPU oracle: a plain torch-fp32 functional restatement of the reference's GCP-tree CEM rollout.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this module, and only as the checker / reported baseline -- the
product package (`video_gcp_b200/`) never imports it and has no CPU fallback.

Parity status: PINNED.  The reference has no tests or golden vectors for this path, so
the oracle is pinned against the reference ITSELF: `oracle/make_golden.py` imports the unmodified
reference from /root/reference (with the import shims in `oracle/refshim.py` and the one documented
semantic patch, truncating integer midpoint), runs it on seeded synthetic weights / inputs / noise and
stores its outputs in `tests/golden/`.  `tests/test_oracle.py` checks every function here against those
fixtures.

Every function cites the reference code it restates (paths relative to /root/reference).  Tensors are
fp32, NCHW; `sd` is a state dict using the reference's key names.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

DEPTH = 8                      # hierarchy_levels (experiments/control/25room/gcp_tree/mod_hyper.py:38)
N_NODES = 2 ** DEPTH - 1       # 255
LRELU = 0.2                    # blox/torch/layers.py:48
BN_EPS = 1e-5
GN_EPS = 1e-5


# --------------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------------
def _lin(sd, name, x, conv_centre):
    """Linear layer; for the conv builder the weight is a 3x3 conv kernel applied to a 1x1 map with
    padding 1, so only the centre tap touches data (blox/torch/layers.py:96-115)."""
    if conv_centre:
        w = sd[name + ".conv.weight"][:, :, 1, 1]
        b = sd.get(name + ".conv.bias")
    else:
        w = sd[name + ".linear.weight"]
        b = sd.get(name + ".linear.bias")
    return F.linear(x, w, b)


def mlp(sd, prefix, x, n_layers=3, conv=True):
    """BaseProcessingNet / Predictor (blox/torch/layers.py:219-238, blox/torch/subnetworks.py:16-29):
    input(+bias, LReLU) -> n x [linear(no bias) -> GroupNorm(8) -> LReLU] -> head(+bias)."""
    x = F.leaky_relu(_lin(sd, prefix + ".input", x, conv), LRELU)
    for i in range(n_layers):
        name = "%s.pyramid-%d" % (prefix, i)
        x = _lin(sd, name, x, conv)
        x = F.group_norm(x, 8, sd[name + ".norm.weight"], sd[name + ".norm.bias"], GN_EPS)
        x = F.leaky_relu(x, LRELU)
    return _lin(sd, prefix + ".head", x, conv)


def _bn(sd, name, x):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                        sd[name + ".weight"], sd[name + ".bias"], False, 0.0, BN_EPS)


def encoder(sd, img):
    """ConvEncoder (blox/torch/encoder_decoder.py:31-53) + GetIntermediatesSequential
    (blox/torch/modules.py:37-52), eval-mode BN.  Returns e [B,128] and skips (s0 [B,16,16,16],
    s2 [B,64,4,4])."""
    p = "encoder.net.net."
    x = F.leaky_relu(F.conv2d(img, sd[p + "input.conv.weight"], sd[p + "input.conv.bias"], 2, 1), LRELU)
    s0 = x
    x = F.leaky_relu(_bn(sd, p + "pyramid-0.norm", F.conv2d(x, sd[p + "pyramid-0.conv.weight"], None, 2, 1)), LRELU)
    x = F.leaky_relu(_bn(sd, p + "pyramid-1.norm", F.conv2d(x, sd[p + "pyramid-1.conv.weight"], None, 2, 1)), LRELU)
    s2 = x
    e = F.conv2d(x, sd[p + "head.weight"], sd[p + "head.bias"])
    return e[:, :, 0, 0], (s0, s2)


def length_logits(sd, e0, eg):
    """LengthPredictorModule.forward (gcp/prediction/models/auxilliary_models/misc.py:38-51)."""
    return mlp(sd, "length_pred.p", torch.cat([e0, eg], 1))


def _up_pad_conv(x, w, b):
    """ConvBlockDec (blox/torch/layers.py:128-150): bilinear x2 -> ZeroPad2d(1,2,1,2) -> conv k4."""
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    return F.conv2d(F.pad(x, (1, 2, 1, 2)), w, b)


def decoder(sd, lat, s0, s2, return_all=False):
    """DecoderModule.forward for the DLM head (blox/torch/encoder_decoder.py:56-97,150-218,341-356;
    SkipInputSequential blox/torch/modules.py:55-68).  lat [M,128]; s0 [M,16,16,16]; s2 [M,64,4,4]
    (already replicated per node as in decode_seq :358-372).  Returns images [M,3,32,32]."""
    p = "decoder.net.net."
    x = F.conv_transpose2d(lat[:, :, None, None], sd[p + "net.conv.weight"])
    x = F.relu(_bn(sd, p + "net.norm", x))                                           # [M,64,4,4]
    x = torch.cat([x, s2], 1)
    x = F.relu(_bn(sd, p + "pyramid-1.norm", _up_pad_conv(x, sd[p + "pyramid-1.conv.weight"], None)))  # [M,32,8,8]
    x = F.relu(_bn(sd, p + "pyramid-0.norm", _up_pad_conv(x, sd[p + "pyramid-0.conv.weight"], None)))  # [M,16,16,16]
    x = torch.cat([x, s0], 1)
    feat = torch.tanh(_up_pad_conv(x, sd[p + "additional_conv_layer.conv.weight"],
                                   sd[p + "additional_conv_layer.conv.bias"]))      # [M,16,32,32]
    y = F.conv2d(F.pad(feat, (1, 2, 1, 2)), sd["decoder.net.gen_head.conv.weight"],
                 sd["decoder.net.gen_head.conv.bias"])                               # [M,30,32,32]
    mu = torch.sigmoid(y[:, :15]).reshape(-1, 5, 3, 32, 32)                          # HalfSigmoid + reshape
    images = mu.mean(1) * 2 - 1                                                      # ImageDLM.mean
    if return_all:
        return images, feat, y
    return images


def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    """torch.nn.LSTMCell, gate order i,f,g,o."""
    g = F.linear(x, w_ih, b_ih) + F.linear(h, w_hh, b_hh)
    i, f, gg, o = g.chunk(4, 1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


def df_index(level, j, depth=DEPTH):
    """In-order (depth-first) index of node j of `level` in a tree of `depth` levels
    (gcp/prediction/utils/tree_utils.py:222-232 depthfirst2layers, inverted)."""
    return (2 * j + 1) * 2 ** (depth - 1 - level) - 1


def tree_module_prefix(sd, level):
    """State-dict prefix of the TreeModule that predicts `level`: one module per level with untied_layers
    (UntiedLayersTree, gcp/prediction/models/tree/untied_layers_tree.py:9-17), else a single `tree_module.`
    (gcp/prediction/models/tree/tree.py:18-21; the 9-room config)."""
    return "tree_module.tree_modules.%d." % level if "tree_module.tree_modules.0.prior.input.conv.weight" in sd else "tree_module."


def interleave(a, b):
    """gcp/prediction/utils/tree_utils.py:202-205 on [B,n,...] tensors."""
    return torch.stack((a, b), 2).reshape(a.shape[0], 2 * a.shape[1], *a.shape[2:])


def tree_rollout(sd, e0, eg, z):
    """SubgoalTreeLayer.produce_tree + TreeModule.produce_subgoal + SplitLinTreeHiddenStatePredictorModel
    (gcp/prediction/utils/tree_utils.py:21-44; gcp/prediction/models/tree/tree_module.py:67-114;
    gcp/prediction/models/tree/tree_lstm.py:30-49; blox/torch/recurrent_modules.py:195-223,286-295).

    e0, eg [B,128]; z [B,2^depth - 1,256] (depth-first node order; the tree depth is read off z: 255 nodes = 8 levels).
    Returns dict of depth-first tensors: e [B,n,128], mu/log_sigma [B,n,256], hidden [B,n,3072].
    """
    B = e0.shape[0]
    n_nodes = z.shape[1]
    depth = int(np.log2(n_nodes + 1))
    assert 2 ** depth - 1 == n_nodes
    eL, eR = e0[:, None], eg[:, None]
    hL = hR = None
    e_df = torch.zeros(B, n_nodes, 128)
    mu_df = torch.zeros(B, n_nodes, 256)
    ls_df = torch.zeros(B, n_nodes, 256)
    h_df = torch.zeros(B, n_nodes, 3072)
    for lvl in range(depth):
        n = 2 ** lvl
        tm = tree_module_prefix(sd, lvl)
        idx = [df_index(lvl, j, depth) for j in range(n)]
        eps = z[:, idx].reshape(B * n, 256)
        el, er = eL.reshape(B * n, 128), eR.reshape(B * n, 128)
        pz = mlp(sd, tm + "prior", torch.cat([el, er], 1))
        mu, log_sigma = pz[:, :256], pz[:, 256:]
        zeta = log_sigma.exp() * eps + mu                                            # Gaussian.reparametrize
        if lvl == 0:
            init = mlp(sd, tm + "lstm_initializer.net", torch.cat([el, er, zeta], 1))
            hl, hr = init[:, :3072], init[:, 3072:]
        else:
            hl, hr = hL.reshape(B * n, 3072), hR.reshape(B * n, 3072)
        sp = tm + "subgoal_pred."
        s = [F.linear(torch.cat([hl[:, 512 * k:512 * (k + 1)], hr[:, 512 * k:512 * (k + 1)]], 1),
                      sd[sp + "projections.%d.weight" % k], sd[sp + "projections.%d.bias" % k])
             for k in range(6)]
        ctx0 = e0.repeat_interleave(n, 0)
        ctxg = eg.repeat_interleave(n, 0)
        x = F.linear(torch.cat([el, er, zeta, ctx0, ctxg], 1), sd[sp + "embed.weight"], sd[sp + "embed.bias"])
        new_state = []
        for i in range(3):
            h, c = lstm_cell(x, s[2 * i], s[2 * i + 1],
                             sd[sp + "lstm.%d.weight_ih" % i], sd[sp + "lstm.%d.weight_hh" % i],
                             sd[sp + "lstm.%d.bias_ih" % i], sd[sp + "lstm.%d.bias_hh" % i])
            new_state += [h, c]
            x = h
        e = F.linear(x, sd[sp + "output.weight"], sd[sp + "output.bias"])
        hid = torch.cat(new_state, 1)
        e_df[:, idx] = e.reshape(B, n, 128)
        mu_df[:, idx] = mu.reshape(B, n, 256)
        ls_df[:, idx] = log_sigma.reshape(B, n, 256)
        h_df[:, idx] = hid.reshape(B, n, 3072)
        e3, h3 = e.reshape(B, n, 128), hid.reshape(B, n, 3072)
        if lvl == 0:
            hL, hR = hl.reshape(B, 1, 3072), hr.reshape(B, 1, 3072)
        eL, eR = interleave(eL, e3), interleave(e3, eR)
        hL, hR = interleave(hL, h3), interleave(h3, hR)
    return dict(e=e_df, mu=mu_df, log_sigma=ls_df, hidden=h_df)


# --------------------------------------------------------------------------------------------------
# integer part: balanced pruning
# --------------------------------------------------------------------------------------------------
def balanced_keep_mask(end_ind, depth=DEPTH):
    """BalancedEvalBinding.get_all_samples + BalancedBinding.__call__/comp_timestep/get_init_inds
    (gcp/evaluation/evaluation_matching.py:192-206; gcp/prediction/models/tree/frame_binding.py:42-65).

    Recursion on integer intervals starting (l, r) = (-1, end_ind + 1); t = trunc((l + r) / 2) (int64
    division, torch-1.3 semantics); the node is kept iff t != l and t != r; children get (l, t), (t, r).
    Returns keep [255] bool and timestep [255] int64, both in depth-first node order.
    """
    keep = np.zeros(2 ** depth - 1, dtype=bool)
    tstep = np.zeros(2 ** depth - 1, dtype=np.int64)

    def rec(l, r, lvl, j):
        if lvl == depth:
            return
        t = int((l + r) / 2)            # C-style truncation toward zero
        i = df_index(lvl, j, depth)
        tstep[i] = t
        keep[i] = (t != l) and (t != r)
        rec(l, t, lvl + 1, 2 * j)
        rec(t, r, lvl + 1, 2 * j + 1)

    rec(-1, int(end_ind) + 1, 0, 0)
    return keep, tstep


def prune_indices(end_ind, depth=DEPTH):
    """Depth-first indices of the kept nodes, in order: exactly end_ind + 1 of them."""
    keep, _ = balanced_keep_mask(end_ind, depth)
    return np.nonzero(keep)[0]


# --------------------------------------------------------------------------------------------------
# full model forward as the planner sees it
# --------------------------------------------------------------------------------------------------
def rollout(sd, I_0, I_g, z, end_ind, decode=True):
    """BaseGCPModel.forward in val_mode with injected z and injected end_ind
    (gcp/prediction/models/base_gcp.py:140-161,184-262; gcp/prediction/models/tree/tree.py:42-67).

    I_0, I_g [B,3,32,32] in [-1,1]; z [B,255,256]; end_ind [B] int.
    Returns dict: e0, eg, seq_len_logits, tree (df tensors), images_df [B,255,3,32,32],
    existence [B,255], pruned images / latents (lists), model_enc_seq [B,Lmax,128], actions
    [B,Lmax-1,2], regressed_state [B,Lmax,2].
    """
    out = {}
    e0, (s0, s2) = encoder(sd, I_0)
    eg, _ = encoder(sd, I_g)
    out["e0"], out["eg"] = e0, eg
    out["seq_len_logits"] = length_logits(sd, e0, eg)
    tree = tree_rollout(sd, e0, eg, z)
    out["tree"] = tree
    B = e0.shape[0]
    n_nodes = z.shape[1]
    depth = int(np.log2(n_nodes + 1))
    if decode:
        lat = tree["e"].reshape(B * n_nodes, 128)
        imgs = decoder(sd, lat, s0.repeat_interleave(n_nodes, 0), s2.repeat_interleave(n_nodes, 0))
        out["images_df"] = imgs.reshape(B, n_nodes, 3, 32, 32)
    # existence predictor (gcp/prediction/models/tree/frame_binding.py:67-78); result unused by pruning
    ex = mlp(sd, tree_module_prefix(sd, 0) + "binding.existence_predictor", tree["e"].reshape(-1, 128))
    out["existence"] = ex.reshape(B, n_nodes)
    # balanced pruning with the (injected) predicted length
    idxs = [prune_indices(int(t), depth) for t in end_ind]
    out["prune_idx"] = idxs
    if decode:
        out["pruned_images"] = [out["images_df"][b, torch.as_tensor(ix)] for b, ix in enumerate(idxs)]
    lat_seqs = [tree["e"][b, torch.as_tensor(ix)] for b, ix in enumerate(idxs)]
    out["pruned_latents"] = lat_seqs
    enc_seq = torch.nn.utils.rnn.pad_sequence(lat_seqs, batch_first=True)            # base_gcp.py:242
    out["model_enc_seq"] = enc_seq
    # inverse model on consecutive pairs (inverse_mdl.py:110-134), state regressor (base_gcp.py:252-256)
    pairs = torch.cat([enc_seq[:, :-1], enc_seq[:, 1:]], 2)
    out["actions"] = mlp(sd, "inv_mdl.action_pred", pairs.reshape(-1, 256), conv=False).reshape(B, -1, 2)
    out["regressed_state"] = mlp(sd, "state_regressor", enc_seq.reshape(-1, 128), conv=False).reshape(B, -1, 2)
    return out


def simulator_rollout(sd, state, goal, samples, end_ind, append_latent=True):
    """GCPImageSimulator.rollout (gcp/planning/cem/cem_simulator.py:14-43,80-96) with injected end_ind.

    state, goal: numpy [1,32,32,3] in [0,1] (or 0..255); samples numpy [B,255,256].
    Returns dict of python lists of numpy arrays: predictions [L,3072(+128)], actions, states, latents.
    """
    B = samples.shape[0]

    def env2planner(img):
        img = torch.tensor(np.repeat(img, B, 0), dtype=torch.float32)
        if img.max() > 1.0:
            img = img / 255.0
        return img.permute(0, 3, 1, 2) * 2 - 1.0

    out = rollout(sd, env2planner(state), env2planner(goal),
                  torch.tensor(samples, dtype=torch.float32), end_ind)
    end = np.maximum(np.asarray(end_ind), 1)
    preds = []
    for b in range(B):
        r = out["pruned_images"][b].reshape(end[b] + 1, -1)
        if append_latent:
            r = torch.cat([r, out["pruned_latents"][b]], -1)
        preds.append(r.numpy())
    cap = lambda v: [v[b, :end[b] + 1].numpy() for b in range(B)]
    return dict(predictions=preds, actions=cap(out["actions"]), states=cap(out["regressed_state"]),
                latents=cap(out["model_enc_seq"]))


def infer_action(sd, current_img, target_latent):
    """ImageCEMPolicy._infer_action (gcp/planning/planner_policy.py:215-221): closed-loop execution step.
    current_img numpy [1,H,W,3] (env range, [0,1] or 0..255); target_latent numpy [128] (next latent of the plan).
    enc = encoder(env2planner(img)) (cem_simulator.py:87-96); action = inv_mdl.action_pred(enc, target)
    (InverseModel.run_single, inverse_mdl.py:221-224).  Returns (action [2], enc [128])."""
    img = torch.as_tensor(np.asarray(current_img), dtype=torch.float32)
    if img.max() > 1.0:
        img = img / 255.0
    img = img.permute(0, 3, 1, 2) * 2 - 1.0
    enc, _ = encoder(sd, img)
    tgt = torch.as_tensor(np.asarray(target_latent), dtype=torch.float32)[None]
    act = mlp(sd, "inv_mdl.action_pred", torch.cat([enc, tgt], 1), conv=False)
    return act[0].numpy(), enc[0].numpy()


# --------------------------------------------------------------------------------------------------
# sequential GCP (config 3): VRNN prior rollout
# --------------------------------------------------------------------------------------------------
def seq_latent_rollout(sd, e0, eg, z, n_lstm=3):
    """SequentialRecModule.forward + VRNNCell.forward/init_state with injected z, prior branch
    (gcp/prediction/models/sequential.py:33-58; blox/torch/models/vrnn.py:54-110;
    blox/torch/recurrent_modules.py:21-53,195-223,239-259).  The inference LSTM / q(z) are computed by the
    reference but never reach the rollout outputs the planner reads, so they are not restated.

    e0, eg [B,128]; z [B,T,256].  Per step t: p_z = prior(x_t); zeta = exp(log_sigma) * z_t + mu;
    x_{t+1} = gen_lstm(cat(x_t, zeta, e0, eg)); x_0 = e0; LSTM state = init_module(cat(e0, eg)), laid out
    [h0,c0,h1,c1,h2,c2] (var2state, recurrent_modules.py:183-187).
    Returns encodings [B,T,128], mu/log_sigma [B,T,256]."""
    cell = "dense_rec.lstm.cell."
    g = cell + "gen_lstm."
    ctx = torch.cat([e0, eg], 1)
    state = mlp(sd, g + "init_module", ctx, n_layers=1, conv=False)
    H = state.shape[1] // (2 * n_lstm)
    hs = [state[:, 2 * i * H:(2 * i + 1) * H] for i in range(n_lstm)]
    cs = [state[:, (2 * i + 1) * H:(2 * i + 2) * H] for i in range(n_lstm)]
    x = e0
    enc, mus, lss = [], [], []
    for t in range(z.shape[1]):
        pz = mlp(sd, cell + "prior", x)
        mu, log_sigma = pz[:, :256], pz[:, 256:]
        zeta = log_sigma.exp() * z[:, t] + mu
        h_in = F.linear(torch.cat([x, zeta, ctx], 1), sd[g + "embed.weight"], sd[g + "embed.bias"])
        for i in range(n_lstm):
            hs[i], cs[i] = lstm_cell(h_in, hs[i], cs[i], sd[g + "lstm.%d.weight_ih" % i], sd[g + "lstm.%d.weight_hh" % i],
                                     sd[g + "lstm.%d.bias_ih" % i], sd[g + "lstm.%d.bias_hh" % i])
            h_in = hs[i]
        x = F.linear(h_in, sd[g + "output.weight"], sd[g + "output.bias"])
        enc.append(x)
        mus.append(mu)
        lss.append(log_sigma)
    return dict(encodings=torch.stack(enc, 1), mu=torch.stack(mus, 1), log_sigma=torch.stack(lss, 1))


def seq_rollout(sd, I_0, I_g, z, given_end_ind, decode=True):
    """SequentialModel forward in val_mode with injected z, default phase='train' as the simulator calls it
    (gcp/prediction/models/base_gcp.py:140-161,234-262,361-374; gcp/prediction/models/sequential.py:33-58,
    78-94).  z [B,199,256]; given_end_ind [B] = inputs.end_ind (the simulator passes rollout_len - 1).

    Returns e0, eg, seq_len_logits, encodings [B,199,128], mu, log_sigma, images [B,200,3,32,32] (frame 0 is
    I_0 itself), model_enc_seq = pad(cat(e0, encodings)[:given_end_ind + 1]), actions, regressed_state."""
    out = {}
    e0, (s0, s2) = encoder(sd, I_0)
    eg, _ = encoder(sd, I_g)
    out["e0"], out["eg"] = e0, eg
    out["seq_len_logits"] = length_logits(sd, e0, eg)
    lat = seq_latent_rollout(sd, e0, eg, z)
    out.update(lat)
    B, T = z.shape[:2]
    if decode:
        imgs = decoder(sd, lat["encodings"].reshape(B * T, 128), s0.repeat_interleave(T, 0), s2.repeat_interleave(T, 0))
        out["images"] = torch.cat([I_0[:, None], imgs.reshape(B, T, 3, 32, 32)], 1)
    full = torch.cat([e0[:, None], lat["encodings"]], 1)
    seqs = [full[b, :int(e) + 1] for b, e in enumerate(given_end_ind)]
    enc_seq = torch.nn.utils.rnn.pad_sequence(seqs, batch_first=True)
    out["model_enc_seq"] = enc_seq
    pairs = torch.cat([enc_seq[:, :-1], enc_seq[:, 1:]], 2)
    out["actions"] = mlp(sd, "inv_mdl.action_pred", pairs.reshape(-1, 256), conv=False).reshape(B, -1, 2)
    out["regressed_state"] = mlp(sd, "state_regressor", enc_seq.reshape(-1, 128), conv=False).reshape(B, -1, 2)
    return out


def seq_simulator_rollout(sd, state, goal, samples, end_ind, append_latent=True, rollout_len=200):
    """GCPImageSimulator.rollout over the sequential model (gcp/planning/cem/cem_simulator.py:14-43,80-96) with
    the sampled rollout length replaced by the injected `end_ind`.  samples numpy [B,199,256]."""
    B = samples.shape[0]

    def env2planner(img):
        img = torch.tensor(np.repeat(img, B, 0), dtype=torch.float32)
        if img.max() > 1.0:
            img = img / 255.0
        return img.permute(0, 3, 1, 2) * 2 - 1.0

    out = seq_rollout(sd, env2planner(state), env2planner(goal), torch.tensor(samples, dtype=torch.float32),
                      np.full(B, rollout_len - 1))
    end = np.maximum(np.asarray(end_ind), 1)
    full = torch.cat([out["e0"][:, None], out["encodings"]], 1)
    preds = []
    for b in range(B):
        r = out["images"][b, :end[b] + 1].reshape(end[b] + 1, -1)
        if append_latent:
            r = torch.cat([r, full[b, :end[b] + 1]], -1)
        preds.append(r.numpy())
    cap = lambda v: [v[b, :end[b] + 1].numpy() for b in range(B)]
    return dict(predictions=preds, actions=cap(out["actions"]), states=cap(out["regressed_state"]),
                latents=cap(out["model_enc_seq"]))


# --------------------------------------------------------------------------------------------------
# adaptive-binding GCP-tree (config 4): pixel-copy decoder + distance-predictor pruning
# --------------------------------------------------------------------------------------------------
def decoder_pixel_copy(sd, lat, s0, s2, I_0, I_g, return_all=False):
    """DecoderModule.forward with PixelCopyDecoder and the Gaussian output head (blox/torch/encoder_decoder.py:
    56-97,235-259,197-203,341-356): same trunk as `decoder`; gen = tanh(conv(pad(feat))), mask =
    softmax_channel(conv(pad(feat))), image = mask_0 I_0 + mask_1 I_g + mask_2 gen.
    lat [M,128]; s0, s2, I_0, I_g already replicated per node (decode_seq :358-372)."""
    p = "decoder.net.net."
    x = F.conv_transpose2d(lat[:, :, None, None], sd[p + "net.conv.weight"])
    x = F.relu(_bn(sd, p + "net.norm", x))
    x = torch.cat([x, s2], 1)
    x = F.relu(_bn(sd, p + "pyramid-1.norm", _up_pad_conv(x, sd[p + "pyramid-1.conv.weight"], None)))
    x = F.relu(_bn(sd, p + "pyramid-0.norm", _up_pad_conv(x, sd[p + "pyramid-0.conv.weight"], None)))
    x = torch.cat([x, s0], 1)
    feat = torch.tanh(_up_pad_conv(x, sd[p + "additional_conv_layer.conv.weight"],
                                   sd[p + "additional_conv_layer.conv.bias"]))
    fp = F.pad(feat, (1, 2, 1, 2))
    gen = torch.tanh(F.conv2d(fp, sd["decoder.net.gen_head.conv.weight"], sd["decoder.net.gen_head.conv.bias"]))
    mask = torch.softmax(F.conv2d(fp, sd["decoder.net.mask_head.conv.weight"], sd["decoder.net.mask_head.conv.bias"]), 1)
    images = (mask.unsqueeze(2) * torch.stack([I_0, I_g, gen], 1)).sum(1)          # mask_and_merge :253-259
    if return_all:
        return images, mask, gen
    return images


def adaptive_rollout(sd, I_0, I_g, z, threshold=0.5, decode=True):
    """Adaptive-binding TreeModel forward in val_mode with injected z (gcp/prediction/models/base_gcp.py:140-161;
    tree.py:42-67; adaptive_binding/adaptive.py:62-77).  The tree recursion is the GCP-tree's; images come from
    the pixel-copy decoder; AdaptiveBinding.prune_sequence drops node n > 0 (depth-first order) when
    sigmoid(distance_predictor(e_{n-1}, e_n)) > learned_pruning_threshold.

    Returns e0, eg, seq_len_logits, tree (df tensors), images_df [B,255,3,32,32], distances [B,254],
    keep [B,255] bool, pruned_images / pruned_latents (lists)."""
    out = {}
    e0, (s0, s2) = encoder(sd, I_0)
    eg, _ = encoder(sd, I_g)
    out["e0"], out["eg"] = e0, eg
    out["seq_len_logits"] = length_logits(sd, e0, eg)
    tree = tree_rollout(sd, e0, eg, z)
    out["tree"] = tree
    B = e0.shape[0]
    rep = lambda t: t.repeat_interleave(N_NODES, 0)
    if decode:
        imgs = decoder_pixel_copy(sd, tree["e"].reshape(B * N_NODES, 128), rep(s0), rep(s2), rep(I_0), rep(I_g))
        out["images_df"] = imgs.reshape(B, N_NODES, 3, 32, 32)
    lat = tree["e"]
    pairs = torch.cat([lat[:, :-1], lat[:, 1:]], 2).reshape(-1, 256)
    dist = mlp(sd, "tree_module.tree_modules.0.binding.distance_predictor", pairs).reshape(B, N_NODES - 1)
    out["distances"] = dist
    close = torch.sigmoid(dist) > threshold
    keep = ~torch.cat([torch.zeros_like(close[:, :1]), close], 1)
    out["keep"] = keep
    if decode:
        out["pruned_images"] = [out["images_df"][b][keep[b]] for b in range(B)]
    out["pruned_latents"] = [lat[b][keep[b]] for b in range(B)]
    return out


# --------------------------------------------------------------------------------------------------
# costs, elites, refit
# --------------------------------------------------------------------------------------------------
def l2_image_cost(image_seqs, goal_raw, dense_cost=True, final_step_weight=1.0):
    """L2ImageCost._compute + CostFcn.__call__ (gcp/planning/cem/cost_fcn.py:9-22,65-72).
    image_seqs: list of numpy [L,3,32,32]; goal_raw numpy [1,32,32,3] in [0,1]."""
    goal = goal_raw.transpose(0, 3, 1, 2) * 2 - 1.0
    costs = []
    for seq in image_seqs:
        c = np.sqrt(np.sum((seq - goal) ** 2, axis=(1, 2, 3)))
        c[-1] *= final_step_weight
        costs.append(np.sum(c) if dense_cost else c[-1])
    return np.array(costs)


def learned_cost(sd, latent_seqs, goal_seqs):
    """LearnedCostEstimate.__call__ list branch (gcp/planning/cem/cost_fcn.py:84-97) with
    TestTimeCostModel.forward (gcp/prediction/models/auxilliary_models/cost_mdl.py:138-145):
    sum_t MLP(cat(x_t, x_{t+1})) over cat(latents, goal latents)."""
    costs = []
    for seq, goal in zip(latent_seqs, goal_seqs):
        s = torch.cat([torch.as_tensor(seq), torch.as_tensor(goal)])
        c = mlp(sd, "cost_mdl.cost_pred", torch.cat([s[:-1], s[1:]], 1), conv=False)
        costs.append(float(c.sum()))
    return np.array(costs)


def image_wrapped_learned_cost(sd, latent_seqs):
    """ImageWrappedLearnedCostFcn.__call__ (gcp/planning/cem/cost_fcn.py:108-116): every candidate's
    'goal' is the LAST candidate's whole latent rollout (the reference's own HACK)."""
    return learned_cost(sd, latent_seqs, [latent_seqs[-1]] * len(latent_seqs))


def elites(scores, n_candidates, elite_frac=0.1):
    """CEMPlanner._get_best_rollouts (gcp/planning/cem/cem_planner.py:124-135): argsort, first
    int(batch_size * elite_frac)."""
    k = int(n_candidates * elite_frac)
    return np.argsort(scores, kind="stable")[:k]


def refit(samples, elite_idx):
    """FlatCEMSampler.fit (gcp/planning/cem/sampler.py:44-46): mean / std (ddof=0) over elites."""
    d = samples[elite_idx]
    return d.mean(0), d.std(0)


# --------------------------------------------------------------------------------------------------
# algorithmic work, for roofline arithmetic
# --------------------------------------------------------------------------------------------------
MAC_PER_ROLLOUT = 8.377e9
FLOP_PER_ROLLOUT = 2 * MAC_PER_ROLLOUT
