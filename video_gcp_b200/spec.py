"""Parameter inventory of the GCP-tree model, keyed exactly like the reference's `state_dict()`.

The reference registers the decoder under several aliases and lets the LSTM initializer keep a handle to
the TreeLSTM cell (gcp/prediction/models/tree/tree.py:15-24, tree_module.py:28-65,
blox/torch/recurrent_modules.py:297-303), so one tensor shows up under several keys; `aliases()`
reproduces that.  `canonical_entries()` lists every distinct tensor once.

Nothing here is copied from a dump of the reference: the names are derived from the same
hyper-parameters by the same construction rules; `tests/test_spec.py` checks the result against a
manifest generated from the reference.
"""
import math
from collections import OrderedDict

# kinds drive initialisation (model.py) and synthetic weights (synthetic.py)
W, B_, BN_W, BN_B, BN_RM, BN_RV, BN_NBT, GN_W, GN_B, LSTM_W, LSTM_B, ZEROS = (
    "w", "b", "bn_w", "bn_b", "bn_rm", "bn_rv", "bn_nbt", "gn_w", "gn_b", "lstm_w", "lstm_b", "zeros")
ONES = "ones"


def _bn(out, prefix, c):
    out[prefix + ".weight"] = ((c,), BN_W)
    out[prefix + ".bias"] = ((c,), BN_B)
    out[prefix + ".running_mean"] = ((c,), BN_RM)
    out[prefix + ".running_var"] = ((c,), BN_RV)
    out[prefix + ".num_batches_tracked"] = ((), BN_NBT)


def _predictor(out, prefix, d_in, d_mid, d_out, n_layers, conv, ksz=(3, 3)):
    """BaseProcessingNet naming (blox/torch/layers.py:219-238)."""
    lname = "conv" if conv else "linear"
    tail = tuple(ksz) if conv else ()
    out["%s.input.%s.weight" % (prefix, lname)] = ((d_mid, d_in) + tail, W)
    out["%s.input.%s.bias" % (prefix, lname)] = ((d_mid,), B_)
    for i in range(n_layers):
        out["%s.pyramid-%d.%s.weight" % (prefix, i, lname)] = ((d_mid, d_mid) + tail, W)
        out["%s.pyramid-%d.norm.weight" % (prefix, i)] = ((d_mid,), GN_W)
        out["%s.pyramid-%d.norm.bias" % (prefix, i)] = ((d_mid,), GN_B)
    out["%s.head.%s.weight" % (prefix, lname)] = ((d_out, d_mid) + tail, W)
    out["%s.head.%s.bias" % (prefix, lname)] = ((d_out,), B_)


def _init_lstm_cell(out, prefix, d_in, d_out, H, L, reset_in, nz_mid):
    """InitLSTMCell naming (blox/torch/recurrent_modules.py:126-162,239-253): CustomLSTMCell + init_module."""
    out[prefix + ".initial_hidden"] = ((1, 2 * H * L), ZEROS)
    out[prefix + ".embed.weight"] = ((H, d_in), W)
    out[prefix + ".embed.bias"] = ((H,), B_)
    for i in range(L):
        out[prefix + ".lstm.%d.weight_ih" % i] = ((4 * H, H), LSTM_W)
        out[prefix + ".lstm.%d.weight_hh" % i] = ((4 * H, H), LSTM_W)
        out[prefix + ".lstm.%d.bias_ih" % i] = ((4 * H,), LSTM_B)
        out[prefix + ".lstm.%d.bias_hh" % i] = ((4 * H,), LSTM_B)
    out[prefix + ".output.weight"] = ((d_out, H), W)
    out[prefix + ".output.bias"] = ((d_out,), B_)
    _predictor(out, prefix + ".init_module", reset_in, nz_mid, 2 * H * L, 1, False)


def n_conv_layers(img_sz):
    n = math.log2(img_sz)
    assert n == round(n) and n >= 3
    return int(n)


def decoder_entries(hp, prefix="decoder"):
    out = OrderedDict()
    n = n_conv_layers(hp.img_sz)
    p = prefix + ".net.net."
    top = hp.ngf * 2 ** (n - 3)
    out[p + "net.conv.weight"] = ((hp.nz_enc, top, 4, 4), W)          # ConvTranspose2d: [in, out, k, k]
    _bn(out, p + "net.norm", top)
    for i in reversed(range(n - 3)):
        f_out = hp.ngf * 2 ** i
        f_in = f_out * 2
        if hp.use_skips and (i + 1) % hp.skips_stride == 0:
            f_in *= 2
        out[p + "pyramid-%d.conv.weight" % i] = ((f_out, f_in, 4, 4), W)
        _bn(out, p + "pyramid-%d.norm" % i, f_out)
    f_in = hp.ngf * (2 if hp.use_skips and 0 % hp.skips_stride == 0 else 1)
    out[p + "additional_conv_layer.conv.weight"] = ((hp.ngf, f_in, 4, 4), W)
    out[p + "additional_conv_layer.conv.bias"] = ((hp.ngf,), B_)
    n_out = 10 * hp.input_nc if hp.decoder_distribution == "discrete_logistic_mixture" else hp.input_nc
    out[prefix + ".net.gen_head.conv.weight"] = ((n_out, hp.ngf, 4, 4), W)
    out[prefix + ".net.gen_head.conv.bias"] = ((n_out,), B_)
    if hp.add_weighted_pixel_copy:
        # PixelCopyDecoder (blox/torch/encoder_decoder.py:235-259): mask over [I_0, I_g, generated]
        out[prefix + ".net.mask_head.conv.weight"] = ((3, hp.ngf, 4, 4), W)
        out[prefix + ".net.mask_head.conv.bias"] = ((3,), B_)
    if hp.decoder_distribution == "gaussian":
        # ProbabilisticConvDecoder (encoder_decoder.py:171-173): constant output log-sigma + its updater handle
        out[prefix + ".log_sigma"] = ((), ZEROS)
        out[prefix + ".sigma_updater.parameter"] = ((), ZEROS)
    return out


def canonical_entries(hp):
    """OrderedDict key -> (shape, kind) of every distinct tensor."""
    out = OrderedDict()
    n = n_conv_layers(hp.img_sz)
    # ---- encoder (blox/torch/encoder_decoder.py:31-53)
    p = "encoder.net.net."
    out[p + "input.conv.weight"] = ((hp.ngf, hp.input_nc, 4, 4), W)
    out[p + "input.conv.bias"] = ((hp.ngf,), B_)
    for i in range(n - 3):
        f_in = hp.ngf * 2 ** i
        out[p + "pyramid-%d.conv.weight" % i] = ((2 * f_in, f_in, 4, 4), W)
        _bn(out, p + "pyramid-%d.norm" % i, 2 * f_in)
    out[p + "head.weight"] = ((hp.nz_enc, hp.ngf * 2 ** (n - 3), 4, 4), W)
    out[p + "head.bias"] = ((hp.nz_enc,), B_)
    # ---- decoder
    out.update(decoder_entries(hp))
    # ---- training-time inference encoders (unused in rollout; kept for checkpoint compatibility)
    if hp.seq_enc == "conv":
        k = (hp.conv_inf_enc_kernel_size,)
        _predictor(out, "inf_encoder.net", hp.nz_enc + 1, hp.nz_mid, hp.nz_enc, hp.conv_inf_enc_layers, True, k)
        _predictor(out, "inf_key_encoder.0.net", hp.nz_enc + 1, hp.nz_mid, hp.nz_enc,
                   hp.conv_inf_enc_layers, True, k)
    _predictor(out, "inf_key_encoder.1.net", hp.nz_enc, hp.nz_mid, hp.nz_attn_key, 1, True)
    # ---- heads
    if hp.regress_length:
        _predictor(out, "length_pred.p", 2 * hp.nz_enc, hp.nz_mid, hp.max_seq_len, hp.n_processing_layers, True)
    if hp.attach_inv_mdl:
        n_act = hp.inv_mdl_params["n_actions"]
        _predictor(out, "inv_mdl.action_pred", 2 * 128, 128, n_act, 3, False)   # InverseModel defaults
    if hp.attach_cost_mdl:
        _predictor(out, "cost_mdl.cost_pred", 2 * 128, 128, 1, 3, False)        # CostModel defaults
    if hp.attach_state_regressor:
        _predictor(out, "state_regressor", hp.nz_enc, hp.nz_mid, hp.state_dim, hp.n_processing_layers, False)
    H, L = hp.nz_mid_lstm, hp.n_lstm_layers
    if hp.dense_rec_type == "svg":
        # ---- sequential GCP: VRNNCell (blox/torch/models/vrnn.py:24-52) behind SequentialRecModule
        # (gcp/prediction/models/sequential.py:15-31); no tree modules
        cell = "dense_rec.lstm.cell."
        ctx = 2 * hp.nz_enc if hp.context_every_step else 0
        for name, d_in in (("inf_lstm", hp.nz_enc + ctx), ("gen_lstm", hp.nz_enc + hp.nz_vae + ctx)):
            _init_lstm_cell(out, cell + name, d_in, hp.nz_enc, H, L, 2 * hp.nz_enc, hp.nz_mid)
        _predictor(out, cell + "inf", hp.nz_enc, hp.nz_mid, 2 * hp.nz_vae, hp.n_processing_layers, True)
        _predictor(out, cell + "prior", hp.nz_enc, hp.nz_mid, 2 * hp.nz_vae, hp.n_processing_layers, True)
        return out
    # ---- tree modules (one per level when untied)
    n_mod = hp.hierarchy_levels if hp.untied_layers else 1
    state_dim = 2 * H * L
    pred_in = 2 * hp.nz_enc + hp.nz_vae + (2 * hp.nz_enc if hp.context_every_step else 0)
    for k in range(n_mod):
        tm = "tree_module.tree_modules.%d." % k if hp.untied_layers else "tree_module."
        _predictor(out, tm + "prior", 2 * hp.nz_enc, hp.nz_mid, 2 * hp.nz_vae, hp.n_processing_layers, True)
        _predictor(out, tm + "inference.q", 3 * hp.nz_enc, hp.nz_mid, 2 * hp.nz_vae, hp.n_processing_layers, True)
        if hp.attentive_inference:
            # AttentiveInference (gcp/prediction/models/adaptive_binding/attentive_inference.py); training only
            at = tm + "inference.attention."
            _predictor(out, at + "query_net", 2 * hp.nz_enc, hp.nz_mid, hp.nz_attn_key, hp.n_processing_layers, True)
            for i in range(hp.n_attention_layers):
                al = at + "attention_layers.%d." % i
                out[al + "temperature"] = ((1,), ONES)
                for nm, d in (("q_linear", hp.nz_attn_key), ("k_linear", hp.nz_attn_key), ("v_linear", hp.nz_enc),
                              ("out", hp.nz_enc)):
                    out[al + nm + ".weight"] = ((d, d), W)
                    out[al + nm + ".bias"] = ((d,), B_)
                _predictor(out, at + "predictor_layers.%d" % i, hp.nz_enc, hp.nz_mid, hp.nz_attn_key, 2, True)
            out[at + "out.weight"] = ((hp.nz_enc, hp.nz_enc), W)
            out[at + "out.bias"] = ((hp.nz_enc,), B_)
        sp = tm + "subgoal_pred."
        out[sp + "initial_hidden"] = ((1, state_dim), ZEROS)
        out[sp + "embed.weight"] = ((H, pred_in), W)
        out[sp + "embed.bias"] = ((H,), B_)
        for i in range(L):
            out[sp + "lstm.%d.weight_ih" % i] = ((4 * H, H), LSTM_W)
            out[sp + "lstm.%d.weight_hh" % i] = ((4 * H, H), LSTM_W)
            out[sp + "lstm.%d.bias_ih" % i] = ((4 * H,), LSTM_B)
            out[sp + "lstm.%d.bias_hh" % i] = ((4 * H,), LSTM_B)
        out[sp + "output.weight"] = ((hp.nz_enc, H), W)
        out[sp + "output.bias"] = ((hp.nz_enc,), B_)
        for i in range(2 * L):
            out[sp + "projections.%d.weight" % i] = ((H, 2 * H), W)
            out[sp + "projections.%d.bias" % i] = ((H,), B_)
        _predictor(out, tm + "lstm_initializer.net", 2 * hp.nz_enc + hp.nz_vae, hp.init_mlp_mid_sz,
                   2 * state_dim, hp.init_mlp_layers, True)
        if "dtw" in hp.matching_type:
            # AdaptiveBinding (gcp/prediction/models/adaptive_binding/adaptive.py:18-31)
            out[tm + "binding.temp"] = ((1,), ONES)
            _predictor(out, tm + "binding.distance_predictor", 2 * hp.nz_enc, hp.nz_mid, 1, hp.n_processing_layers, True)
        else:
            _predictor(out, tm + "binding.existence_predictor", hp.nz_enc, hp.nz_mid, 1, hp.n_processing_layers, True)
    return out


def aliases(hp):
    """List of (alias_prefix, canonical_prefix): every key starting with canonical_prefix also appears
    with alias_prefix substituted."""
    al = [("dense_rec.decoder", "decoder")]
    if hp.dense_rec_type == "svg":
        return al
    n_mod = hp.hierarchy_levels if hp.untied_layers else 1
    for k in range(n_mod):
        tm = "tree_module.tree_modules.%d." % k if hp.untied_layers else "tree_module."
        al.append((tm + "decoder", "decoder"))
        al.append((tm + "binding.decoder", "decoder"))
        al.append((tm + "lstm_initializer._cell", tm + "subgoal_pred"))
    return al


def full_manifest(hp):
    """key -> shape for every key of the reference's state_dict (aliases expanded)."""
    canon = canonical_entries(hp)
    out = OrderedDict((k, tuple(v[0])) for k, v in canon.items())
    for alias, target in aliases(hp):
        for k, v in canon.items():
            if k.startswith(target + "."):
                out[alias + k[len(target):]] = tuple(v[0])
    return out
