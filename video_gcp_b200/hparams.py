"""Hyper-parameters of the GCP-tree rollout path.

Mirrors the defaults the reference layers together (gcp/prediction/models/auxilliary_models/
base_model.py:34-70 + gcp/prediction/hyperparameters.py:4-150) for the keys that reach the rollout
path, and its override rule (`override_defaults`, base_model.py:27-32): overriding a key with the value
it already has is an error, unknown keys are an error.
"""
from .types import AttrDict

# defaults of the keys the hot path reads (same values as the reference defaults)
_DEFAULTS = dict(
    batch_size=-1, max_seq_len=-1, n_actions=-1, state_dim=-1, img_sz=32, input_nc=3,
    n_conv_layers=None, use_convs=True, use_batchnorm=True, normalization='batch',
    predictor_normalization='group', checkpt_path=None, dataset_class=None,
    ngf=4, nz_enc=32, nz_vae=32, nz_vae2=256, nz_mid=32, nz_mid_lstm=32, n_lstm_layers=1,
    n_processing_layers=3, conv_inf_enc_kernel_size=3, conv_inf_enc_layers=1,
    n_attention_heads=1, n_attention_layers=1, nz_attn_key=32, init_mlp_layers=3, init_mlp_mid_sz=32,
    action_activation=None, device=None, context_every_step=True,
    kl_weight=1., kl_weight_burn_in=None, entropy_weight=.0, length_pred_weight=1.,
    dense_img_rec_weight=1., dense_action_rec_weight=1., free_nats=0,
    use_skips=True, skips_stride=2, add_weighted_pixel_copy=False, pixel_shift_decoder=False,
    skip_from_parents=False, seq_enc='none', regress_actions=False, learn_attn_temp=True,
    attention_temperature=1.0, attach_inv_mdl=False, attach_cost_mdl=False, run_cost_mdl=True,
    attach_state_regressor=False, action_conditioned_pred=False, learn_beta=True, initial_sigma=1.0,
    separate_cnn_start_goal_encoder=False, decoder_distribution='gaussian', use_conv_lstm=False,
    prior_type='learned', var_inf='standard', hierarchy_levels=3, attentive_inference=False,
    non_goal_conditioned=False, tree_lstm='', lstm_init='zero', matching_type='latent',
    regress_index=False, regress_length=False, inv_mdl_params={}, train_inv_mdl_full_seq=False,
    cost_mdl_params={}, learned_pruning_threshold=0.5, untied_layers=False, supervised_decoder=False,
    states_inference=False, dense_rec_type='none', one_step_planner='discrete', binding='frames',
    randomize_length=False, randomize_start=False, learn_matching_temp=True, matching_temp=1.0,
)


class HParams(AttrDict):
    """Attribute-style hyper-parameter bag with the reference's override semantics."""

    def override_defaults(self, params):
        for name, value in params.items():
            if name not in self:
                raise AttributeError("unknown hyperparameter %r" % name)
            if value == self[name]:
                # same rule as base_model.py:29-30
                raise ValueError("attribute is {} is identical to default value!!".format(name))
            self[name] = value
        return self


def default_hparams():
    return HParams(_DEFAULTS)


def gcp_tree_25room_config(**extra):
    """The 25-room GCP-tree planner model config: experiments/control/25room/gcp_tree/mod_hyper.py:33-55
    layered over experiments/prediction/base_configs/{base_tree,gcp_tree}.py (with
    `add_weighted_pixel_copy` popped as the experiment does)."""
    cfg = AttrDict(
        one_step_planner='sh_pred', binding='loss', seq_enc='conv', tree_lstm='split_linear',
        lstm_init='mlp', dense_rec_type='node_prob', matching_type='balanced',
        state_dim=2, ngf=16, max_seq_len=200, hierarchy_levels=8, nz_mid_lstm=512, n_lstm_layers=3,
        nz_mid=128, nz_enc=128, nz_vae=256, regress_length=True, attach_state_regressor=True,
        attach_inv_mdl=True,
        inv_mdl_params=AttrDict(n_actions=2, use_convs=False, build_encoder=False),
        untied_layers=True, decoder_distribution='discrete_logistic_mixture',
    )
    cfg.update(extra)
    return cfg


def gcp_tree_9room_config(**extra):
    """experiments/control/9room/gcp_tree/mod_hyper.py:33-54: the 9-room planner model -- 7 levels (127 nodes), 100
    frames, and `untied_layers` left at its default False (ONE TreeModule for all levels)."""
    cfg = gcp_tree_25room_config()
    cfg.update(max_seq_len=100, hierarchy_levels=7)
    cfg.pop("untied_layers")            # the 9-room experiment leaves it at the default (False)
    cfg.update(extra)
    return cfg


def gcp_adaptive_25room_config(**extra):
    """The adaptive-binding GCP-tree model with the 25-room network sizes (config 4): experiments/prediction/
    base_configs/gcp_adaptive.py:6-11 over base_tree.py:11-20, sizes of experiments/prediction/25room/gcp_tree/conf.py.
    Pixel-copy decoder with a Gaussian output distribution, distance-predictor pruning, no auxiliary heads."""
    cfg = AttrDict(
        one_step_planner='sh_pred', binding='loss', seq_enc='conv', tree_lstm='split_linear',
        lstm_init='mlp', dense_rec_type='node_prob', add_weighted_pixel_copy=True,
        matching_type='dtw_image', learn_matching_temp=False, attentive_inference=True,
        ngf=16, max_seq_len=200, hierarchy_levels=8, nz_mid_lstm=512, n_lstm_layers=3,
        nz_mid=128, nz_enc=128, nz_vae=256, regress_length=True, untied_layers=True,
    )
    cfg.update(extra)
    return cfg


def gcp_sequential_25room_config(**extra):
    """The 25-room sequential GCP model config: experiments/prediction/25room/gcp_sequential/conf.py:20-43
    layered over experiments/prediction/base_configs/gcp_sequential.py:9-14 (`add_weighted_pixel_copy`
    popped as the experiment does; `attach_cost_mdl` is a training-time head and off by default here)."""
    cfg = AttrDict(
        one_step_planner='continuous', dense_rec_type='svg', hierarchy_levels=0,
        state_dim=2, ngf=16, max_seq_len=200, nz_mid_lstm=1024, n_lstm_layers=3,
        nz_mid=128, nz_enc=128, nz_vae=256, regress_length=True, attach_state_regressor=True,
        attach_inv_mdl=True,
        inv_mdl_params=AttrDict(n_actions=2, use_convs=False, build_encoder=False),
        decoder_distribution='discrete_logistic_mixture',
    )
    cfg.update(extra)
    return cfg


def build_hparams(params):
    """defaults + overrides, the way BaseGCPModel.__init__ does it (base_gcp.py:30-41)."""
    hp = default_hparams()
    hp.override_defaults(params)
    assert hp.batch_size != -1, "batch_size must be overridden (base_gcp.py:36)"
    return hp
