"""`TreeModel`: drop-in for the reference's GCP-tree model on the planner path.

Mirrors the surface the planner and simulator use (gcp/prediction/models/base_gcp.py:29-66,140-161;
gcp/prediction/models/tree/tree.py:14-67; gcp/planning/planner_policy.py:36-51):

    model = TreeModel(params, logger); model.to(device); model.device = device
    CheckpointHandler.load_weights(...) -> model.load_state_dict(state_dict, strict=False); model.eval()
    with model.val_mode():
        out = model(inputs)          # inputs: AttrDict(I_0, I_g, z, start_ind, end_ind); mutated in place
    model.dense_rec.get_sample_with_len(i, L, out, inputs, 'basic'[, name='e_g_prime'])

`state_dict()` has exactly the reference's keys (including its aliases), so reference checkpoints load.
All arithmetic runs in libgcpb200.so on a B200; there is no PyTorch / CPU fallback -- calling the model
without a CUDA device or outside `val_mode()` (training-time inference networks) raises.
"""
from contextlib import contextmanager

import torch
import torch.nn as nn

from . import spec
from .engine import Engine
from .hparams import build_hparams
from .types import AttrDict


class _Node(nn.Module):
    """Anonymous container used to reproduce the reference's dotted parameter names."""


def _attach(root, dotted, tensor, is_buffer):
    parts = dotted.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, _Node())
        mod = mod._modules[p]
    if is_buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], tensor)


def _init_tensor(shape, kind):
    """Random init with the reference's scales (xavier / kaiming-uniform style fan-in scaling, forget-gate
    bias 1, BatchNorm identity statistics)."""
    if kind in (spec.W, spec.LSTM_W):
        fan_in = shape[1] if (len(shape) == 2 or tuple(shape[2:]) == (3, 3)) else int(torch.tensor(shape[1:]).prod())
        a = (3.0 / max(fan_in, 1)) ** 0.5 if kind == spec.W else 1.0 / shape[1] ** 0.5
        return torch.empty(shape).uniform_(-a, a)
    if kind == spec.LSTM_B:
        b = torch.zeros(shape)
        n = shape[0]
        b[n // 4:n // 2] = 1.0
        return b
    if kind in (spec.BN_W, spec.GN_W, spec.BN_RV, spec.ONES):
        return torch.ones(shape)
    if kind == spec.BN_NBT:
        return torch.zeros(shape, dtype=torch.long)
    return torch.zeros(shape)


class TreeView:
    """Read access to the rollout's node tensors the way `outputs.tree.df.* / .bf.*` exposes them
    (gcp/prediction/utils/tree_utils.py:90-102,185-199)."""

    def __init__(self, fields, depth):
        self._fields = fields
        self.depth = depth
        bf = []
        for lvl in range(depth):
            bf += [(2 * j + 1) * 2 ** (depth - 1 - lvl) - 1 for j in range(2 ** lvl)]
        self._bf_index = bf

    @property
    def size(self):
        return 2 ** self.depth - 1

    @property
    def df(self):
        return _Access(self, False)

    @property
    def bf(self):
        return _Access(self, True)


class _Access:
    def __init__(self, tree, bf):
        self._tree, self._bf = tree, bf

    def __getattr__(self, item):
        f = self._tree._fields
        if item not in f:
            raise AttributeError(item)
        t = f[item]
        if self._bf:
            t = t[:, torch.as_tensor(self._tree._bf_index, device=t.device)]
        return t

    __getitem__ = __getattr__


class TreeDenseRec(nn.Module):
    """Stand-in for gcp/prediction/models/tree/tree_dense_rec.py: balanced ('basic') pruning only."""

    def __init__(self, model):
        super().__init__()
        object.__setattr__(self, "_model", model)

    def _pruned(self, outputs, name):
        key = "_pruned_" + name
        if key not in outputs:
            src = outputs.tree.df.images if name == "images" else outputs.tree.df.e_g_prime
            outputs[key] = self._model.engine.prune_gather(src, outputs.end_ind)
        return outputs[key]

    def get_sample_with_len(self, i_ex, length, outputs, inputs, pruning_scheme, name=None):
        if pruning_scheme != "basic":
            raise NotImplementedError("only the balanced 'basic' pruning scheme is on the planner path")
        name = "images" if name is None else name
        seq = self._pruned(outputs, name)[i_ex, :int(outputs.end_ind[i_ex]) + 1]
        shape = (3, 32, 32) if name == "images" else (seq.shape[-1], 1, 1)
        return seq.reshape(seq.shape[0], *shape), None

    def get_all_samples_with_len(self, length, outputs, inputs, pruning_scheme, name=None):
        if pruning_scheme != "basic":
            raise NotImplementedError("only the balanced 'basic' pruning scheme is on the planner path")
        name = "images" if name is None else name
        buf = self._pruned(outputs, name)
        shape = (3, 32, 32) if name == "images" else (buf.shape[-1], 1, 1)
        ends = outputs.end_ind.tolist()
        return [buf[i, :e + 1].reshape(e + 1, *shape) for i, e in enumerate(ends)], None


class _GCPModelBase(nn.Module):
    """Shared surface of the reference's BaseGCPModel on the planner path (base_gcp.py:29-66,140-161): parameters
    registered under the reference's state-dict keys, lazy engine + weight packing, `val_mode()`."""

    ENGINE_KIND = "tree"

    def _check_config(self, hp):
        raise NotImplementedError

    def _make_dense_rec(self):
        raise NotImplementedError

    def __init__(self, params, logger=None, max_candidates=1024):
        super().__init__()
        self._logger = logger
        self._hp = build_hparams(params)
        hp = self._hp
        self._check_config(hp)
        canon = spec.canonical_entries(hp)
        tensors = {}
        for key, (shape, kind) in canon.items():
            is_buf = kind in (spec.BN_RM, spec.BN_RV, spec.BN_NBT)
            t = _init_tensor(tuple(shape), kind)
            tensors[key] = (t if is_buf else nn.Parameter(t, requires_grad=False), is_buf)
            _attach(self, key, tensors[key][0], is_buf)
        for alias, target in spec.aliases(hp):
            for key in canon:
                if key.startswith(target + "."):
                    _attach(self, alias + key[len(target):], tensors[key][0], tensors[key][1])
        self._canonical_keys = list(canon.keys())
        self.device = torch.device("cpu")
        self._max_candidates = max_candidates
        self._engine = None
        self._dirty = True
        self._val_mode = False
        self._use_pred_length = False
        self.return_prior = False       # also return p_z (mu / log_sigma) per node
        self.return_images = True
        self.inject_end_ind = None      # parity harness: replaces the sampled rollout length
        # True (set by the device-resident simulator around its call): the rollout outputs that the reference pads to
        # the longest sequence of the batch keep their full [B,200] buffers instead, which saves the host
        # synchronisation on `end_ind.max()` in the middle of a CEM step
        self.defer_length_sync = False
        self.seed = 0
        self.__dict__["dense_rec_impl"] = self._make_dense_rec()

    # the reference registers `dense_rec` as a module holding the decoder alias; ours only needs methods
    def __getattr__(self, name):
        if name == "dense_rec" and "dense_rec_impl" in self.__dict__:
            return _DenseRecProxy(self.__dict__["dense_rec_impl"], super().__getattr__("dense_rec"))
        return super().__getattr__(name)

    # ------------------------------------------------------------------------------------------
    @property
    def engine(self):
        if self._engine is None:
            dev = self.device if isinstance(self.device, torch.device) else torch.device(self.device)
            if dev.type != "cuda":
                dev = next(self.parameters()).device
            self._engine = Engine(dev, self._max_candidates, attach_cost_mdl=self._hp.attach_cost_mdl,
                                  model=self.ENGINE_KIND, hierarchy_levels=max(int(self._hp.hierarchy_levels), 2),
                                  max_seq_len=int(self._hp.max_seq_len), tied_layers=not self._hp.untied_layers)
            self._dirty = True
        if self._dirty:
            sd = {k: v for k, v in nn.Module.state_dict(self).items() if k in set(self._canonical_keys)}
            self._engine.load_weights(sd)
            self._dirty = False
        return self._engine

    def load_state_dict(self, state_dict, strict=True, **kw):
        r = super().load_state_dict(state_dict, strict=strict, **kw)
        self._dirty = True
        return r

    def pack_weights(self):
        """Force re-packing after in-place parameter edits."""
        self._dirty = True
        return self.engine

    @contextmanager
    def val_mode(self, pred_length=True):
        self._val_mode, self._use_pred_length = True, pred_length
        try:
            yield
        finally:
            self._val_mode, self._use_pred_length = False, False

    def _rollout_args(self, inputs, n_rows):
        """Checks shared by the rollouts: val_mode only, injected noise, which rollout length to use."""
        if not self._val_mode:
            raise NotImplementedError("only the prior rollout (`with model.val_mode():`) is implemented; the "
                                      "training-time inference path is out of scope for libgcpb200")
        if "z" not in inputs:
            raise NotImplementedError("the rollout needs injected noise `inputs.z` (as the CEM simulator provides)")
        z = inputs.z
        if z.dim() == 5:
            z = z[..., 0, 0]
        assert z.shape[1] == n_rows, "z must be [B,%d,256]" % n_rows
        inject = self.inject_end_ind
        if not (self._use_pred_length and self._hp.length_pred_weight > 0) and "end_ind" in inputs:
            inject = inputs.end_ind      # base_gcp.py:222: predicted length only when _use_pred_length
        return z, inject


class TreeModel(_GCPModelBase):
    """GCP-tree model.  Two configurations are supported, selected by the hyper-parameters exactly as in the
    reference (same class, gcp/prediction/models/tree/tree.py:14): the balanced 25-room planner model
    (mod_hyper.py) and the adaptive-binding model (base_configs/gcp_adaptive.py: pixel-copy decoder +
    distance-predictor pruning, no auxiliary heads)."""
    ENGINE_KIND = "tree"

    def _check_config(self, hp):
        if ("dtw" in hp.matching_type and hp.add_weighted_pixel_copy and hp.decoder_distribution == "gaussian"
                and hp.hierarchy_levels == 8 and hp.nz_enc == 128 and hp.nz_vae == 256 and hp.nz_mid == 128
                and hp.nz_mid_lstm == 512 and hp.n_lstm_layers == 3 and hp.ngf == 16 and hp.img_sz == 32
                and hp.max_seq_len == 200 and hp.untied_layers and hp.tree_lstm == "split_linear"
                and hp.lstm_init == "mlp" and hp.use_skips and not hp.attach_inv_mdl and not hp.attach_state_regressor):
            self.ENGINE_KIND = "tree_adaptive"
            return
        # network sizes are those of the shipped GCP-tree configurations; tree depth, sequence length and tied / untied
        # layers are context parameters: 25-room = 8 levels, 200 frames, untied (experiments/control/25room/gcp_tree/
        # mod_hyper.py:33-54), 9-room = 7 levels, 100 frames, tied (experiments/control/9room/gcp_tree/mod_hyper.py:33-54)
        shape_ok = (2 <= hp.hierarchy_levels <= 8 and hp.max_seq_len % 4 == 0 and 4 <= hp.max_seq_len <= 256
                    and hp.max_seq_len <= 2 ** hp.hierarchy_levels - 1)
        if not (shape_ok and hp.nz_enc == 128 and hp.nz_vae == 256 and hp.nz_mid == 128
                and hp.nz_mid_lstm == 512 and hp.n_lstm_layers == 3 and hp.ngf == 16 and hp.img_sz == 32
                and hp.tree_lstm == "split_linear"
                and hp.lstm_init == "mlp" and hp.matching_type == "balanced" and hp.use_skips
                and hp.decoder_distribution == "discrete_logistic_mixture" and not hp.add_weighted_pixel_copy):
            raise NotImplementedError("libgcpb200 is specialised to the network sizes of the shipped GCP-tree configurations "
                                      "(experiments/control/{25room,9room}/gcp_tree/mod_hyper.py); tree depth 2..8 and "
                                      "max_seq_len (a multiple of 4, <= 2^depth - 1) are free")

    def _make_dense_rec(self):
        return TreeDenseRec(self)

    # ------------------------------------------------------------------------------------------
    # Training-phase forward + loss (BASELINE config 1; train.py:155-157 and the validation pass train.py:204-206)
    def _check_train_config(self):
        hp = self._hp
        if self.ENGINE_KIND != "tree":
            raise NotImplementedError("the training-phase forward is implemented for the balanced 25-room GCP-tree only")
        ok = (hp.hierarchy_levels == 8 and hp.max_seq_len == 200 and hp.untied_layers
              and hp.attach_inv_mdl and hp.attach_state_regressor and hp.attach_cost_mdl and hp.get("run_cost_mdl", True)
              and hp.regress_length and hp.get("seq_enc", "conv") == "conv"
              and not hp.get("supervised_decoder", False) and not hp.get("train_inv_mdl_full_seq", False)
              and hp.kl_weight == 1.0 and hp.length_pred_weight == 1.0 and hp.dense_img_rec_weight == 1.0
              and hp.entropy_weight == 0.0 and not hp.get("kl_weight_burn_in", None) and hp.get("free_nats", 0) == 0)
        if not ok:
            raise NotImplementedError("gcpb200_forward_loss is specialised to the 25-room prediction config "
                                      "(experiments/prediction/25room/gcp_tree/conf.py: inverse model, state regressor "
                                      "and cost model attached, conv sequence encoder, unit loss weights)")

    @staticmethod
    def sample_aux_indices(end_ind, temp_dist=1):
        """The auxiliary heads' frame pairs, drawn from `np.random` with the reference's calls in the reference's order
        (run_auxilliary_models, base_gcp.py:250-260): inverse model first -- B scalar t0 draws, then one vector of B
        offsets (inverse_mdl.py:88-98) -- then per sequence the cost model's start and end index (cost_mdl.py:105-106).
        With the same `np.random.seed` the model therefore trains on the same pairs as the reference."""
        import numpy as np
        end_ind = np.asarray(end_ind).astype(np.int64)
        B = end_ind.shape[0]
        t0 = np.zeros(B, dtype=np.int64)
        for b in range(B):
            assert end_ind[b] >= temp_dist
            t0[b] = np.random.randint(0, end_ind[b] - temp_dist + 1, 1)[0]
        t1 = t0 + np.random.randint(1, temp_dist + 1, B)
        cs, ce = np.zeros(B, dtype=np.int64), np.zeros(B, dtype=np.int64)
        for b in range(B):
            cs[b] = np.random.randint(0, end_ind[b], 1)[0]
            ce[b] = np.random.randint(cs[b] + 1, end_ind[b] + 1, 1)[0]
        return dict(inv_t0=t0, inv_t1=t1, cost_start=cs, cost_end=ce)

    def _forward_train(self, inputs, phase):
        """`model(inputs)` outside val_mode with a ground-truth sequence: BaseGCPModel.forward(phase='train')
        (base_gcp.py:140-161) = run_encoder + length predictor + the tree under the approximate posterior + decoder +
        matching + auxiliary heads, and -- because the device computes them in the same call -- every loss term.
        One gcpb200_forward_loss call; `loss()` / `get_total_loss()` below only hand the results out.

        Randomness: `inputs.eps` ([B,255,256], depth-first) replaces the posterior's N(0,1) draws if present, else the
        device Philox stream (seed = self.seed); `inputs.aux_indices` (dict inv_t0, inv_t1, cost_start, cost_end)
        replaces the np.random draws of `sample_aux_indices`; `inputs.cost_target` ([B,1]) replaces the device-side
        EuclideanPathLength."""
        self._check_train_config()
        eng = self.engine
        dev = eng.device
        traj = inputs.traj_seq
        B, T = traj.shape[:2]
        if B > 128:
            raise NotImplementedError("gcpb200_forward_loss takes at most 128 sequences per call")
        outputs = AttrDict()
        inputs.reference_tensor = traj
        if "start_ind" not in inputs:
            inputs.start_ind = torch.zeros(B, dtype=torch.long, device=dev)
        end_ind = inputs.end_ind
        aux = inputs.get("aux_indices", None)
        if aux is None:
            aux = self.sample_aux_indices(end_ind.detach().cpu().numpy())    # inverse_mdl.py:93 syncs the same way
        if "eps" in inputs:
            eps = inputs.eps
        else:
            eps = eng.sample_noise(B, std_scalar=1.0, seed=self.seed)
            self.seed += 1
        want = ["nll_per_frame", "kl_per_seq", "e_0", "e_g", "enc_traj_seq", "inf_enc_seq", "seq_len_logits", "e_df",
                "p_mu", "p_log_sigma", "q_mu", "q_log_sigma", "match_timesteps", "existence", "model_enc_seq",
                "regressed_state", "inv_actions", "cost_pred"]
        if self.return_images:
            want.append("images_df")
        ct = inputs.get("cost_target", None)
        res = eng.forward_loss(traj, inputs.pad_mask, end_ind, inputs.traj_seq_states, inputs.actions, eps,
                               aux["inv_t0"], aux["inv_t1"], aux["cost_start"], aux["cost_end"], cost_target=ct,
                               I_0=inputs.I_0, I_g=inputs.I_g, want=tuple(want))
        lmax = int(end_ind.max()) + 1
        sp = lambda t: t[..., None, None]
        inputs.e_0, inputs.e_g = sp(res["e_0"]), sp(res["e_g"])
        inputs.enc_traj_seq, inputs.inf_enc_seq = sp(res["enc_traj_seq"]), sp(res["inf_enc_seq"])
        inputs.model_enc_seq = res["model_enc_seq"][:, :lmax]
        outputs.seq_len_logits = res["seq_len_logits"]
        outputs.end_ind = end_ind
        fields = dict(e_g_prime=sp(res["e_df"]), p_z_mu=sp(res["p_mu"]), p_z_log_sigma=sp(res["p_log_sigma"]),
                      q_z_mu=sp(res["q_mu"]), q_z_log_sigma=sp(res["q_log_sigma"]),
                      match_timesteps=res["match_timesteps"])
        if "images_df" in res:
            fields["images"] = res["images_df"]
        outputs.tree = TreeView(fields, self._hp.hierarchy_levels)
        outputs.dense_rec = AttrDict()
        outputs.existence_predictor = AttrDict(existence=res["existence"])
        outputs.actions = res["inv_actions"]
        outputs.regressed_state = res["regressed_state"][:, :lmax]
        outputs.cost = res["cost_pred"][:, None]
        outputs.aux_indices = aux
        outputs.eps = eps
        outputs["_train_losses"] = res["losses"].clone()      # the engine's buffer is overwritten by the next call
        # the per-frame / per-sequence breakdowns are small: own copies, so a validation forward before logging does not
        # change the previous step's `losses.*.breakdown`.  Every other field above is a VIEW of an engine buffer that
        # the next forward of this model overwrites (same lifetime rule as the rollout outputs, see Engine.rollout)
        outputs["_nll_per_frame"], outputs["_kl_per_seq"] = res["nll_per_frame"].clone(), res["kl_per_seq"].clone()
        return outputs

    def loss(self, inputs, outputs, log_error_arr=False):
        """BaseGCPModel.loss + TreeModel.loss (base_gcp.py:264-288, tree.py:70-75, tree_module.py:116-130): the same
        names, `.value` (0-dim device tensor) and `.weight` per term.  The values were reduced on the device by the
        forward call; `breakdown` carries the per-frame NLL / per-sequence KL the reference exposes as error_mat sums."""
        if "_train_losses" not in outputs:
            raise NotImplementedError("loss() needs the outputs of a training-phase forward (model(inputs) outside "
                                      "val_mode with inputs.traj_seq)")
        from . import _C
        vec = outputs["_train_losses"]
        weights = dict(len_pred=self._hp.length_pred_weight, action_reconst=1.0, cost_estimation=1.0,
                       state_regression=1.0, dense_img_rec=self._hp.dense_img_rec_weight, kl=self._hp.kl_weight,
                       existence_predictor=1.0, entropy=self._hp.entropy_weight)
        losses = AttrDict()
        for i, name in enumerate(_C.LOSS_NAMES[:8]):
            losses[name] = AttrDict(value=vec[i], weight=weights[name])
        losses.dense_img_rec.breakdown = outputs["_nll_per_frame"]
        losses.kl.breakdown = outputs["_kl_per_seq"]
        self.__dict__["_last_train_total"] = (losses, {k: v.weight for k, v in losses.items()}, vec[8])
        return losses

    def get_total_loss(self, inputs, losses):
        """base_gcp.py:290-301: sum of value * weight over the positively weighted terms, divided by the number of
        elements of one sequence (T*3*32*32).  The device reduced it with the configuration's weights in the forward
        call; changed weights are refused rather than silently ignored."""
        last = self.__dict__.get("_last_train_total")
        if last is None or last[0] is not losses or any(losses[k].weight != w for k, w in last[1].items()):
            raise NotImplementedError("get_total_loss takes the unmodified result of loss() of the latest forward")
        return AttrDict(value=last[2])

    def step(self):
        pass

    # ------------------------------------------------------------------------------------------
    def forward(self, inputs, phase="train"):
        if not self._val_mode and "traj_seq" in inputs:
            return self._forward_train(inputs, phase)
        z, inject = self._rollout_args(inputs, 2 ** self._hp.hierarchy_levels - 1)
        eng = self.engine
        dev = eng.device
        if z.is_cuda or not (z.dtype == torch.float32 and z.is_pinned()):
            z = z.to(device=dev, dtype=torch.float32)      # pinned fp32 host noise is uploaded by the library itself
        z = z.contiguous()
        B = z.shape[0]
        shared = bool(inputs.get("images_shared", False))
        outputs = AttrDict()
        inputs.reference_tensor = inputs.I_0
        if "start_ind" not in inputs:
            inputs.start_ind = torch.zeros(B, dtype=torch.long, device=dev)
        # planner mode (optional `inputs.planner_mode`, set by GCPImageSimulator.rollout_device for the CEM planner): decode
        # only the nodes balanced pruning keeps, and / or reduce the L2 image cost inside the decoder
        pm = inputs.get("planner_mode", None)
        kw, want_images, want_heads = {}, self.return_images, True
        if pm is not None and self.ENGINE_KIND == "tree":
            want_images = want_images and bool(pm.get("images", True))
            want_heads = bool(pm.get("heads", True))     # False: a cost-only rollout (no existence / action / state heads)
            kept = bool(pm.get("kept_only", True))
            kw = dict(decode_kept_only=kept, want_existence=want_heads, want_aux=want_heads,
                      tree_kept_only=kept and not want_heads and not self.return_prior,
                      sort_sampled_lengths=bool(pm.get("sort_lengths", False)))
            if pm.get("l2", None) is not None:
                kw.update(l2_goal=inputs.I_g[0], l2_dense=bool(pm["l2"][0]), l2_final_step_weight=float(pm["l2"][1]),
                          l2_out=pm.get("l2_out", None))
        res = eng.rollout(inputs.I_0, inputs.I_g, z, end_ind=inject, seed=self.seed, images_shared=shared,
                          want_images=want_images, want_prior=self.return_prior,
                          prune_threshold=self._hp.learned_pruning_threshold, **kw)
        self.seed += 1
        if "l2_cost" in res:
            outputs.l2_cost = res["l2_cost"]
            outputs.l2_spec = (bool(pm["l2"][0]), float(pm["l2"][1]))
        inputs.e_0 = res["e_0"][..., None, None]
        inputs.e_g = res["e_g"][..., None, None]
        outputs.seq_len_logits = res["seq_len_logits"]
        outputs.end_ind = res["end_ind"]
        outputs.z_device = res["z"]
        fields = dict(e_g_prime=res["e_df"][..., None, None])
        if "images_df" in res:
            fields["images"] = res["images_df"]
        if "mu_df" in res:
            fields["p_z_mu"] = res["mu_df"][..., None, None]
            fields["p_z_log_sigma"] = res["log_sigma_df"][..., None, None]
        outputs.tree = TreeView(fields, self._hp.hierarchy_levels)
        outputs.dense_rec = AttrDict()
        if self.ENGINE_KIND == "tree_adaptive":
            # AdaptiveBinding.prune_sequence (adaptive.py:62-77): the kept nodes as index lists; pruned_prediction is
            # materialised from them on first use
            outputs.distance_predictor = AttrDict(distances=res["distances"])
            outputs.pruned_nodes, outputs.pruned_len = res["pruned_nodes"], res["pruned_len"]
            if "images_df" in res:
                outputs.pruned_prediction = _LazyNodePruned(self, outputs, res["images_df"])
            return outputs
        outputs["_lmax"] = lambda e=res["end_ind"]: int(e.max()) + 1      # length the reference pads to (host sync)
        if not want_heads:
            return outputs
        outputs.existence_predictor = AttrDict(existence=res["existence"])
        lmax = self._hp.max_seq_len if self.defer_length_sync else outputs["_lmax"]()
        inputs.model_enc_seq = res["model_enc_seq"][:, :lmax]
        outputs.actions = res["actions"][:, :lmax - 1]
        outputs.regressed_state = res["regressed_state"][:, :lmax]
        outputs["_pruned_e_g_prime"] = res["model_enc_seq"]
        if "images_df" in res:
            outputs.pruned_prediction = _LazyPruned(self, outputs)
        return outputs


class SequentialDenseRec(nn.Module):
    """Sampling API of gcp/prediction/models/sequential.py:78-101 ('basic' scheme): frames are already in time
    order, frame 0 is the start image and `encodings` gets e_0 prepended."""

    def __init__(self, model):
        super().__init__()
        object.__setattr__(self, "_model", model)

    def get_sample_with_len(self, i_ex, length, outputs, inputs, pruning_scheme, name=None):
        if pruning_scheme != "basic":
            raise NotImplementedError("only the 'basic' pruning scheme is on the planner path")
        length = int(length)
        if name is None:
            return outputs.dense_rec.images[i_ex, :length], None
        if name == "encodings":
            return torch.cat((inputs.e_0[i_ex][None], outputs.dense_rec.encodings[i_ex]), 0)[:length], None
        return outputs.dense_rec[name][i_ex, :length], None

    def get_all_samples_with_len(self, end_idxs, outputs, inputs, pruning_scheme, name=None):
        return [self.get_sample_with_len(b, int(end_idxs[b]) + 1, outputs, inputs, pruning_scheme, name=name)[0]
                for b in range(end_idxs.shape[0])], None


class SequentialModel(_GCPModelBase):
    """Drop-in for the reference's sequential GCP model on the planner path
    (gcp/prediction/models/sequential.py:104-134; experiments/prediction/25room/gcp_sequential/conf.py)."""
    ENGINE_KIND = "sequential"

    def _check_config(self, hp):
        if not (hp.dense_rec_type == "svg" and hp.nz_enc == 128 and hp.nz_vae == 256 and hp.nz_mid == 128
                and hp.nz_mid_lstm == 1024 and hp.n_lstm_layers == 3 and hp.ngf == 16 and hp.img_sz == 32
                and hp.max_seq_len == 200 and hp.use_skips and hp.context_every_step and hp.prior_type == "learned"
                and not hp.action_conditioned_pred
                and hp.decoder_distribution == "discrete_logistic_mixture" and not hp.add_weighted_pixel_copy):
            raise NotImplementedError("libgcpb200 is specialised to the 25-room sequential GCP configuration "
                                      "(experiments/prediction/25room/gcp_sequential/conf.py)")

    def _make_dense_rec(self):
        return SequentialDenseRec(self)

    def forward(self, inputs, phase="train"):
        z, inject = self._rollout_args(inputs, self._hp.max_seq_len - 1)
        eng = self.engine
        dev = eng.device
        z = z.to(device=dev, dtype=torch.float32).contiguous()
        B = z.shape[0]
        shared = bool(inputs.get("images_shared", False))
        outputs = AttrDict()
        inputs.reference_tensor = inputs.I_0
        if "start_ind" not in inputs:
            inputs.start_ind = torch.zeros(B, dtype=torch.long, device=dev)
        # phase == 'train' (the simulator's default): model_enc_seq is cut at inputs.end_ind, not at the predicted
        # length (base_gcp.py:238-239 -> get_matched_pruned_seqs); otherwise at outputs.end_ind (sequential.py:133-134)
        given = inputs.end_ind if (phase == "train" and "end_ind" in inputs) else None
        res = eng.seq_rollout(inputs.I_0, inputs.I_g, z, end_ind=inject, given_end_ind=given, seed=self.seed,
                              images_shared=shared, want_images=self.return_images, want_prior=self.return_prior)
        self.seed += 1
        inputs.e_0 = res["e_0"][..., None, None]
        inputs.e_g = res["e_g"][..., None, None]
        outputs.seq_len_logits = res["seq_len_logits"]
        outputs.end_ind = res["end_ind"]
        outputs.z_device = res["z"]
        dr = AttrDict(encodings=res["encodings"][..., None, None])
        if "images" in res:
            dr.images = res["images"]
        if "mu" in res:
            dr.p_z_mu = res["mu"][..., None, None]
            dr.p_z_log_sigma = res["log_sigma"][..., None, None]
        outputs.dense_rec = dr
        if phase != "train":
            raise NotImplementedError("only the simulator's default phase='train' aux path is implemented")
        outputs["_lmax"] = (lambda e=given: int(e.max()) + 1) if given is not None else (lambda: self._hp.max_seq_len)
        lmax = self._hp.max_seq_len if self.defer_length_sync else outputs["_lmax"]()
        inputs.model_enc_seq = res["model_enc_seq"][:, :lmax]
        outputs.actions = res["actions"][:, :lmax - 1]
        outputs.regressed_state = res["regressed_state"][:, :lmax]
        return outputs


class _LazyPruned:
    """`outputs.pruned_prediction`: list of [L_i,3,32,32]; materialised on first use."""

    def __init__(self, model, outputs):
        self._m, self._o, self._v = model, outputs, None

    def _get(self):
        if self._v is None:
            self._v = self._m.dense_rec.get_all_samples_with_len(None, self._o, None, "basic")[0]
        return self._v

    def __iter__(self):
        return iter(self._get())

    def __len__(self):
        return len(self._get())

    def __getitem__(self, i):
        return self._get()[i]


class _LazyNodePruned:
    """`outputs.pruned_prediction` of the adaptive model: list of [L_i,3,32,32]; gathered on first use."""

    def __init__(self, model, outputs, images_df):
        self._m, self._o, self._img, self._v = model, outputs, images_df, None

    def _get(self):
        if self._v is None:
            buf = self._m.engine.gather_nodes(self._img, self._o.pruned_nodes, self._o.pruned_len)
            self._v = [buf[i, :n].reshape(n, 3, 32, 32) for i, n in enumerate(self._o.pruned_len.tolist())]
        return self._v

    def __iter__(self):
        return iter(self._get())

    def __len__(self):
        return len(self._get())

    def __getitem__(self, i):
        return self._get()[i]


class _DenseRecProxy:
    """Gives `model.dense_rec` both the parameter container (state-dict alias) and the sampling API."""

    def __init__(self, impl, container):
        self._impl, self._container = impl, container

    def __getattr__(self, name):
        if hasattr(self._impl, name):
            return getattr(self._impl, name)
        return getattr(self._container, name)
