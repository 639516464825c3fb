"""Thin Python owner of one gcpb200 context: weight upload and typed wrappers over the C ABI.

PyTorch is used for device memory, streams and host<->device copies only; every arithmetic step of the
rollout runs inside libgcpb200.so.
"""
import ctypes as C

import torch

from . import _C

N_NODES, NZ_ENC, NZ_VAE, MAX_LEN = 255, 128, 256, 200
SEQ_STEPS = MAX_LEN - 1


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Engine:
    """One context = one device.  Not thread-safe (same as the reference model object)."""

    def __init__(self, device, max_candidates=1024, attach_cost_mdl=False, decoder_slot_chunk=0, model="tree",
                 lib=None, reserved0=0, hierarchy_levels=8, max_seq_len=200, tied_layers=False):
        """lib / reserved0: test hooks -- tests/verify_lib.py passes the separately built verification library
        (tests/cuda/libgcpb200_verify.so) and its SIMT cross-check switch; the product never sets them."""
        if not torch.cuda.is_available():
            raise _C.GcpB200Error("video_gcp_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.lib = lib if lib is not None else _C.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _C.GcpB200Error("Engine device must be a CUDA device, got %s" % device)
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.max_candidates = int(max_candidates)
        self.attach_cost_mdl = bool(attach_cost_mdl)
        self.model = model
        kind = {"tree": _C.MODEL_TREE, "sequential": _C.MODEL_SEQUENTIAL, "tree_adaptive": _C.MODEL_TREE_ADAPTIVE}[model]
        # tree shape: 25-room = 8 levels / 200 frames / one TreeModule per level; 9-room = 7 / 100 / tied
        self.depth, self.max_len, self.tied = int(hierarchy_levels), int(max_seq_len), bool(tied_layers)
        self.n_nodes = (1 << self.depth) - 1 if model != "sequential" else N_NODES
        self.seq_steps = self.max_len - 1
        cfg = _C.Config(self.index, self.max_candidates, int(attach_cost_mdl), int(reserved0),
                        int(decoder_slot_chunk), kind, self.depth if model != "sequential" else 0, self.max_len, int(self.tied))
        h = C.c_void_p()
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_create(C.byref(h), C.byref(cfg)))
        self.h = h
        self.weights_loaded = False
        self._bufs = {}          # persistent output buffers (no allocator traffic on the hot path)

    def _check(self, rc):
        _C.check(rc, self.lib)

    def _buf(self, name, shape, dtype=torch.float32):
        """Output buffer reused across calls: contents are valid until the next call that writes `name`."""
        key = (name, tuple(shape), dtype)
        t = self._bufs.get(key)
        if t is None:
            for k in [k for k in self._bufs if k[0] == name]:
                del self._bufs[k]
            t = torch.empty(*shape, device=self.device, dtype=dtype)
            self._bufs[key] = t
        return t

    def close(self):
        if getattr(self, "h", None):
            self.lib.gcpb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def load_weights(self, state_dict):
        """state_dict: reference key names -> tensors (any device); packed + uploaded by the library."""
        keep, arr = [], []
        for k, v in state_dict.items():
            if not torch.is_floating_point(v):
                continue
            t = v.detach().to("cpu", torch.float32).contiguous()
            if t.dim() > 4 or t.dim() == 0:
                continue
            keep.append(t)
            shape = (C.c_int64 * 4)(*(list(t.shape) + [0] * (4 - t.dim())))
            arr.append(_C.Tensor(k.encode(), C.c_void_p(t.data_ptr()), t.dim(), shape))
        tens = (_C.Tensor * len(arr))(*arr)
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_load_weights(self.h, tens, len(arr)))
        self.weights_loaded = True

    # ------------------------------------------------------------------------------------------
    def rollout(self, I_0, I_g, z, end_ind=None, seed=0, images_shared=False, want_images=True,
                want_prior=False, want_existence=True, want_aux=True, want_logits=True, fresh=False,
                prune_threshold=0.5, decode_kept_only=False, l2_goal=None, l2_dense=True, l2_final_step_weight=1.0,
                l2_out=None, tree_kept_only=False, sort_sampled_lengths=False):
        """Device tensors in, dict of device tensors out.  z: [B,255,256] fp32, either on the device or a PINNED host
        tensor; a host tensor is uploaded by the library level by level on its own copy stream, overlapped with
        the encoder and the upper tree levels (out["z"] is the device copy, valid in stream order after the call).
        Outputs live in persistent buffers owned by the engine (overwritten by the next rollout) unless
        fresh=True.

        Planner mode: decode_kept_only=True decodes only the nodes balanced pruning keeps (out["images_df"] then holds
        exactly those nodes' images, the other entries keep whatever the buffer held); l2_goal ([3,32,32] in [-1,1]) adds
        out["l2_cost"] ([B], written to `l2_out` if given), the L2 image cost reduced inside the decoder-tail kernel --
        with want_images=False no image is written at all.  tree_kept_only=True (with decode_kept_only, want_existence=False)
        also restricts the tree recursion to the (node, candidate tile) pairs some candidate keeps; sort_sampled_lengths=True
        hands the sampled lengths to the candidates in descending order (same joint distribution; see gcpb200.h), which is
        what makes that restriction effective."""
        dev = self.device
        B = z.shape[0]
        f32 = dict(device=dev, dtype=torch.float32)
        assert z.dtype == torch.float32 and z.is_contiguous() and tuple(z.shape[1:]) == (self.n_nodes, NZ_VAE)
        z_host = None
        if not z.is_cuda:
            if not z.is_pinned():
                z = z.pin_memory()
            z_host, z = z, self._buf("z_dev", (B, self.n_nodes, NZ_VAE))
            self._z_host_ref = z_host          # keep the host buffer alive until the next call
        I_0 = I_0.to(**f32).contiguous()
        I_g = I_g.to(**f32).contiguous()
        if fresh:
            mk = lambda name, shape, dtype=torch.float32: torch.empty(*shape, device=dev, dtype=dtype)
        else:
            mk = self._buf
        out = dict(z=z, e_0=mk("e_0", (B, NZ_ENC)), e_g=mk("e_g", (B, NZ_ENC)), end_ind=mk("end_ind", (B,), torch.int64),
                   e_df=mk("e_df", (B, self.n_nodes, NZ_ENC)))
        if want_logits:
            out["seq_len_logits"] = mk("seq_len_logits", (B, self.max_len))
        if want_prior:
            out["mu_df"] = mk("mu_df", (B, self.n_nodes, NZ_VAE))
            out["log_sigma_df"] = mk("log_sigma_df", (B, self.n_nodes, NZ_VAE))
        if want_images:
            out["images_df"] = mk("images_df", (B, self.n_nodes, 3, 32, 32))
        adaptive = self.model == "tree_adaptive"
        if adaptive:
            # AdaptiveBinding: distance-predictor logits + the kept (depth-first) node list per candidate
            want_existence = want_aux = False
            out["distances"] = mk("distances", (B, self.n_nodes - 1))
            out["pruned_nodes"] = mk("pruned_nodes", (B, self.n_nodes), torch.int32)
            out["pruned_len"] = mk("pruned_len", (B,), torch.int32)
        if want_existence:
            out["existence"] = mk("existence", (B, self.n_nodes))
        if want_aux:
            out["model_enc_seq"] = mk("model_enc_seq", (B, self.max_len, NZ_ENC))
            out["actions"] = mk("actions", (B, self.max_len, 2))
            out["regressed_state"] = mk("regressed_state", (B, self.max_len, 2))
        if end_ind is not None:
            end_ind = end_ind.to(device=dev, dtype=torch.int64).contiguous()
        if l2_goal is not None:
            l2_goal = l2_goal.to(**f32).contiguous()
            assert l2_goal.numel() == 3072
            out["l2_cost"] = self._cost_out(l2_out, B) if l2_out is not None else mk("l2_cost", (B,))
            self._l2_goal_ref = l2_goal
        io = _C.RolloutIO(
            _ptr(I_0), _ptr(I_g), int(images_shared), _ptr(z), _ptr(z_host), _ptr(end_ind), int(seed), int(B),
            _ptr(out["e_0"]), _ptr(out["e_g"]), _ptr(out.get("seq_len_logits")), _ptr(out["end_ind"]),
            _ptr(out["e_df"]), _ptr(out.get("mu_df")), _ptr(out.get("log_sigma_df")), _ptr(out.get("images_df")),
            _ptr(out.get("existence")), _ptr(out.get("model_enc_seq")), _ptr(out.get("actions")),
            _ptr(out.get("regressed_state")), _ptr(out.get("distances")), _ptr(out.get("pruned_nodes")),
            _ptr(out.get("pruned_len")), float(prune_threshold), int(bool(decode_kept_only)), _ptr(l2_goal),
            _ptr(out.get("l2_cost")), int(bool(l2_dense)), float(l2_final_step_weight), int(bool(tree_kept_only)),
            int(bool(sort_sampled_lengths)))
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_rollout(self.h, C.byref(io), _stream()))
        return out

    def seq_rollout(self, I_0, I_g, z, end_ind=None, given_end_ind=None, seed=0, images_shared=False, want_images=True,
                    want_prior=False, want_aux=True, want_logits=True, fresh=False):
        """Sequential GCP rollout.  z: [B,199,256] fp32 device tensor.  Returns dict of device tensors (persistent
        buffers unless fresh=True): e_0, e_g, end_ind, encodings [B,199,128], images [B,200,3,32,32], ..."""
        dev = self.device
        B = z.shape[0]
        f32 = dict(device=dev, dtype=torch.float32)
        z = z.to(**f32).contiguous()
        assert tuple(z.shape[1:]) == (self.seq_steps, NZ_VAE)
        I_0 = I_0.to(**f32).contiguous()
        I_g = I_g.to(**f32).contiguous()
        if fresh:
            mk = lambda name, shape, dtype=torch.float32: torch.empty(*shape, device=dev, dtype=dtype)
        else:
            mk = self._buf
        out = dict(z=z, e_0=mk("e_0", (B, NZ_ENC)), e_g=mk("e_g", (B, NZ_ENC)), end_ind=mk("end_ind", (B,), torch.int64),
                   encodings=mk("encodings", (B, self.seq_steps, NZ_ENC)))
        if want_logits:
            out["seq_len_logits"] = mk("seq_len_logits", (B, self.max_len))
        if want_prior:
            out["mu"] = mk("seq_mu", (B, self.seq_steps, NZ_VAE))
            out["log_sigma"] = mk("seq_log_sigma", (B, self.seq_steps, NZ_VAE))
        if want_images:
            out["images"] = mk("seq_images", (B, self.max_len, 3, 32, 32))
        if want_aux:
            out["model_enc_seq"] = mk("model_enc_seq", (B, self.max_len, NZ_ENC))
            out["actions"] = mk("actions", (B, self.max_len, 2))
            out["regressed_state"] = mk("regressed_state", (B, self.max_len, 2))
        i64 = lambda t: None if t is None else t.to(device=dev, dtype=torch.int64).contiguous()
        end_ind, given_end_ind = i64(end_ind), i64(given_end_ind)
        io = _C.SeqIO(
            _ptr(I_0), _ptr(I_g), int(images_shared), _ptr(z), _ptr(end_ind), _ptr(given_end_ind), int(seed), int(B),
            _ptr(out["e_0"]), _ptr(out["e_g"]), _ptr(out.get("seq_len_logits")), _ptr(out["end_ind"]),
            _ptr(out["encodings"]), _ptr(out.get("mu")), _ptr(out.get("log_sigma")), _ptr(out.get("images")),
            _ptr(out.get("model_enc_seq")), _ptr(out.get("actions")), _ptr(out.get("regressed_state")))
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_seq_rollout(self.h, C.byref(io), _stream()))
        return out

    def _cost_out(self, out, B):
        """Destination of a [B] cost vector: the caller's `out` (contiguous fp32 on this device; e.g. a slice of a larger
        vector when a planner rolls out in chunks) or the persistent buffer the next cost call overwrites."""
        if out is None:
            return self._buf("cost_l2", (B,))
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (B,)
        return out

    def cost_l2_seq(self, images, end_ind, goal, dense=True, final_step_weight=1.0, out=None):
        """L2 image cost over time-ordered image sequences images [B,n_frames,3,32,32] cut at end_ind."""
        B, n_frames = images.shape[:2]
        cost = self._cost_out(out, B)
        goal = goal.to(device=self.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_cost_l2_seq(self.h, _ptr(images), int(n_frames), _ptr(end_ind.contiguous()), _ptr(goal), B,
                                                  int(dense), float(final_step_weight), _ptr(cost), _stream()))
        return cost

    def gather_nodes(self, src_df, nodes, length):
        """src_df [B,255,D...] -> [B,255,D] with rows nodes[c,:length[c]] in order, zeros after (adaptive pruning)."""
        B = src_df.shape[0]
        flat = src_df.reshape(B, self.n_nodes, -1).contiguous()
        D = flat.shape[2]
        dst = torch.empty(B, self.n_nodes, D, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_gather_nodes(self.h, _ptr(flat), _ptr(nodes.contiguous()), _ptr(length.contiguous()), B, D,
                                                   _ptr(dst), _stream()))
        return dst

    def cost_l2_nodes(self, images_df, nodes, length, goal, dense=True, final_step_weight=1.0, out=None):
        B = images_df.shape[0]
        cost = self._cost_out(out, B)
        goal = goal.to(device=self.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_cost_l2_nodes(self.h, _ptr(images_df), _ptr(nodes.contiguous()), _ptr(length.contiguous()),
                                                    _ptr(goal), B, int(dense), float(final_step_weight), _ptr(cost), _stream()))
        return cost

    def prune_gather(self, src_df, end_ind):
        """src_df [B,255,D...] -> [B,200,D] with frames 0..end_ind in order, zeros after."""
        B = src_df.shape[0]
        flat = src_df.reshape(B, self.n_nodes, -1).contiguous()
        D = flat.shape[2]
        dst = torch.empty(B, self.max_len, D, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_prune_gather(self.h, _ptr(flat), _ptr(end_ind.contiguous()), B, D, _ptr(dst), _stream()))
        return dst

    def cost_l2(self, images_df, end_ind, goal, dense=True, final_step_weight=1.0, out=None):
        B = images_df.shape[0]
        cost = self._cost_out(out, B)
        goal = goal.to(device=self.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_cost_l2(self.h, _ptr(images_df), _ptr(end_ind.contiguous()), _ptr(goal), B, int(dense),
                                              float(final_step_weight), _ptr(cost), _stream()))
        return cost

    def cost_learned(self, e_df, end_ind, goal_seq, out=None):
        B = e_df.shape[0]
        cost = torch.empty(B, device=self.device, dtype=torch.float32) if out is None else self._cost_out(out, B)
        goal_seq = goal_seq.to(device=self.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_cost_learned(self.h, _ptr(e_df), _ptr(end_ind.contiguous()), B, _ptr(goal_seq),
                                                   goal_seq.shape[0], _ptr(cost), _stream()))
        return cost

    def cost_pairs(self, lat, idx1, idx2, seg_off=None):
        """Learned pairwise cost of row pairs (idx1[i], idx2[i]) of the latent table lat [R,128]; with seg_off
        (int32 [n_seg+1]) the pair costs are summed per segment.  Index arrays may be host lists / numpy."""
        as_i32 = lambda v: torch.as_tensor(v, dtype=torch.int32).to(self.device).contiguous()
        lat = lat.to(device=self.device, dtype=torch.float32).contiguous()
        idx1, idx2 = as_i32(idx1), as_i32(idx2)
        n = int(idx1.shape[0])
        assert idx2.shape[0] == n and lat.shape[-1] == NZ_ENC
        n_seg = 0
        if seg_off is not None:
            seg_off = as_i32(seg_off)
            n_seg = int(seg_off.shape[0]) - 1
        cost = torch.empty(n_seg if seg_off is not None else n, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_cost_pairs(self.h, _ptr(lat), _ptr(idx1), _ptr(idx2), n, _ptr(seg_off), n_seg,
                                                 _ptr(cost), _stream()))
        return cost

    def infer_action(self, img, target_latent, want_enc=False):
        """Closed-loop step: action = inv_mdl(cat(encoder(img), target_latent)).  img [n,3,32,32] in [-1,1]."""
        f32 = dict(device=self.device, dtype=torch.float32)
        img = img.to(**f32).contiguous()
        target_latent = target_latent.to(**f32).reshape(img.shape[0], NZ_ENC).contiguous()
        n = img.shape[0]
        action = torch.empty(n, 2, **f32)
        enc = torch.empty(n, NZ_ENC, **f32) if want_enc else None
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_infer_action(self.h, _ptr(img), _ptr(target_latent), n, _ptr(action), _ptr(enc),
                                                   _stream()))
        return (action, enc) if want_enc else action

    def topk(self, cost, k):
        N = cost.shape[0]
        idx = torch.empty(k, device=self.device, dtype=torch.int32)
        val = torch.empty(k, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_topk(self.h, _ptr(cost.contiguous()), N, k, _ptr(idx), _ptr(val), _stream()))
        return idx, val

    def refit(self, z, elite_idx):
        mean = torch.empty(self.n_nodes, NZ_VAE, device=self.device, dtype=torch.float32)
        std = torch.empty(self.n_nodes, NZ_VAE, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_refit(self.h, _ptr(z), _ptr(elite_idx.contiguous()), elite_idx.shape[0], _ptr(mean),
                                            _ptr(std), _stream()))
        return mean, std

    def sample_noise(self, n, mean=None, std=None, std_scalar=1.0, seed=0, first_candidate_id=0, clip=float("inf"),
                     out=None):
        z = out if out is not None else torch.empty(n, self.n_nodes, NZ_VAE, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_sample_noise(self.h, _ptr(mean), _ptr(std), float(std_scalar), int(seed),
                                                   int(first_candidate_id), int(n), float(min(clip, 3.0e38)), _ptr(z),
                                                   _stream()))
        return z

    def sample_noise_ids(self, ids, mean=None, std=None, std_scalar=1.0, seed=0, clip=float("inf")):
        """Noise of the global candidate ids in `ids` (int32 cuda tensor)."""
        n = ids.shape[0]
        z = torch.empty(n, self.n_nodes, NZ_VAE, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_sample_noise_ids(self.h, _ptr(mean), _ptr(std), float(std_scalar), int(seed),
                                                       _ptr(ids.contiguous()), int(n), float(min(clip, 3.0e38)), _ptr(z),
                                                       _stream()))
        return z

    # ------------------------------------------------------------------------------------------
    def forward_loss(self, traj_seq, pad_mask, end_ind, states, actions, eps, inv_t0, inv_t1, cost_start, cost_end,
                     cost_target=None, I_0=None, I_g=None, want=("nll_per_frame", "kl_per_seq")):
        """Training-phase forward + loss (gcpb200_forward_loss; reference: `model(inputs)` + `model.loss` +
        `model.get_total_loss` in .train() mode, gcp/prediction/train.py:155-157,204-206).  Device tensors in; returns a
        dict with `losses` ([9] device tensor, order _C.LOSS_NAMES) and the optional outputs named in `want`
        (any field of gcpb200_train_io).  cost_target None: EuclideanPathLength of the ground-truth frames, on the device."""
        dev = self.device
        f32 = dict(device=dev, dtype=torch.float32)
        i64 = dict(device=dev, dtype=torch.int64)
        B, T = traj_seq.shape[:2]
        assert T == self.max_len and tuple(traj_seq.shape[2:]) == (3, 32, 32)
        traj_seq = traj_seq.to(**f32).contiguous()
        end_ind = torch.as_tensor(end_ind).to(**i64).contiguous()
        if I_0 is None:
            I_0 = traj_seq[:, 0]
        if I_g is None:
            I_g = traj_seq[torch.arange(B, device=dev), end_ind]
        ins = dict(traj_seq=traj_seq, pad_mask=pad_mask.to(**f32).contiguous(), end_ind=end_ind,
                   I_0=I_0.to(**f32).contiguous(), I_g=I_g.to(**f32).contiguous(), states=states.to(**f32).contiguous(),
                   actions=actions.to(**f32).contiguous(), eps=eps.to(**f32).contiguous(),
                   inv_t0=torch.as_tensor(inv_t0).to(**i64).contiguous(), inv_t1=torch.as_tensor(inv_t1).to(**i64).contiguous(),
                   cost_start=torch.as_tensor(cost_start).to(**i64).contiguous(),
                   cost_end=torch.as_tensor(cost_end).to(**i64).contiguous(),
                   cost_target=None if cost_target is None else torch.as_tensor(cost_target).to(**f32).reshape(-1).contiguous())
        assert tuple(ins["eps"].shape) == (B, self.n_nodes, NZ_VAE)
        assert ins["cost_target"] is None or ins["cost_target"].numel() == B
        shapes = dict(nll_per_frame=(B, T), kl_per_seq=(B,), e_0=(B, NZ_ENC), e_g=(B, NZ_ENC), enc_traj_seq=(B, T, NZ_ENC),
                      inf_enc_seq=(B, T, NZ_ENC), seq_len_logits=(B, T), e_df=(B, self.n_nodes, NZ_ENC),
                      p_mu=(B, self.n_nodes, NZ_VAE), p_log_sigma=(B, self.n_nodes, NZ_VAE), q_mu=(B, self.n_nodes, NZ_VAE),
                      q_log_sigma=(B, self.n_nodes, NZ_VAE), match_timesteps=(B, self.n_nodes), images_df=(B, self.n_nodes, 3, 32, 32),
                      existence=(B, self.n_nodes), model_enc_seq=(B, T, NZ_ENC), regressed_state=(B, T, 2), inv_actions=(B, 2),
                      cost_pred=(B,))
        out = dict(losses=self._buf("train_losses", (len(_C.LOSS_NAMES),)))
        for name in want:
            out[name] = self._buf("train_" + name, shapes[name], torch.int32 if name == "match_timesteps" else torch.float32)
        io = _C.TrainIO()
        for k, v in ins.items():
            setattr(io, k, _ptr(v))
        io.B = B
        for k, v in out.items():
            setattr(io, k, _ptr(v))
        self._train_refs = ins            # keep inputs alive until the (stream-ordered) call has run
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_forward_loss(self.h, C.byref(io), _stream()))
        return out

    # ------------------------------------------------------------------------------------------
    # DTW family.  Device tensors in / out; outputs are persistent buffers (valid until the next call).
    def cdist_mean(self, x, y):
        """batch_cdist(x, y, reduction='mean') (blox/torch/ops.py:62-91).  x [B,n,...], y [B,m,...] -> [B,n,m]."""
        x = x.to(self.device, torch.float32).flatten(2).contiguous()
        y = y.to(self.device, torch.float32).flatten(2).contiguous()
        B, n, dim = x.shape
        assert y.shape[0] == B and y.shape[2] == dim, (x.shape, y.shape)
        out = self._buf("cdist", (B, n, y.shape[1]))
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_cdist_mean(self.h, _ptr(x), _ptr(y), B, n, y.shape[1], dim, _ptr(out), _stream()))
        self._dtw_refs = (x, y)
        return out

    def soft_dtw(self, cost, temp=1.0, end_inds=None, want_bf=False, want_tables=False):
        """soft_dtw(cost / temp, end_inds) (probabilistic_dtw.py:82-121).  Returns dict(w [B,r,c], rowsum_max [1],
        optionally w_bf = depthfirst2breadthfirst(normalize(w, 1)) and the float64 forward / backward tables)."""
        cost = cost.to(self.device, torch.float32).contiguous()
        B, r, c = cost.shape
        ei = None if end_inds is None else torch.as_tensor(end_inds).to(self.device, torch.int64).contiguous()
        assert ei is None or tuple(ei.shape) == (B,)
        ws = self._buf("soft_dtw_ws", (2, B, r, c), torch.float64)
        assert ws.numel() * 8 == self.lib.gcpb200_soft_dtw_workspace(B, r, c)
        out = dict(w=self._buf("soft_dtw_w", (B, r, c)), rowsum_max=self._buf("soft_dtw_rowsum", (1,)))
        if want_bf:
            out["w_bf"] = self._buf("soft_dtw_w_bf", (B, r, c))
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_soft_dtw(self.h, _ptr(cost), float(temp), _ptr(ei), B, r, c, _ptr(ws), _ptr(out["w"]),
                                               _ptr(out.get("w_bf")), _ptr(out["rowsum_max"]), _stream()))
        self._dtw_refs = (cost, ei)
        if want_tables:
            out["forward"], out["backward"] = ws[0], ws[1]
        return out

    def dtw(self, cost, end_ind=None, want_matches=True):
        """c_dtw / batched_dtw (gcp/evaluation/dtw_utils.py:77-130,201-241).  cost [B,r,c] fp32 or float64.  Returns
        dict(acc [B,r+1,c+1] f64 padded table, dist [B], path_p / path_q [B,r+c-1] right-aligned, path_len [B],
        match_inds [B,c])."""
        f64 = cost.dtype == torch.float64
        cost = cost.to(self.device, torch.float64 if f64 else torch.float32).contiguous()
        B, r, c = cost.shape
        ei = None if end_ind is None else torch.as_tensor(end_ind).to(self.device, torch.int64).contiguous()
        i32 = torch.int32
        out = dict(acc=self._buf("dtw_acc", (B, r + 1, c + 1), torch.float64), dist=self._buf("dtw_dist", (B,), torch.float64),
                   path_p=self._buf("dtw_p", (B, r + c - 1), i32), path_q=self._buf("dtw_q", (B, r + c - 1), i32),
                   path_len=self._buf("dtw_len", (B,), i32))
        if want_matches:
            out["match_inds"] = self._buf("dtw_match", (B, c), i32)
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_dtw(self.h, _ptr(cost), int(f64), _ptr(ei), B, r, c, _ptr(out["acc"]), _ptr(out["dist"]),
                                          _ptr(out["path_p"]), _ptr(out["path_q"]), _ptr(out["path_len"]),
                                          _ptr(out.get("match_inds")), _stream()))
        self._dtw_refs = (cost, ei)
        return out

    def gather_rows(self, src, idx):
        """src[idx] for a device int32 index vector (rows of a multiple of 4 floats)."""
        src = src.to(self.device, torch.float32).contiguous()
        idx = idx.to(self.device, torch.int32).contiguous()
        row = src[0].numel()
        out = torch.empty((idx.shape[0],) + tuple(src.shape[1:]), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_gather_rows(self.h, _ptr(src), _ptr(idx), idx.shape[0], row, _ptr(out), _stream()))
        return out

    def launch_count(self):
        return int(self.lib.gcpb200_launch_count(self.h))

    PHASES = ("encoder_length", "tree_recursion", "decoder_gemm", "decoder_tail", "heads", "rollout_total")

    def profile_enable(self, on=True):
        self._check(self.lib.gcpb200_profile_enable(self.h, int(on)))

    def profile_read(self):
        """dict phase -> ms accumulated since the last read, plus decoder-tail image / launch counts."""
        ms = (C.c_double * 6)()
        imgs, launches = C.c_int64(), C.c_int64()
        with torch.cuda.device(self.index):
            self._check(self.lib.gcpb200_profile_read(self.h, ms, C.byref(imgs), C.byref(launches)))
        out = {k: ms[i] for i, k in enumerate(self.PHASES)}
        out["tail_images"], out["tail_launches"] = imgs.value, launches.value
        return out
