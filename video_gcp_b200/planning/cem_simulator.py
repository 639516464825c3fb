"""Simulator interface of the CEM planner (reference: gcp/planning/cem/cem_simulator.py:7-96).

`rollout()` keeps the reference contract (numpy in, AttrDict of per-candidate numpy lists out).
`rollout_device()` is the B200-first variant the planner uses: everything stays in HBM and only what the
caller asks for is copied back.
"""
import numpy as np
import torch

from ..types import AttrDict


class DeviceRollouts:
    """Result of one batched rollout, resident on the device."""

    def __init__(self, model, inputs, outputs, goal_chw):
        self.model, self.inputs, self.outputs, self.goal_chw = model, inputs, outputs, goal_chw
        self.end_ind = torch.max(outputs.end_ind, torch.ones_like(outputs.end_ind))   # cem_simulator.py:31
        self.z = outputs.z_device          # device copy of the noise that was rolled out
        self.l2_cost = outputs.get("l2_cost", None)     # planner mode: L2 image cost reduced inside the decoder
        self.l2_spec = outputs.get("l2_spec", None)     # (dense_cost, final_step_weight) it was computed with
        self.sequential = "tree" not in outputs
        if self.sequential:
            # sequential model: frames / latents are stored in time order (frame 0 = start image, latent 0 = e_0)
            self.images_seq = outputs.dense_rec.get("images")
            self.enc_seq = outputs.dense_rec.encodings[..., 0, 0]
            return
        self.images_df = outputs.tree.df.images if "images" in outputs.tree._fields else None
        self.has_heads = "actions" in outputs      # False: a cost-only rollout, to_host() is not available
        self.e_df = outputs.tree.df.e_g_prime[..., 0, 0]

    def __len__(self):
        return int(self.end_ind.shape[0])

    def to_host(self, append_latent, idx=None):
        """AttrDict(predictions, actions, states, latents) of numpy lists, for candidates `idx` (default all).
        One device-side gather per field, one D2H copy per field into pinned host memory (a pageable destination costs
        ~10x the copy time at 250 MB of elite frames); above 16 MB the per-candidate arrays are views of those host blocks,
        which stay alive as long as any of the arrays does; a smaller result (the plan) is copied out of them."""
        if not getattr(self, "has_heads", True):
            raise RuntimeError("this rollout was made cost-only (planner_mode heads=False): it carries no frames / actions / states")
        eng = self.model.engine
        ends = self.end_ind.tolist()
        n_all = len(ends)
        sel = list(range(n_all)) if idx is None else [int(i) for i in idx]
        take = (lambda t: t) if idx is None else (lambda t, i=torch.as_tensor(sel, device=self.end_ind.device): t[i])
        end_sel = take(self.end_ind)
        if self.sequential:
            # cem_simulator.py:45-58 with SequentialRecModule.get_sample_with_len (sequential.py:78-94); `latents`
            # is inputs.model_enc_seq capped to the predicted length (cem_simulator.py:41)
            img = take(self.images_seq).reshape(len(sel), self.images_seq.shape[1], -1)
            lat = torch.cat([take(self.inputs.e_0)[:, None, :, 0, 0], take(self.enc_seq)], 1)
        else:
            img = eng.prune_gather(take(self.images_df), end_sel)
            lat = eng.prune_gather(take(self.e_df), end_sel)
        if append_latent:
            img = torch.cat([img, lat], -1)
        # the reference pads these two to the longest sequence of the batch (pad_sequence, base_gcp.py:242); every list entry
        # below is cut to its own length anyway, so the full-length device buffers are copied as they are: the pinned
        # staging shapes then never change between calls (a new shape is a fresh cudaHostAlloc, milliseconds per plan)
        act = take(self.outputs.actions)
        sta = take(self.outputs.regressed_state)

        def pinned(t):
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
            return h
        img, lat, act, sta = pinned(img), pinned(lat), pinned(act), pinned(sta)
        torch.cuda.current_stream(self.end_ind.device).synchronize()
        img, lat, act, sta = img.numpy(), lat.numpy(), act.numpy(), sta.numpy()
        if img.nbytes <= (16 << 20):
            # a plan-sized result leaves the staging blocks at once (0.3 ms of host memcpy): arrays that are views of pinned
            # blocks keep them out of the host allocator's cache for as long as the caller holds the plan, and every
            # block that is missing when the next plan arrives is a cudaHostAlloc (2-45 ms measured)
            img, lat, act, sta = img.copy(), lat.copy(), act.copy(), sta.copy()
        out = AttrDict(predictions=[], actions=[], states=[], latents=[])
        lmax = self.outputs["_lmax"]()  # the reference's padded length: its action tensor has lmax - 1 steps
        for n, i in enumerate(sel):
            L = ends[i] + 1
            out.predictions.append(img[n, :L])
            out.actions.append(act[n, :min(L, lmax - 1)])
            out.states.append(sta[n, :min(L, lmax)])
            out.latents.append(lat[n, :L])
        return out


class GCPSimulator:
    """Implements the simulator interface for GCP models."""

    def __init__(self, model, append_latent):
        self._model = model
        self._append_latent = append_latent
        self._logs = []

    def _postprocess_inputs(self, input_dict):
        return input_dict

    def rollout_device(self, state, goal_state, samples, rollout_len, planner_mode=None):
        """samples: numpy [B,255,256] ([B,199,256] for the sequential model) or a CUDA tensor (kept on device).
        planner_mode (tree model): dict(kept_only=True, images=True/False, l2=(dense_cost, final_step_weight) or None,
        l2_out=[B] destination or None) -- decode only the frames the planner reads and fold the L2 image cost into the
        decoder (see gcpb200_rollout_io.decode_kept_only).  None = the reference behaviour, every node decoded.
        Returns DeviceRollouts."""
        dev = self._model.engine.device
        B = samples.shape[0]
        if isinstance(samples, torch.Tensor):
            # CUDA tensors stay where they are; pinned fp32 host tensors are uploaded by the library, overlapped
            z = samples if (samples.is_cuda or (samples.dtype == torch.float32 and samples.is_pinned())) \
                else samples.to(device=dev, dtype=torch.float32)
        else:
            z = torch.as_tensor(np.ascontiguousarray(samples, dtype=np.float32)).pin_memory()
        # start / goal are converted on the host (12 KB each) and uploaded once: every candidate shares them
        input_dict = AttrDict(
            I_0=torch.as_tensor(np.asarray(state), dtype=torch.float32),
            I_g=torch.as_tensor(np.asarray(goal_state), dtype=torch.float32),
            start_ind=torch.zeros(B, dtype=torch.long, device=dev),
            end_ind=torch.full((B,), rollout_len - 1, dtype=torch.long, device=dev),
            z=z, images_shared=True)
        if planner_mode is not None:
            input_dict.planner_mode = planner_mode
        input_dict = self._postprocess_inputs(input_dict)
        input_dict.I_0 = input_dict.I_0.to(dev, non_blocking=True)
        input_dict.I_g = input_dict.I_g.to(dev, non_blocking=True)
        # device-resident path: nothing on the host needs the batch's longest length, so the model skips that sync
        defer = getattr(self._model, "defer_length_sync", False)
        self._model.defer_length_sync = True
        try:
            with self._model.val_mode():
                out = self._model(input_dict)
        finally:
            self._model.defer_length_sync = defer
        return DeviceRollouts(self._model, input_dict, out, input_dict.I_g[0])

    def rollout(self, state, goal_state, samples, rollout_len, prune=False):
        """Reference contract: AttrDict of python lists with one numpy array per candidate."""
        return self.rollout_device(state, goal_state, samples, rollout_len).to_host(self._append_latent)

    def dump_logs(self, dump_file='rollout_dump.pkl'):
        self._logs = []


class GCPImageSimulator(GCPSimulator):
    def _postprocess_inputs(self, input_dict):
        input_dict = super()._postprocess_inputs(input_dict)
        if input_dict.z.dim() == 3:
            input_dict.z = input_dict.z[..., None, None]
        input_dict.I_0 = self._env2planner(input_dict.I_0)
        input_dict.I_g = self._env2planner(input_dict.I_g)
        return input_dict

    @staticmethod
    def _env2planner(img):
        """[0..1] or [0..255] HWC environment images -> [-1..1] CHW planner images (cem_simulator.py:87-96)."""
        if img.max() > 1.0:
            img = img / 255.0
        if len(img.shape) == 5:
            img = img[0]
        if len(img.shape) == 4:
            img = img.permute(0, 3, 1, 2)
        return (img * 2 - 1.0).contiguous()
