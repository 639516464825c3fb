"""CEM cost functions (reference: gcp/planning/cem/cost_fcn.py).  Each accepts either the reference's
list-of-numpy rollouts or a `DeviceRollouts`; both routes run the same device kernels."""
import numpy as np
import torch

from .cem_simulator import DeviceRollouts


def _lists_to_device(engine, seqs, width):
    """Pack per-candidate [L_i, width] arrays into [B,max_seq_len,width] + end_ind on the device."""
    B = len(seqs)
    buf = np.zeros((B, engine.max_len, width), dtype=np.float32)
    end = np.zeros((B,), dtype=np.int64)
    for i, s in enumerate(seqs):
        s = np.asarray(s, dtype=np.float32).reshape(len(s), -1)
        buf[i, :len(s)] = s
        end[i] = len(s) - 1
    return torch.as_tensor(buf).to(engine.device), torch.as_tensor(end).to(engine.device)


class CostFcn:
    """Base class (cost_fcn.py:9-22): dense sum over steps or last step only, final step weighted."""

    def __init__(self, dense_cost, final_step_weight=1.0, *unused_args):
        self._dense_cost = dense_cost
        self._final_step_weight = final_step_weight


class ImageCost:
    LATENT_SIZE = 128

    @property
    def input_dim(self):
        return self.LATENT_SIZE

    def _split_state_rollout(self, rollouts):
        from ..types import AttrDict
        imgs, lats = [], []
        for r in rollouts:
            flat = r[..., :-self.input_dim]
            res = int(np.sqrt(flat.shape[1] / 3))
            imgs.append(flat.reshape(flat.shape[0], 3, res, res))
            lats.append(r[..., -self.input_dim:])
        return AttrDict(image_rollout=imgs, latent_rollout=lats)


class L2ImageCost(CostFcn, ImageCost):
    """L2 distance to the goal image per frame (cost_fcn.py:65-76)."""

    def __init__(self, dense_cost, final_step_weight=1.0, *unused_args, engine=None):
        super().__init__(dense_cost, final_step_weight)
        self.engine = engine

    def fused_spec(self):
        """(dense_cost, final_step_weight): what the rollout needs to reduce this cost inside the decoder-tail kernel."""
        return (bool(self._dense_cost), float(self._final_step_weight))

    def device_cost(self, ro, out=None):
        """[B] costs of a DeviceRollouts; `out`: caller-owned destination (else an engine buffer that the next cost
        call overwrites).  A rollout made in planner mode already carries the cost (same bits as the kernel below gives on
        its frames up to summation order, identical between the two decode modes)."""
        if getattr(ro, "l2_cost", None) is not None and ro.l2_spec == self.fused_spec():
            if out is None:
                return ro.l2_cost
            if out.data_ptr() != ro.l2_cost.data_ptr():
                out.copy_(ro.l2_cost)
            return out
        if ro.sequential:
            return ro.model.engine.cost_l2_seq(ro.images_seq, ro.end_ind, ro.goal_chw, self._dense_cost,
                                               self._final_step_weight, out=out)
        return ro.model.engine.cost_l2(ro.images_df, ro.end_ind, ro.goal_chw, self._dense_cost, self._final_step_weight,
                                       out=out)

    def __call__(self, cem_outputs, goal):
        if isinstance(cem_outputs, DeviceRollouts):
            return self.device_cost(cem_outputs).cpu().numpy()
        if self.engine is None:
            raise RuntimeError("L2ImageCost on host lists needs `engine=` (there is no CPU implementation)")
        imgs = [np.asarray(r)[:, :3072] for r in cem_outputs]
        buf, end = _lists_to_device(self.engine, imgs, 3072)
        # frames are already in order: present them as a depth-first array whose frame t sits at node t
        images_df, node_end = _as_df(self.engine, buf, end)
        goal_chw = torch.as_tensor(np.asarray(goal, dtype=np.float32)).to(self.engine.device)[0].permute(2, 0, 1) * 2 - 1.0
        return self.engine.cost_l2(images_df, node_end, goal_chw.contiguous(), self._dense_cost,
                                   self._final_step_weight).cpu().numpy()


def _as_df(engine, seq, end):
    """Scatter ordered frames [B,200,D] into a depth-first [B,255,D] array at the balanced-pruning
    positions, so the device cost kernels (which gather through the pruning map) see them unchanged."""
    B, _, D = seq.shape
    df = torch.zeros(B, engine.n_nodes, D, device=seq.device)
    from ..pruning import frame_nodes
    for i, e in enumerate(end.tolist()):
        nodes = torch.as_tensor(frame_nodes(e, engine.depth), device=seq.device)
        df[i, nodes] = seq[i, :e + 1]
    return df, end


class LearnedCostEstimate:
    """Learned pairwise cost summed along the latent sequence (cost_fcn.py:79-101).  The network is the
    model's `cost_mdl.cost_pred` (TestTimeCostModel loads the same weights from the checkpoint)."""

    def __init__(self, config=None, model=None):
        self._model = model if model is not None else (config or {}).get("model")
        if self._model is None:
            raise ValueError("LearnedCostEstimate needs the TreeModel that owns cost_mdl weights (`model=`)")

    @property
    def input_dim(self):
        return 128

    @property
    def engine(self):
        return self._model.engine

    def device_cost(self, ro, goal_seq, out=None):
        return ro.model.engine.cost_learned(ro.e_df, ro.end_ind, goal_seq, out=out)

    def pairs_device(self, lat, idx1, idx2, seg_off=None):
        """Pair costs over rows of a device latent table (per pair, or summed per segment): what the hierarchical
        optimiser evaluates (tree_optimizer.py:96-99,145-150).  Returns a device tensor."""
        return self._model.engine.cost_pairs(lat, idx1, idx2, seg_off)

    def __call__(self, start_enc, goal_enc):
        eng = self._model.engine
        if isinstance(start_enc, np.ndarray):
            # cost of single start / goal pairs (cost_fcn.py:84-87): [n,128] x 2 -> [n,1]
            a = np.asarray(start_enc, dtype=np.float32).reshape(-1, 128)
            b = np.asarray(goal_enc, dtype=np.float32).reshape(-1, 128)
            n = a.shape[0]
            lat = torch.as_tensor(np.concatenate([a, b])).to(eng.device)
            return eng.cost_pairs(lat, np.arange(n), n + np.arange(n)).cpu().numpy()[:, None]
        if isinstance(start_enc, list):
            out = np.zeros(len(start_enc), dtype=np.float32)
            buf, end = _lists_to_device(eng, start_enc, 128)
            df, end = _as_df(eng, buf, end)
            # group candidates by identical goal sequence (one kernel call per distinct goal)
            groups = {}
            for i, g in enumerate(goal_enc):
                groups.setdefault(id(g), (g, []))[1].append(i)
            for g, idx in groups.values():
                it = torch.as_tensor(idx, device=eng.device)
                gs = torch.as_tensor(np.asarray(g, dtype=np.float32).reshape(len(g), -1)).to(eng.device)
                out[idx] = eng.cost_learned(df[it].contiguous(), end[it].contiguous(), gs).cpu().numpy()
            return out
        raise ValueError("Dimensionality of input to learned cost function not supported!")


class ImageLearnedCostEstimate(LearnedCostEstimate, ImageCost):
    pass


class ImageWrappedLearnedCostFcn(LearnedCostEstimate, ImageCost):
    """Unpacks image+latent rollouts; every candidate's goal is the LAST candidate's latent rollout
    (the reference's own HACK, cost_fcn.py:108-116)."""

    chunkable = False       # the "goal" is the last candidate of the WHOLE batch: one rollout call per CEM iteration

    def device_cost(self, ro, goal_seq=None, out=None):
        last = len(ro) - 1
        L = int(ro.end_ind[last]) + 1
        goal = ro.model.engine.prune_gather(ro.e_df[last:last + 1], ro.end_ind[last:last + 1])[0, :L]
        return ro.model.engine.cost_learned(ro.e_df, ro.end_ind, goal, out=out)

    def __call__(self, start_enc, goal_enc=None):
        if isinstance(start_enc, DeviceRollouts):
            return self.device_cost(start_enc).cpu().numpy()
        lat = self._split_state_rollout(start_enc).latent_rollout
        goal = [lat[-1] for _ in range(len(lat))]
        return super().__call__(lat, goal)
