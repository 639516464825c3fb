"""Hierarchical latent optimiser for GCP-tree planning, device-resident
(reference: gcp/planning/tree_optimizer.py:7-203, image variant :164-190).

Same algorithm and the same public surface (`sample()`, `optimize(rollouts, goal)`, `fully_optimized`) as the
reference's ImageHierarchicalTreeLatentOptimizer: the tree's latents are optimised one layer at a time -- best-of-N
subgoal by learned pairwise cost to both parents (:88-121), then recursion into both halves (:123-152), and a dense
best-of-N over the remaining "segments" in the last layer (:82-86,145-150).

B200-first representation.  The reference slices numpy copies of every rollout ([L, 3072+128] per candidate) and calls
the cost network once per (sub)sequence.  Here a batch of rollouts is ONE latent table in HBM (`FrameTable`: row id =
frame), a rollout is an int array of row ids, slicing is index arithmetic, and every cost evaluation of one optimiser
node is a single `gcpb200_cost_pairs` launch chain over (row, row) pairs; only [n] float costs come back.  Images are
gathered for the chosen plan only.  The reference's edge cases are kept exactly: too-short rollouts are replaced by
the inf / 0 / inf dummy (:158-162) whose costs are NaN, NaN wins np.argmin (first one), and the too-short rollouts
compete through the dense segment cost (:147-150).

Sampling: `rng="numpy"` (default) draws with np.random.normal in exactly the reference's order -- including the
draws it throws away -- so a seeded run proposes bit-identical latents; `rng="device"` draws only what is used from the
engine's Philox stream and keeps the latents in HBM.
"""
import numpy as np
import torch

from ..pruning import frame_nodes

INF_ROW = -1            # frame of a dummy sequence filled with inf (cost NaN)
CHOICE_HOOK = None      # tests only: callable(costs, default_choice) -> choice, to replay a fixed decision trace


def _choose(costs):
    k = int(np.argmin(costs))          # NaN counts as the minimum and the first one wins, as in the reference
    return k if CHOICE_HOOK is None else int(CHOICE_HOOK(np.asarray(costs, dtype=np.float64).reshape(-1), k))


class FrameTable:
    """All frames of one batch of rollouts: latent table on the device + lazy image access.
    rollouts[i] is the int64 array of table rows of candidate i's frames, in time order."""

    def __init__(self, engine, lat, rollouts, image_fn):
        self.engine = engine
        self.zero_row = int(lat.shape[0])
        self.lat = torch.cat([lat, torch.zeros(1, lat.shape[1], device=lat.device, dtype=lat.dtype)])
        self.rollouts = rollouts
        self._image_fn = image_fn

    @classmethod
    def from_device(cls, ro):
        """ro: DeviceRollouts of the tree model (frames = balanced-pruned nodes, cem_simulator.py:45-58)."""
        seq = ro.outputs["_pruned_e_g_prime"]                       # [B,200,128] pruned latents, zero padded
        B, T = int(seq.shape[0]), int(seq.shape[1])
        ends = ro.end_ind.tolist()
        rollouts = [np.arange(c * T, c * T + e + 1, dtype=np.int64) for c, e in enumerate(ends)]

        def image_fn(rows):
            c = torch.as_tensor([r // T for r in rows], device=seq.device)
            node = torch.as_tensor([frame_nodes(ends[r // T])[r % T] for r in rows], device=seq.device)
            return ro.images_df[c, node].cpu().numpy()

        return cls(ro.model.engine, seq.reshape(B * T, -1), rollouts, image_fn)

    @classmethod
    def from_host(cls, engine, rollouts, latent_dim):
        """rollouts: list of numpy [L_i, 3072 + latent_dim] as GCPSimulator.rollout returns them (append_latent)."""
        lens = [int(r.shape[0]) for r in rollouts]
        offs = np.concatenate([[0], np.cumsum(lens)])
        lat = torch.as_tensor(np.concatenate([np.asarray(r)[:, -latent_dim:] for r in rollouts]).astype(np.float32))
        rows = [np.arange(offs[i], offs[i + 1], dtype=np.int64) for i in range(len(lens))]
        flat = np.concatenate([np.asarray(r)[:, :-latent_dim] for r in rollouts])

        def image_fn(ids):
            x = flat[np.asarray(ids, dtype=np.int64)]
            res = int(np.sqrt(x.shape[1] / 3))
            return x.reshape(len(ids), 3, res, res)

        return cls(engine, lat.to(engine.device), rows, image_fn)

    def images(self, rows):
        """numpy [n,3,H,W] of the given rows (dummy rows: inf / zeros)."""
        rows = [int(r) for r in rows]
        real = [r for r in rows if r != INF_ROW and r != self.zero_row]
        got = self._image_fn(real) if real else None
        shape = got.shape[1:] if got is not None else (3, 32, 32)
        out = np.zeros((len(rows),) + tuple(shape), dtype=np.float32)
        j = 0
        for i, r in enumerate(rows):
            if r == INF_ROW:
                out[i] = np.inf
            elif r != self.zero_row:
                out[i] = got[j]
                j += 1
        return out


class ImageHierarchicalTreeLatentOptimizer:
    """Optimises latent distributions for GCP-tree layers recursively, one layer at a time."""

    def __init__(self, latent_dim, sampling_rates, depth, subgoal_cost_fcn, ll_cost_fcn, final_layer_samples,
                 engine=None, rng="numpy", seed=0, _counter=None):
        self._latent_dim = latent_dim
        self._depth = depth
        self._subgoal_cost_fcn = subgoal_cost_fcn
        self._ll_cost_fcn = ll_cost_fcn
        for f in (subgoal_cost_fcn, ll_cost_fcn):
            if not hasattr(f, "pairs_device"):
                raise NotImplementedError("the device optimiser needs learned cost functions (LearnedCostEstimate family)")
        self._is_optimized = False
        self._opt_z = None
        self._latest_z_samples = None
        self._engine, self._rng, self._seed = engine, rng, int(seed)
        self._counter = _counter if _counter is not None else [0, None, 0]   # device rng, shared by the whole tree:
        # [next Philox candidate id, pool of latent rows drawn for the current sample() call, rows handed out]
        sampling_rates = list(sampling_rates)
        if sampling_rates:
            self._n_samples = sampling_rates.pop(0)
            self._n_latents = 1
            mk = lambda: type(self)(latent_dim, sampling_rates.copy(), depth - 1, subgoal_cost_fcn, ll_cost_fcn,
                                    final_layer_samples, engine=engine, rng=rng, seed=seed, _counter=self._counter)
            self._children = [[mk() for _ in range(self._n_samples)] for _ in range(2)]
        else:
            self._n_samples = final_layer_samples
            self._n_latents = 2 ** depth - 1
            self._children = None
        self.mean, self.std = 0.0, 1.0          # N(0, 1) per latent element; never refit (sampler.py:113-115)

    # ------------------------------------------------------------------ sampling (tree_optimizer.py:46-72,142-143)
    def _sample(self, n):
        if self._rng == "numpy":
            return np.random.normal(loc=self.mean, scale=self.std, size=(self._n_samples, self._n_latents, self._latent_dim))[:n]
        rows = n * self._n_latents
        pool = self._counter[1]
        z = pool[self._counter[2]:self._counter[2] + rows]
        self._counter[2] += rows
        return z.reshape(n, self._n_latents, self._latent_dim)

    def _rows_needed(self, below):
        """Latent rows one sample() call draws (device rng: drawn as ONE Philox launch, then sliced)."""
        n = 0 if self._is_optimized else (1 if below else self._n_samples) * self._n_latents
        if self._children is not None:
            k = 1 if (self._is_optimized or below) else self._n_samples
            nb = below or not self._is_optimized
            n += sum(self._children[0][i]._rows_needed(nb) + self._children[1][i]._rows_needed(nb) for i in range(k))
        return n

    def sample(self, below_opt_layer=False, _top=True):
        """Latents of all layers, concatenated in depth-first node order: N for the layer being optimised, one for the
        layers above (their optimum) and below (not used for the decision)."""
        if _top and self._rng != "numpy":
            rows = self._rows_needed(below_opt_layer)
            blocks = max(-(-rows // 255), 1)
            pool = self._engine.sample_noise(blocks, None, None, 1.0, self._seed, self._counter[0])
            self._counter[0] += blocks
            self._counter[1:] = [pool.reshape(-1, self._latent_dim), 0]
        if self._is_optimized:
            z = self._opt_z[None]
        else:
            z = self._sample(1 if below_opt_layer else self._n_samples)
            self._latest_z_samples = z
        next_below = below_opt_layer or not self._is_optimized
        if self._children is None:
            return z.copy() if isinstance(z, np.ndarray) else z
        dev = not isinstance(z, np.ndarray)
        samples = []
        for child_left, child_right, z_i in zip(self._children[0], self._children[1], z):
            z_left, z_right = child_left.sample(next_below, False), child_right.sample(next_below, False)
            assert z_left.shape == z_right.shape          # latent tree needs to be balanced
            n = z_left.shape[0]
            mid = z_i[0].expand(n, 1, -1) if dev else np.tile(z_i[0], (n, 1, 1))
            samples.append(torch.cat([z_left, mid, z_right], 1) if dev else np.concatenate([z_left, mid, z_right], 1))
        return torch.cat(samples) if dev else np.concatenate(samples)

    # ------------------------------------------------------------------ optimisation
    def optimize(self, all_rollouts, goal):
        """all_rollouts: DeviceRollouts, or the reference's list of numpy [L_i, 3072+128]; goal: the goal image
        [1,H,W,3] in [0,1].  Returns (best plan images [n,3,H,W], cost)."""
        if isinstance(all_rollouts, (list, tuple)):
            tab = FrameTable.from_host(self._engine or self._subgoal_cost_fcn.engine, list(all_rollouts),
                                       self._subgoal_cost_fcn.input_dim)
        else:
            tab = FrameTable.from_device(all_rollouts)
        plan, cost = self._optimize(tab, tab.rollouts, None)
        imgs = tab.images([r for r in plan if r is not None])
        goal_chw = np.asarray(goal, dtype=np.float32)[0].transpose(2, 0, 1)
        out, j = [], 0
        for r in plan:
            if r is None:
                out.append(goal_chw)             # the raw goal image, as the reference appends it (:109-113)
            else:
                out.append(imgs[j])
                j += 1
        return np.stack(out), cost

    def _optimize(self, tab, rollouts, goal_row):
        """goal_row None: the (image) goal of the whole plan; else the table row of the parent's subgoal frame."""
        if self._children is None:
            return self._optimize_segment(tab, rollouts, goal_row)
        if not self._is_optimized:
            return self._optimize_subgoal(tab, rollouts, goal_row)
        return self._recurse_optimization(tab, rollouts, goal_row)

    def _pair_costs(self, tab, a, b):
        """cost_fcn(lat[a], lat[b]) per pair; NaN where a dummy inf frame is involved (tree_optimizer.py:96-99)."""
        a, b = np.asarray(a, dtype=np.int64), np.asarray(b, dtype=np.int64)
        bad = (a == INF_ROW) | (b == INF_ROW)
        out = np.full(len(a), np.nan, dtype=np.float32)
        if (~bad).any():
            out[~bad] = self._subgoal_cost_fcn.pairs_device(tab.lat, a[~bad], b[~bad]).cpu().numpy()
        return out

    def _segment_costs(self, tab, rollouts, goal_row):
        """ll_cost_fcn on cat(latents, goal latent) per rollout (:145-150,167-174): one launch chain for all pairs."""
        i1, i2, off, bad = [], [], [0], []
        for r in rollouts:
            g = int(r[-1]) if goal_row is None else int(goal_row)
            seq = np.concatenate([r, [g]])
            dummy = bool((seq == INF_ROW).any())
            bad.append(dummy)
            if not dummy:
                i1.append(seq[:-1])
                i2.append(seq[1:])
            off.append(off[-1] + (0 if dummy else len(seq) - 1))
        out = np.full(len(rollouts), np.nan, dtype=np.float32)
        if i1:
            cost = self._ll_cost_fcn.pairs_device(tab.lat, np.concatenate(i1), np.concatenate(i2), seg_off=np.asarray(off))
            cost = cost.cpu().numpy()
            for k, dummy in enumerate(bad):
                if not dummy:
                    out[k] = cost[k]
        return out

    def _best_of_n_segments(self, tab, rollouts, goal_row):
        cost = self._segment_costs(tab, rollouts, goal_row)
        k = _choose(cost)
        return list(rollouts[k]), cost[k], k

    def _optimize_segment(self, tab, rollouts, goal_row):
        best, cost, k = self._best_of_n_segments(tab, rollouts, goal_row)
        self._opt_z = self._latest_z_samples[k]
        self._is_optimized = True
        return best, cost

    def _optimize_subgoal(self, tab, rollouts, goal_row):
        n = len(rollouts)
        start = np.array([r[0] for r in rollouts], dtype=np.int64)
        sub = np.array([r[len(r) // 2] for r in rollouts], dtype=np.int64)
        goal = np.array([r[-1] if goal_row is None else goal_row for r in rollouts], dtype=np.int64)
        c = self._pair_costs(tab, np.concatenate([start, sub]), np.concatenate([sub, goal]))
        total = c[:n] + c[n:]
        k = _choose(total)
        self._opt_z = self._latest_z_samples[k]
        plan = [int(start[k])]
        if len(rollouts[k]) // 2 != 0:                 # subgoal == start only if the sequence is too short (:104-105)
            plan.append(int(sub[k]))
        if goal_row is None:
            plan.append(None)                          # the very final goal is appended once (:106-111)
        self._children = [c_[:1] for c_ in self._children]
        self._n_samples = 1
        self._is_optimized = True
        return plan, total[k:k + 1]

    def _recurse_optimization(self, tab, rollouts, goal_row):
        n_all = len(rollouts)
        assert n_all % self._n_samples == 0
        per = n_all // self._n_samples
        best_costs, best_plans = [], []
        for i, (child_left, child_right) in enumerate(zip(self._children[0], self._children[1])):
            group = list(rollouts[i * per:(i + 1) * per])
            short = []
            for j, r in enumerate(group):
                if len(r) < 3:                         # too short for hierarchical expansion -> dummy (:131-136)
                    short.append(r)
                    group[j] = np.array([INF_ROW, tab.zero_row, INF_ROW], dtype=np.int64)
            mids = [len(r) // 2 for r in group]
            via = int(group[0][mids[0]])               # "across batch dimension all of the subgoals are identical"
            plan_l, cost_l = child_left._optimize(tab, [r[:m] for r, m in zip(group, mids)], via)
            plan_r, cost_r = child_right._optimize(tab, [r[m:] for r, m in zip(group, mids)], goal_row)
            plan, cost = plan_l + plan_r, cost_l + cost_r
            if short:
                plan_s, cost_s, _ = self._best_of_n_segments(tab, short, goal_row)
                if cost_s < cost or np.isnan(cost):
                    plan, cost = plan_s, cost_s
            best_plans.append(plan)
            best_costs.append(cost)
        k = _choose(np.array(best_costs, dtype=np.float32).reshape(len(best_costs)))
        return best_plans[k], best_costs[k]

    @property
    def fully_optimized(self):
        if self._children is None:
            return self._is_optimized
        return self._is_optimized and all(c.fully_optimized for c in self._children[0]) \
            and all(c.fully_optimized for c in self._children[1])
