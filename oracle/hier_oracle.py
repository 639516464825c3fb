"""CPU oracle of the hierarchical GCP-tree latent optimiser (numpy restatement).

TEST INFRASTRUCTURE ONLY (same rules as gcp_oracle.py: imported by tests/ only, never by the product).

Restates, in a flat functional style, what the reference's image-based hierarchical planner computes:
  gcp/planning/tree_optimizer.py:7-203   HierarchicalTreeLatentOptimizer / ImageHierarchicalTreeLatentOptimizer
  gcp/planning/cem/sampler.py:83-143     HierarchicalTreeCEMSampler / ImageHierarchicalTreeCEMSampler
  gcp/planning/cem/cem_planner.py:55-96,166-218   CEMPlanner.__call__ with HierarchicalCEMPlanner._get_best_rollouts
  gcp/planning/cem/cost_fcn.py:79-101    LearnedCostEstimate (ndarray and list branches)

Parity status: PINNED by oracle/make_golden_hier.py, which runs the unmodified reference planner
(HierarchicalImageCEMPlanner + ImageHierarchicalTreeCEMSampler + ImageLearnedCostEstimate over the reference
TreeModel / GCPImageSimulator) with np.random seeded and stores its samples, per-iteration choices, costs and plan in
tests/golden/hier_plan.npz; tests/test_oracle_hier.py replays this restatement against it.

A "rollout" is a numpy array [L, 3072 + 128] (flattened image, latent) exactly as GCPSimulator.rollout returns it with
append_latent=True.  `normal(loc, scale, size)` is the Gaussian source (np.random.normal in the reference); drawing
order matters and is kept, including the draws whose results are thrown away.
"""
import numpy as np
import torch

from . import gcp_oracle as O

LAT = 128


def injected_end_ind(call_idx, n):
    """Rollout lengths injected into the model (reference, oracle and product alike) for the parity fixtures: a fixed
    function of the rollout-call index.  Candidates 1 and 4 are short on purpose so that the optimiser's too-short /
    dummy-sequence branches (tree_optimizer.py:131-152) run."""
    r = np.random.default_rng([77, int(call_idx)])
    e = r.integers(8, 200, size=n)
    if n > 1:
        e[1] = 2 + call_idx
    if n > 4:
        e[4] = 4
    return e.astype(np.int64)


ARGMIN_LOG = []     # (choice, costs) of every argmin the optimiser takes, in order (compared with the reference's)


def _argmin(a):
    k = int(np.argmin(a))                      # NaN counts as the minimum, first one wins -- as in the reference
    ARGMIN_LOG.append((k, np.asarray(a, dtype=np.float64).reshape(-1).copy()))
    return k


def pair_cost(sd, a, b):
    """LearnedCostEstimate.__call__, ndarray branch (cost_fcn.py:84-87) -> [n,1]."""
    with torch.no_grad():
        x = torch.cat([torch.as_tensor(np.asarray(a), dtype=torch.float32),
                       torch.as_tensor(np.asarray(b), dtype=torch.float32)], 1)
        return O.mlp(sd, "cost_mdl.cost_pred", x, conv=False).numpy()


def path_cost(sd, seqs, goals):
    """LearnedCostEstimate.__call__, list branch (cost_fcn.py:88-97) -> [n]."""
    out = []
    with torch.no_grad():
        for s, g in zip(seqs, goals):
            x = torch.cat([torch.as_tensor(np.asarray(s), dtype=torch.float32),
                           torch.as_tensor(np.asarray(g), dtype=torch.float32)])
            c = O.mlp(sd, "cost_mdl.cost_pred", torch.cat([x[:-1], x[1:]], 1), conv=False)
            out.append(c.sum().numpy())
    return np.array(out)


def _img(r):
    """image part of a rollout [L,3200] -> [L,3,32,32] (tree_optimizer.py:181-190)."""
    flat = r[..., :-LAT]
    return flat.reshape(flat.shape[0], 3, 32, 32)


def _lat(r):
    return r[..., -LAT:]


def _dummy(frame):
    """tree_optimizer.py:158-162: a 3-frame stand-in whose cost is maximal (inf / 0 / inf)."""
    return np.stack([np.full_like(frame, np.inf), np.zeros_like(frame), np.full_like(frame, np.inf)])


class LatentNode:
    """One layer-node of the optimiser tree (tree_optimizer.py:11-44)."""

    def __init__(self, sd, dim, rates, depth, n_final, normal):
        self.sd, self.dim, self.depth, self.normal = sd, dim, depth, normal
        self.done, self.best_z, self.last_z = False, None, None
        self.trace = None            # filled by solve(): what was chosen here (for the parity fixtures)
        if rates:
            self.n, self.n_lat = rates[0], 1
            mk = lambda: LatentNode(sd, dim, list(rates[1:]), depth - 1, n_final, normal)
            self.left = [mk() for _ in range(self.n)]
            self.right = [mk() for _ in range(self.n)]
        else:
            self.n, self.n_lat = n_final, 2 ** depth - 1
            self.left = self.right = None

    # ---- sampling (tree_optimizer.py:46-72) ----
    def draw(self, below=False):
        if self.done:
            z = self.best_z.copy()[None]
        else:
            z = self.normal(np.zeros((self.n_lat, self.dim)), np.ones((self.n_lat, self.dim)),
                            (self.n, self.n_lat, self.dim))
            if below:
                z = z[:1]
            self.last_z = z.copy()
        kids_below = below or not self.done
        if self.left is None:
            return z
        rows = []
        for l, r, zi in zip(self.left, self.right, z):
            zl, zr = l.draw(kids_below), r.draw(kids_below)
            assert zl.shape == zr.shape
            rows.append(np.concatenate([zl, np.tile(zi[0], (zl.shape[0], 1, 1)), zr], 1))
        return np.concatenate(rows)

    # ---- optimisation (tree_optimizer.py:74-150) ----
    def solve(self, rollouts, goal):
        if self.left is None:
            return self._segment(rollouts, goal)
        if not self.done:
            return self._subgoal(rollouts, goal)
        return self._descend(rollouts, goal)

    def _best_segment(self, rollouts, goal):
        """tree_optimizer.py:145-150 with the image variant's inputs (:167-174)."""
        lats = [_lat(r) for r in rollouts]
        if goal.ndim > 2:
            goals = [l[-1:] for l in lats]
        else:
            goals = [_lat(goal[None]) for _ in lats]
        cost = path_cost(self.sd, lats, goals)
        k = _argmin(cost)
        return _img(rollouts[k]), cost[k], k, cost

    def _segment(self, rollouts, goal):
        best, cost, k, all_cost = self._best_segment(rollouts, goal)
        self.best_z, self.done = self.last_z[k], True
        self.trace = dict(kind="segment", choice=k, costs=np.asarray(all_cost, dtype=np.float64))
        return best, cost

    def _subgoal(self, rollouts, goal):
        mid = [int(np.floor(r.shape[0] / 2)) for r in rollouts]
        frame_goal = goal.shape[-1] == rollouts[0].shape[-1]       # a frame of a parent rollout vs the goal image
        s_lat = np.stack([_lat(r)[0] for r in rollouts])
        m_lat = np.stack([_lat(r)[m] for r, m in zip(rollouts, mid)])
        g_lat = np.stack([_lat(goal[None])[0] if frame_goal else _lat(r)[-1] for r in rollouts])
        total = pair_cost(self.sd, s_lat, m_lat) + pair_cost(self.sd, m_lat, g_lat)
        k = _argmin(total)
        self.best_z = self.last_z[k]
        plan = [_img(rollouts[k])[0]]
        sub = _img(rollouts[k])[mid[k]]
        if (sub != plan[-1]).any():
            plan.append(sub)
        if not frame_goal:
            plan.append(goal[0].transpose(2, 0, 1))               # raw [0,1] goal image, as the reference appends it
        self.left, self.right = self.left[:1], self.right[:1]
        self.n, self.done = 1, True
        self.trace = dict(kind="subgoal", choice=k, costs=np.asarray(total, dtype=np.float64).reshape(-1))
        return np.stack(plan), total[k]

    def _descend(self, rollouts, goal):
        n_all = len(rollouts)
        bounds = np.linspace(0, n_all, self.n + 1).astype(int) if n_all % self.n == 0 else None
        assert bounds is not None, "np.array_split with uneven parts is not on the planner path"
        plans, costs = [], []
        for i, (l, r) in enumerate(zip(self.left, self.right)):
            group = [x for x in rollouts[bounds[i]:bounds[i + 1]]]
            short = []
            for j, x in enumerate(group):
                if x.shape[0] < 3:
                    short.append(x)
                    group[j] = _dummy(x[0])
            mid = [int(np.floor(x.shape[0] / 2)) for x in group]
            via = group[0][mid[0]]
            pl, cl = l.solve([x[:m] for x, m in zip(group, mid)], via)
            pr, cr = r.solve([x[m:] for x, m in zip(group, mid)], goal)
            plan, cost = np.concatenate([pl, pr]), cl + cr
            if short:
                ps, cs, _, _ = self._best_segment(short, goal)
                if cs < cost or np.isnan(cost):
                    plan, cost = ps, cs
            plans.append(plan)
            costs.append(cost)
        k = _argmin(np.array(costs))
        return plans[k], costs[k]

    @property
    def complete(self):
        if self.left is None:
            return self.done
        return self.done and all(c.complete for c in self.left) and all(c.complete for c in self.right)


def plan(sd, rollout_fn, goal_img, rates=(10, 10), depth=8, n_final=5, n_iters=3, batch_size=10, dim=256,
         normal=np.random.normal):
    """CEMPlanner.__call__ (cem_planner.py:55-96) specialised to HierarchicalImageCEMPlanner +
    ImageHierarchicalTreeCEMSampler (sampler.py:130-143): per iteration sample -> roll out -> optimise one more layer
    -> (re)sample; final rollout of the single optimised latent tree.

    rollout_fn(samples [n,255,256] float64) -> list of n rollouts [L_i,3200] (float32).
    Returns dict(samples per iteration, plans, costs, final samples, final rollout)."""
    root = LatentNode(sd, dim, list(rates), depth, n_final, normal)
    log = dict(samples=[], plans=[], costs=[], n_rollouts=[])
    best = None
    for _ in range(n_iters):
        z = root.draw()
        ro = rollout_fn(z)
        best_plan, cost = root.solve(ro, goal_img)
        if (best_plan[-1] != goal_img[0].transpose(2, 0, 1)).any():          # sampler.py:137-139
            best_plan = np.concatenate((best_plan, goal_img.transpose(0, 3, 1, 2)))
        best = root.draw()                                                    # cem_planner.py:214
        log["samples"].append(z)
        log["plans"].append(best_plan)
        log["costs"].append(np.asarray(cost, dtype=np.float64).reshape(-1))
        log["n_rollouts"].append(len(ro))
    log["final_samples"] = best
    log["final_rollouts"] = rollout_fn(best)
    log["complete"] = root.complete
    log["root"] = root
    return log
