"""Host-side integer helpers for balanced pruning (index glue only; the device kernel
`prune_map_kernel` is what the rollout uses)."""
from functools import lru_cache


@lru_cache(maxsize=None)
def frame_nodes(end_ind, depth=8):
    """Depth-first node index of frame t, t = 0..end_ind: interval recursion from (-1, end_ind+1), midpoint
    with truncating division, node kept iff its timestep differs from both ends
    (gcp/prediction/models/tree/frame_binding.py:42-65)."""
    out = [0] * (end_ind + 1)

    def rec(l, r, lvl, j):
        if lvl == depth:
            return
        t = int((l + r) / 2)
        if t != l and t != r:
            out[t] = (2 * j + 1) * 2 ** (depth - 1 - lvl) - 1
        rec(l, t, lvl + 1, 2 * j)
        rec(t, r, lvl + 1, 2 * j + 1)

    rec(-1, end_ind + 1, 0, 0)
    return tuple(out)
