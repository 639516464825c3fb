"""Edge cases of the C ABI on the GPU: smallest / ragged batches, extreme lengths (work lists that are empty from some level
on), refused argument combinations.  Everything goes through ctypes (video_gcp_b200/_C.py)."""
import numpy as np
import pytest
import torch

from video_gcp_b200 import _C
from video_gcp_b200.synthetic import synthetic_rollout_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need the B200 box"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def engine(dev, sd):
    from video_gcp_b200.engine import Engine
    eng = Engine(dev, max_candidates=256, attach_cost_mdl=True)
    eng.load_weights(sd)
    yield eng
    eng.close()


def test_single_and_ragged_batches_are_subsets_of_a_full_batch(engine, dev):
    """B = 1 and B = 129 (a second 128-row tile holding ONE candidate): every output equals the corresponding rows of a
    256-candidate rollout bit for bit, in the reference mode and in planner mode."""
    inp = synthetic_rollout_inputs(256, seed=61, shared_images=True)
    I0, Ig, z, ei = inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["z"].to(dev), inp["end_ind"].to(dev)
    keys = ("e_df", "images_df", "actions", "regressed_state", "model_enc_seq", "existence", "seq_len_logits")
    full = {k: v.clone() for k, v in engine.rollout(I0, Ig, z, end_ind=ei, images_shared=True, l2_goal=Ig[0]).items()}
    for B in (1, 129):
        part = engine.rollout(I0, Ig, z[:B].contiguous(), end_ind=ei[:B].contiguous(), images_shared=True, l2_goal=Ig[0], fresh=True)
        for k in keys + ("l2_cost",):
            assert torch.equal(part[k], full[k][:B]), (B, k)
        kept = engine.rollout(I0, Ig, z[:B].contiguous(), end_ind=ei[:B].contiguous(), images_shared=True, l2_goal=Ig[0], fresh=True,
                              decode_kept_only=True, tree_kept_only=True, want_existence=False, want_images=False)
        assert torch.equal(kept["l2_cost"], full["l2_cost"][:B])
        assert torch.equal(kept["actions"], full["actions"][:B]) and torch.equal(kept["model_enc_seq"], full["model_enc_seq"][:B])


@pytest.mark.parametrize("length", [1, 2, 3, 199])
def test_uniform_extreme_lengths_in_planner_mode(engine, dev, length):
    """Every candidate at the same extreme length: with end_ind = 1 .. 3 the work lists of the deeper tree levels are EMPTY
    (launches that find zero rows), with 199 they are full; costs and pruned outputs equal the unpruned rollout."""
    B = 200
    inp = synthetic_rollout_inputs(B, seed=62, shared_images=True)
    I0, Ig, z = inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["z"].to(dev)
    ei = torch.full((B,), length, dtype=torch.long, device=dev)
    kw = dict(end_ind=ei, images_shared=True, l2_goal=Ig[0], want_existence=False, fresh=True)
    full = engine.rollout(I0, Ig, z, **kw)
    kept = engine.rollout(I0, Ig, z, decode_kept_only=True, tree_kept_only=True, **kw)
    torch.cuda.synchronize()
    assert torch.isfinite(kept["l2_cost"]).all() and torch.equal(kept["l2_cost"], full["l2_cost"])
    for k in ("actions", "regressed_state", "model_enc_seq"):
        assert torch.equal(kept[k], full[k]), k
    assert torch.equal(engine.prune_gather(kept["images_df"], ei), engine.prune_gather(full["images_df"], ei))
    assert float(kept["model_enc_seq"][:, length + 1:].abs().max() if length < 199 else 0.0) == 0.0


def test_refused_arguments(engine, dev, sd):
    from video_gcp_b200.engine import Engine
    inp = synthetic_rollout_inputs(4, seed=63, shared_images=True)
    I0, Ig, z = inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["z"].to(dev)
    big = synthetic_rollout_inputs(300, seed=63, shared_images=True)["z"].to(dev)
    with pytest.raises(_C.GcpB200Error, match="max_candidates"):
        engine.rollout(I0, Ig, big, images_shared=True)
    with pytest.raises(_C.GcpB200Error, match="tree_kept_only"):
        engine.rollout(I0, Ig, z, images_shared=True, tree_kept_only=True)                          # needs decode_kept_only
    with pytest.raises(_C.GcpB200Error, match="tree_kept_only"):
        engine.rollout(I0, Ig, z, images_shared=True, decode_kept_only=True, tree_kept_only=True)   # existence requested
    with pytest.raises(_C.GcpB200Error, match="bad arguments"):
        engine.topk(torch.zeros(5, device=dev), 6)
    with pytest.raises(_C.GcpB200Error, match="unsupported tree shape"):
        Engine(dev, max_candidates=128, hierarchy_levels=9)
    with pytest.raises(_C.GcpB200Error, match="unsupported tree shape"):
        Engine(dev, max_candidates=128, hierarchy_levels=6, max_seq_len=100)                        # 100 frames need 7 levels
    seq = Engine(dev, max_candidates=128, model="sequential")
    with pytest.raises(_C.GcpB200Error, match="weights not loaded"):
        seq.seq_rollout(I0, Ig, torch.zeros(4, 199, 256, device=dev), images_shared=True)
    seq.close()
    # the engine is still usable after refused calls
    out = engine.rollout(I0, Ig, z, images_shared=True)
    assert torch.isfinite(out["e_df"]).all()
