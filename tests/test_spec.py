"""State-dict compatibility: the key/shape inventory equals the reference's (manifest generated from
the reference by oracle/make_golden.py)."""
import json
import os

from video_gcp_b200 import hparams, spec
from video_gcp_b200.synthetic import synthetic_state_dict


def _manifest(golden_dir):
    with open(os.path.join(golden_dir, "state_dict_manifest.json")) as f:
        return json.load(f)


def test_planner_manifest(golden_dir):
    hp = hparams.build_hparams(hparams.gcp_tree_25room_config(batch_size=1))
    mine = {k: list(v) for k, v in spec.full_manifest(hp).items()}
    ref = _manifest(golden_dir)["planner"]
    assert set(mine) == set(ref)
    assert all(mine[k] == ref[k] for k in ref)


def test_training_manifest(golden_dir):
    hp = hparams.build_hparams(hparams.gcp_tree_25room_config(batch_size=1, attach_cost_mdl=True))
    mine = {k: list(v) for k, v in spec.full_manifest(hp).items()}
    ref = _manifest(golden_dir)["training"]
    assert set(mine) == set(ref)
    assert all(mine[k] == ref[k] for k in ref)


def test_synthetic_weights_deterministic_and_aliased(hp):
    a = synthetic_state_dict(hp, 5)
    b = synthetic_state_dict(hp, 5)
    c = synthetic_state_dict(hp, 6)
    k = "tree_module.tree_modules.3.subgoal_pred.lstm.1.weight_hh"
    assert (a[k] == b[k]).all() and not (a[k] == c[k]).all()
    assert a["dense_rec.decoder.net.gen_head.conv.weight"].data_ptr() == a["decoder.net.gen_head.conv.weight"].data_ptr()
    bias = a["tree_module.tree_modules.0.subgoal_pred.lstm.0.bias_ih"]
    assert bias[512:1024].mean() > 0.8 and abs(bias[:512].mean()) < 0.05


def test_override_defaults_rules():
    import pytest
    hp = hparams.default_hparams()
    with pytest.raises(ValueError):
        hp.override_defaults({"ngf": 4})            # identical to default -> error, as in the reference
    with pytest.raises(AttributeError):
        hp.override_defaults({"no_such_param": 1})
