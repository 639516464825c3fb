"""Checkpoint format of the reference (gcp/prediction/training/checkpoint_handler.py:14-160, train.py:113-122).

A checkpoint is one `torch.save`d dict `{'epoch', 'global_step', 'state_dict', 'optimizer'}` named
`weights_ep<epoch>.pth` in `<exp_path>/weights`.  `TreeModel` / `SequentialModel` register their tensors under the
reference's state-dict keys, so `state_dict` goes straight into `model.load_state_dict`; the engine repacks the
weights (bf16 K-major tiles, folded eval-BatchNorm, composed decoder matrices) on the next call.

The logic lives in the module-level functions; `CheckpointHandler` only re-exports them under the static-method names
the reference's call sites use (planner_policy.py:49-50, train.py:60-66, cost_mdl.py:134-135), with the same
arguments, return values and exceptions.  The reference's one-off "hack_to_fix_checkpoints" scripts are not carried over.
"""
import os
import re

import torch

_CKPT_RE = re.compile(r"^weights_ep(-?\d+)\.pth$")


class NoCheckpointsException(Exception):
    pass


def checkpoint_name(epoch):
    return "weights_ep%s.pth" % (epoch,)


def list_epochs(folder):
    """Epoch numbers of the `weights_ep<N>.pth` files in `folder`; other *.pth files are ignored."""
    folder = os.path.abspath(folder)
    files = [f for f in (os.listdir(folder) if os.path.isdir(folder) else []) if f.endswith(".pth")]
    if not files:
        print("Warning: No checkpoints found at {}!".format(folder))
        raise NoCheckpointsException
    return [int(m.group(1)) for m in map(_CKPT_RE.match, files) if m]


def resolve_checkpoint(which, folder):
    """'latest' -> highest epoch; an integer (or its string) -> that epoch; anything else is a file name, '.pth' added
    when missing."""
    text = str(which)
    if text == "latest":
        name = checkpoint_name(max(list_epochs(folder)))
    elif re.fullmatch(r"-?\d+", text):
        name = checkpoint_name(text)
    else:
        name = text if ".pth" in text else text + ".pth"
    return os.path.join(folder, name)


def submodule_state(state_dict, prefix):
    """Entries below `prefix.` with the prefix stripped (loading a sub-network, e.g. 'cost_mdl', from a full checkpoint)."""
    if prefix is None:
        return state_dict
    cut = len(prefix) + 1
    picked = {key[cut:]: value for key, value in state_dict.items() if key.startswith(prefix)}
    if len(picked) == 0:
        raise ValueError("Did not find submodule {} in checkpoint!".format(prefix))
    return picked


def rename_keys(state_dict, old, new):
    for key in list(state_dict):
        if old in key:
            state_dict[key.replace(old, new)] = state_dict.pop(key)


def load_checkpoint(path, model, with_step_and_optimizer=False, optimizer=None, strict=True, submodule=None):
    """Reads the checkpoint to host memory and loads it into `model` (the engine uploads / repacks lazily).  Returns True,
    or (global_step, next_epoch, True) when the training position and optimiser state are restored as well."""
    if not os.path.isfile(path):
        raise ValueError("Could not find checkpoint file in {}!".format(path))
    blob = torch.load(path, map_location="cpu", weights_only=False)
    model.load_state_dict(submodule_state(blob["state_dict"], submodule), strict=strict)
    if not with_step_and_optimizer:
        return True
    try:
        optimizer.load_state_dict(blob["optimizer"])
    except (RuntimeError, ValueError):
        if strict:
            raise
        print("Could not load optimizer params because of changes in the network + non-strict loading")
    return blob["global_step"], blob["epoch"] + 1, True


def save_checkpoint(folder, model, epoch, global_step=0, optimizer=None):
    """The file ModelTrainer.save_checkpoint writes (train.py:113-122) -- readable by the loader above and by the
    reference's own."""
    os.makedirs(folder, exist_ok=True)
    path = os.path.join(folder, checkpoint_name(epoch))
    torch.save(dict(epoch=epoch, global_step=global_step, state_dict=model.state_dict(),
                    optimizer={} if optimizer is None else optimizer.state_dict()), path)
    return path


class CheckpointHandler:
    """The reference's static interface (same names, argument order and behaviour)."""
    get_ckpt_name = staticmethod(checkpoint_name)
    get_epochs = staticmethod(list_epochs)
    get_resume_ckpt_file = staticmethod(resolve_checkpoint)
    filter = staticmethod(submodule_state)
    rename_parameters = staticmethod(rename_keys)
    save_checkpoint = staticmethod(save_checkpoint)

    @staticmethod
    def load_weights(weights_file, model, load_step_and_opt=False, optimizer=None, dataset_length=None, strict=True,
                     submodule_name=None):
        return load_checkpoint(weights_file, model, load_step_and_opt, optimizer, strict, submodule_name)
