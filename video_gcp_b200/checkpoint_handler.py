"""Checkpoint format of the reference (gcp/prediction/training/checkpoint_handler.py:14-160, train.py:113-122).

A checkpoint is one `torch.save`d dict `{'epoch', 'global_step', 'state_dict', 'optimizer'}` named
`weights_ep<epoch>.pth` in `<exp_path>/weights`.  `TreeModel` / `SequentialModel` register their tensors under the
reference's state-dict keys, so `state_dict` goes straight into `model.load_state_dict`; the engine repacks the
weights (bf16 K-major tiles, folded eval-BatchNorm, composed decoder matrices) on the next call.  Same static-method
surface and error behaviour as the reference class, minus its two one-off "hack_to_fix_checkpoints" scripts.
"""
import glob
import os

import torch


class NoCheckpointsException(Exception):
    pass


def _str2int(s):
    try:
        return int(s)
    except (TypeError, ValueError):
        return None


class CheckpointHandler:
    @staticmethod
    def get_ckpt_name(epoch):
        return 'weights_ep{}.pth'.format(epoch)

    @staticmethod
    def get_epochs(path):
        names = glob.glob(os.path.abspath(path) + "/*.pth")
        if len(names) == 0:
            print("Warning: No checkpoints found at {}!".format(path))
            raise NoCheckpointsException
        stems = [os.path.basename(f).replace('weights_ep', '').replace('.pth', '') for f in names]
        return [e for e in (_str2int(s) for s in stems) if e is not None]

    @staticmethod
    def get_resume_ckpt_file(resume, path):
        if resume == 'latest':
            resume_file = CheckpointHandler.get_ckpt_name(max(CheckpointHandler.get_epochs(path)))
        elif _str2int(resume) is not None:
            resume_file = CheckpointHandler.get_ckpt_name(resume)
        elif '.pth' not in resume:
            resume_file = resume + '.pth'
        else:
            resume_file = resume
        return os.path.join(path, resume_file)

    @staticmethod
    def filter(state_dict, submodule_key):
        """Keeps the entries under `submodule_key` and strips that prefix (checkpoint_handler.py:119-130)."""
        if submodule_key is None:
            return state_dict
        new_dict = {k[len(submodule_key) + 1:]: v for k, v in state_dict.items() if k.startswith(submodule_key)}
        if not new_dict:
            raise ValueError("Did not find submodule {} in checkpoint!".format(submodule_key))
        return new_dict

    @staticmethod
    def rename_parameters(state_dict, old, new):
        for key in [k for k in state_dict if old in k]:
            state_dict[key.replace(old, new)] = state_dict.pop(key)

    @staticmethod
    def load_weights(weights_file, model, load_step_and_opt=False, optimizer=None, dataset_length=None, strict=True,
                     submodule_name=None):
        """checkpoint_handler.py:45-76.  Tensors are read to host memory; the model's engine uploads and packs them."""
        if not os.path.isfile(weights_file):
            raise ValueError("Could not find checkpoint file in {}!".format(weights_file))
        checkpoint = torch.load(weights_file, map_location='cpu', weights_only=False)
        model.load_state_dict(CheckpointHandler.filter(checkpoint['state_dict'], submodule_name), strict=strict)
        if load_step_and_opt:
            try:
                optimizer.load_state_dict(checkpoint['optimizer'])
            except (RuntimeError, ValueError):
                if strict:
                    raise
                print("Could not load optimizer params because of changes in the network + non-strict loading")
            return checkpoint['global_step'], checkpoint['epoch'] + 1, True
        return True

    @staticmethod
    def save_checkpoint(folder, model, epoch, global_step=0, optimizer=None):
        """ModelTrainer.save_checkpoint (train.py:113-122): the file the loaders above -- and the reference's -- read."""
        os.makedirs(folder, exist_ok=True)
        state = {'epoch': epoch, 'global_step': global_step, 'state_dict': model.state_dict(),
                 'optimizer': optimizer.state_dict() if optimizer is not None else {}}
        path = os.path.join(folder, CheckpointHandler.get_ckpt_name(epoch))
        torch.save(state, path)
        return path
