"""Planning policies over the B200 rollout library (reference: gcp/planning/planner_policy.py:13-227).

`ImageCEMPolicy` keeps the reference's call surface -- `policy = ImageCEMPolicy(ag_params, policyparams)`,
`policy.reset()`, `policy.act(t, i_tr, state, images, goal_image) -> AttrDict(actions=...)` -- and its behaviour:
(re)plan with the configured CEM planner when there is no plan / the plan is used up / the replan interval hits
(:101-107), then either play back the planned actions or, with `closed_loop_execution`, re-infer every action from
the CURRENT image and the next latent of the plan with the inverse model (:208-221).  The closed-loop step is one
`gcpb200_infer_action` call (encoder + inverse-model MLP on the device, 8 bytes back).

Only what the reference's infrastructure (gcp/planning/infra) provides around the policy -- agent, environment,
logging -- is left out; `log_outputs_stateful` just clears the planner logs.
"""
import numpy as np
import torch

from ..checkpoint_handler import CheckpointHandler
from ..model import TreeModel
from ..types import AttrDict
from .cem_simulator import GCPImageSimulator

_DEFAULTS = dict(
    params={}, model_cls=None, checkpt_path=None, load_epoch=None, logger=None, verbose=False, max_dump_rollouts=5,
    replan_interval=1, num_max_replans=10,                 # PlannerPolicy._default_hparams (:67-84)
    cem_planner=None, cem_params={},                       # CEMPolicy._default_hparams (:133-143)
    closed_loop_execution=False, act_cond=False,           # ImageCEMPolicy._default_hparams (:188-197)
    state_dict=None, max_candidates=128,                   # ours: weights passed in memory; engine workspace size
)


def resume_ckpt_file(resume, path):
    """CheckpointHandler.get_resume_ckpt_file (gcp/prediction/training/checkpoint_handler.py:31-43)."""
    return CheckpointHandler.get_resume_ckpt_file(resume, path)


class ImageCEMPolicy:
    """CEM planning policy for image-based tasks; follows the plan open loop or closed loop via the inverse model."""

    def __init__(self, ag_params, policyparams, gpu_id=None, ngpu=None, conversion_fcns=None, n_rooms=None):
        hp = AttrDict(_DEFAULTS)
        for k, v in policyparams.items():
            if k == 'type':
                continue
            if k not in hp:
                raise AttributeError("unknown policy parameter %r" % k)
            hp[k] = v
        self._hp = hp
        if hp.act_cond:
            raise NotImplementedError("the action-conditioned simulator is not on the GCP-tree planner path")
        self.verbose = hp.verbose
        self.log_dir = getattr(ag_params, "log_dir", None) if not isinstance(ag_params, dict) else ag_params.get("log_dir")
        T = ag_params["T"] if isinstance(ag_params, dict) else ag_params.T
        params = AttrDict(hp.params)
        params['batch_size'] = 1
        self.max_seq_len = T
        if 'max_seq_len' not in params:
            params['max_seq_len'] = T

        model_cls = hp.model_cls if hp.model_cls is not None else TreeModel
        self.planner = model_cls(params, None, max_candidates=hp.max_candidates)
        if not torch.cuda.is_available():
            raise RuntimeError("ImageCEMPolicy needs a CUDA device (B200); there is no CPU path")
        self.device = torch.device('cuda', gpu_id if gpu_id is not None else torch.cuda.current_device())
        self.planner.to(self.device)
        self.planner.device = self.device
        self.planner._hp.device = self.device
        if hp.state_dict is not None:
            self.planner.load_state_dict(hp.state_dict, strict=False)
        else:
            # planner_policy.py:48-50
            f = CheckpointHandler.get_resume_ckpt_file('latest' if hp.load_epoch is None else hp.load_epoch, hp.checkpt_path)
            CheckpointHandler.load_weights(f, self.planner, strict=False)
        self.planner.eval()

        cem_params = AttrDict(hp.cem_params)
        cem_params.update({'max_seq_len': params['max_seq_len']})
        self._cem_simulator = GCPImageSimulator(self.planner, append_latent=True)
        self._cem_planner = hp.cem_planner(cem_params, self._cem_simulator)
        self.reset()

    def reset(self):
        self.current_exec_step = None
        self.action_plan = None
        self.image_plan = None
        self.latent_plan = None
        self.num_replans = 0
        self.planner_outputs = []

    def act(self, t=None, i_tr=None, state=None, images=None, goal_image=None):
        """images: executed trajectory so far [T,1,H,W,3]; goal_image [1,H,W,3] (planner_policy.py:199-202,86-111)."""
        self._images = images[:, 0]
        self._states = state
        self.t, self.i_tr, self.goal_image = t, i_tr, goal_image
        output = AttrDict()
        if self.image_plan is None \
                or self.image_plan.shape[0] - 1 <= self.current_exec_step \
                or (t % self._hp.replan_interval == 0 and self.num_replans < self._hp.num_max_replans):
            self._plan(images[t], goal_image, t)
            self.num_replans += 1
        output.actions = self.get_action(images[t])
        self.current_exec_step = self.current_exec_step + 1
        return output

    def _plan(self, state, goal, step):
        """Planner directly outputs the action plan via the inverse model (:204-208)."""
        self.image_plan, self.action_plan, self.latent_plan, self.plan_cost = self._cem_planner(state, goal)
        self.current_exec_step = 0

    def get_action(self, current_image):
        if self._hp.closed_loop_execution:
            return self._infer_action(current_image, self.latent_plan[self.current_exec_step + 1])
        assert self.action_plan is not None     # need to attach inverse model to planner to get actions!
        if self.action_plan.size < 1:
            return 0.05 * np.random.rand(2, )
        return self.action_plan[self.current_exec_step]

    def _infer_action(self, current_img, target_latent):
        """Closed-loop execution action from the inverse model (:215-221)."""
        img = torch.as_tensor(np.asarray(current_img), dtype=torch.float32).pin_memory().to(self.device, non_blocking=True)
        img = self._cem_simulator._env2planner(img)
        tgt = torch.as_tensor(np.asarray(target_latent, dtype=np.float32)).pin_memory().to(self.device, non_blocking=True)
        return self.planner.engine.infer_action(img, tgt[None])[0].cpu().numpy()

    def log_outputs_stateful(self, logger=None, global_step=None, phase=None, dump_dir=None, **unused):
        self._cem_planner.log_verbose(logger, global_step, phase, getattr(self, "i_tr", 0), dump_dir)
