"""Drop-in CEM planning stack (simulator, cost functions, samplers, planner) over libgcpb200."""
from .cem_simulator import GCPSimulator, GCPImageSimulator  # noqa: F401
from .cost_fcn import L2ImageCost, LearnedCostEstimate, ImageLearnedCostEstimate, ImageWrappedLearnedCostFcn  # noqa: F401
from .sampler import FlatCEMSampler, SimpleTreeCEMSampler  # noqa: F401
from .cem_planner import CEMPlanner, ImageCEMPlanner  # noqa: F401
