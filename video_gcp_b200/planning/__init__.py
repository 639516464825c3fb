"""Drop-in CEM planning stack (simulator, cost functions, samplers, planner) over libgcpb200."""
from .cem_simulator import GCPSimulator, GCPImageSimulator  # noqa: F401
from .cost_fcn import L2ImageCost, LearnedCostEstimate, ImageLearnedCostEstimate, ImageWrappedLearnedCostFcn  # noqa: F401
from .sampler import (FlatCEMSampler, SimpleTreeCEMSampler, HierarchicalTreeCEMSampler,  # noqa: F401
                      ImageHierarchicalTreeCEMSampler)
from .cem_planner import CEMPlanner, ImageCEMPlanner, HierarchicalCEMPlanner, HierarchicalImageCEMPlanner  # noqa: F401
from .tree_optimizer import ImageHierarchicalTreeLatentOptimizer  # noqa: F401
