"""video_gcp_b200 -- B200-native GCP-tree CEM rollout (drop-in for orybkin/video-gcp's hot path).

Host side: Python mirroring the reference's module / simulator / cost / sampler / planner interfaces.
Device side: one C-ABI shared library (`libgcpb200.so`, built from `csrc/` for sm_100a) that holds every
kernel.  There is no CPU fallback: using the model without the library or without a CUDA device raises.
"""
from .types import AttrDict, ParamDict  # noqa: F401

__version__ = "0.1.0"
