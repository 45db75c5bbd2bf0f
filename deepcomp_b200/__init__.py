"""deepcomp_b200 -- B200-native batched implementation of the DeepCoMP mobile-cellular env step.

Product path only: hand-written sm_100a CUDA behind the C ABI in include/deepcomp_b200.h, driven from Python over
PyTorch device tensors.  Nothing here imports ``oracle/`` and there is no CPU fallback.
"""
from .batched import BatchedMobileEnv, env_seeds, sharing_for_bs  # noqa: F401

__version__ = '0.1.0'
