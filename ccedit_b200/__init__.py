"""ccedit_b200 - B200-native (sm_100a) implementation of CCEdit's denoising hot path.

Host side: a PyTorch-facing mirror of the reference's network classes (same constructor kwargs, forward signatures
and state-dict keys as sgm.modules.diffusionmodules.controlmodel / wrappers); device side: hand-written CUDA kernels
behind the C ABI in include/ccedit_b200.h.
"""
__version__ = "0.1.0"
