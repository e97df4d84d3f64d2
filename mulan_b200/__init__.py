"""mulan_b200 -- B200-native (sm_100a) MuLAN noise-schedule + ELBO hot path.

Only the pieces the hot path needs live here:
  csrc/      hand-written CUDA kernels + the C ABI (include/mulan_b200.h)
  _lib.py    ctypes loader of libmulan_b200.so (no fallback)
  ops.py     raw launches + torch.autograd bindings
  model.py   host-side mirror of the reference's VDM / VDMOutput / loss_fn interface
  dist.py    one-process-per-GPU data-parallel plumbing (NCCL / gloo)
"""
__version__ = '0.1.0'
