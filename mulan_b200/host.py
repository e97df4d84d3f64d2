"""Host-buffer entry (mulan_elbo_host): numpy / CPU-tensor in, numpy out, one C-ABI call.

This is the call a user with HOST arrays makes (and what bench.py's ``e2e`` leg times):
H2D of the inputs, all ELBO kernels, D2H of losses, scalars and gradients happen inside the
one call.  Page-locked inputs/outputs (``torch.empty(..., pin_memory=True).numpy()``) make
the copies true DMA.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import numpy as np

from . import _lib


def _np(v, dtype, shape, name):
  if hasattr(v, 'numpy') and not isinstance(v, np.ndarray):
    v = v.numpy()            # CPU torch tensor (possibly pinned) -> zero-copy view
  if not isinstance(v, np.ndarray) or v.dtype != dtype or tuple(v.shape) != tuple(shape):
    raise TypeError(f'{name}: expected {np.dtype(dtype).name} array of shape {tuple(shape)}')
  if not v.flags['C_CONTIGUOUS']:
    raise ValueError(f'{name}: must be C-contiguous')
  return v


def _ptr(v):
  return None if v is None else C.c_void_p(v.ctypes.data)


class HostOutputs:
  """Preallocated (optionally pinned) output buffers, reusable across calls."""

  def __init__(self, rows: int, dim: int = 3072, want_grad: bool = True, pinned: bool = False):
    def alloc(*shape):
      if pinned:
        import torch
        return torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()
      return np.empty(shape, dtype=np.float32)
    self.losses = alloc(3, rows)
    self.scalars = alloc(6)
    self.a_bar = alloc(rows, dim) if want_grad else None
    self.b_bar = alloc(rows, dim) if want_grad else None
    self.c_bar = alloc(rows, dim) if want_grad else None
    self.n_bar = alloc(rows, dim) if want_grad else None


def elbo_host(x, a, b, c, t, eps0, eps, net=None, *, param: int = _lib.MULAN_PARAM_EPS,
              want_grad: bool = True, denoiser: Optional[Callable] = None,
              out: Optional[HostOutputs] = None, vocab: int = 256, gamma_min: float = -13.3,
              gamma_max: float = 5.0, jax_keys=None):
  """ELBO loss terms (+ gradients of bpd w.r.t. a, b, c and the denoiser output) for host
  arrays.  Mirrors VDM.__call__ + Experiment_VDM.loss_fn (ldm/model_mulan_epsilon.py:280-363,
  ldm/experiment_vdm.py:47-78) once (a, b, c) and the random draws exist.

  denoiser(rows, z_t_ptr, g_net_ptr, net_ptr, stream_ptr) -> int, raw device addresses of one
  chunk of `rows` examples (called once per chunk); when None the supplied host `net` is used.
  jax_keys = ((k0, k1) of eps_0, (k0, k1) of eps): instead of the eps0 / eps arrays (pass None
  for them), the raw threefry keys of the two jax.random.normal draws of VDM.__call__
  (ldm/model_mulan_epsilon.py:315, :327); the draws are then made on the device
  (mulan_elbo_host_keyed) and never cross PCIe.
  Returns dict(loss_recon, loss_klz_prior, loss_diff, scalars[, a_bar, b_bar, c_bar, n_bar]).
  """
  B, D = a.shape
  x = _np(x, np.uint8, (B, D), 'x')
  a, b, c = (_np(v, np.float32, (B, D), n) for v, n in ((a, 'a'), (b, 'b'), (c, 'c')))
  if jax_keys is None:
    eps0, eps = (_np(v, np.float32, (B, D), n) for v, n in ((eps0, 'eps0'), (eps, 'eps')))
  else:
    if eps0 is not None or eps is not None:
      raise ValueError('pass either the eps0 / eps arrays or jax_keys, not both')
    keys = [np.asarray(k, dtype=np.uint32).reshape(2) for k in jax_keys]
    if len(keys) != 2:
      raise ValueError('jax_keys = (key of eps_0, key of eps), two uint32 words each')
  t = _np(t, np.float32, (B,), 't')
  if net is not None:
    net = _np(net, np.float32, (B, D), 'net')
  out = out or HostOutputs(B, D, want_grad)
  cb = _lib.DENOISER_FN()
  if denoiser is not None:
    cb = _lib.DENOISER_FN(lambda user, rows, z, g, n, s: int(denoiser(rows, z, g, n, s) or 0))
  d = _lib.make_desc(B, D, vocab, param, _lib.MULAN_GT_MEAN, 0, gamma_min, gamma_max)
  tail = (_ptr(net), cb, None, 1 if want_grad else 0, _ptr(out.losses), _ptr(out.scalars),
          _ptr(out.a_bar), _ptr(out.b_bar), _ptr(out.c_bar), _ptr(out.n_bar))
  head = (C.byref(d), _ptr(x), _ptr(a), _ptr(b), _ptr(c), _ptr(t))
  if jax_keys is None:
    _lib.check(_lib.load().mulan_elbo_host(*head, _ptr(eps0), _ptr(eps), *tail))
  else:
    _lib.check(_lib.load().mulan_elbo_host_keyed(*head, _ptr(keys[0]), _ptr(keys[1]), *tail))
  res = dict(loss_recon=out.losses[0], loss_klz_prior=out.losses[1], loss_diff=out.losses[2],
             scalars=out.scalars)
  if want_grad:
    res.update(a_bar=out.a_bar, b_bar=out.b_bar, c_bar=out.c_bar, n_bar=out.n_bar)
  return res
