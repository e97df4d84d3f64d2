"""jax.custom_vjp wrappers over the XLA-FFI handlers of mulan_xla_ffi.cc.

STATUS: NOT executed in this repository's image (JAX is not installable there).  It shows
the reference-side binding: `mulan_pre` / `mulan_post` replace the statements of
VDM.__call__ (ldm/model_mulan_epsilon.py:300-347, ldm/model_mulan_velocity.py:208-260)
around the `self.score_model(...)` call, and stay differentiable under
`jax.value_and_grad` / `pmap` / `scan` (ldm/experiment.py:89-91, 339).

Cotangents of (a, b, c): mulan_pre's vjp covers the paths through z_t and the denoiser's
noise-level input, mulan_post's vjp the path through loss_diff (two mulan_bwd_pre launches,
summed by XLA).  The PyTorch binding (mulan_b200/ops.py), which IS tested, fuses the two into
one launch by routing gL back through a `link` tensor.
"""
import ctypes
from functools import partial

import jax
import jax.numpy as jnp
import numpy as np

_so = ctypes.cdll.LoadLibrary('libmulan_xla_ffi.so')
for _name in ('MulanFwdPre', 'MulanFwdPost', 'MulanBwdPost', 'MulanBwdPre'):
  jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(_so, _name)), platform='CUDA')


def _attrs(cfg, param):
  return dict(vocab=np.int32(cfg.vocab_size), param=np.int32(param),
              gt_mode=np.int32(0 if cfg.unet_type == 'vdm' else 1),
              n_timesteps=np.int32(cfg.sm_n_timesteps),     # every handler: T scales the loss
              gamma_min=np.float64(cfg.gamma_min), gamma_max=np.float64(cfg.gamma_max))


@partial(jax.custom_vjp, nondiff_argnums=(0, 1))
def mulan_pre(cfg, param, x, a, b, c, t, eps0, eps):
  return _pre_fwd(cfg, param, x, a, b, c, t, eps0, eps)[0]


def _pre_fwd(cfg, param, x, a, b, c, t, eps0, eps):
  B, D = a.shape
  f32 = lambda *s: jax.ShapeDtypeStruct(s, jnp.float32)
  g_shape = (B,) if cfg.unet_type == 'vdm' else (B, D)
  z_t, g_net, w, rec, klz, var_sums = jax.ffi.ffi_call(
      'MulanFwdPre', (f32(B, D), f32(*g_shape), f32(B, D), f32(B), f32(B), f32(B, 2)))(
          x, a, b, c, t, eps0, eps, **_attrs(cfg, param))
  return (z_t, g_net, rec, klz, var_sums, w), (x, a, b, c, t, eps)


def _pre_bwd(cfg, param, res, cts):
  x, a, b, c, t, eps = res
  z_bar, g_bar, _, _, _, _ = cts     # recon / prior KL: fixed ends, zero gradient
  B, D = a.shape
  f32 = jax.ShapeDtypeStruct((B, D), jnp.float32)
  zeros_b = jnp.zeros((B,), jnp.float32)
  a_bar, b_bar, c_bar = jax.ffi.ffi_call('MulanBwdPre', (f32, f32, f32))(
      x, a, b, c, t, eps, jnp.zeros_like(eps), z_bar, g_bar, zeros_b, **_attrs(cfg, param))
  return (None, a_bar, b_bar, c_bar, None, None, None)


mulan_pre.defvjp(_pre_fwd, _pre_bwd)


@partial(jax.custom_vjp, nondiff_argnums=(0, 1))
def mulan_post(cfg, param, x, a, b, c, t, eps, w, net):
  B = a.shape[0]
  return jax.ffi.ffi_call('MulanFwdPost', jax.ShapeDtypeStruct((B,), jnp.float32))(
      x, a, b, c, t, eps, net, w, **_attrs(cfg, param))


def _post_fwd(cfg, param, x, a, b, c, t, eps, w, net):
  return mulan_post(cfg, param, x, a, b, c, t, eps, w, net), (x, a, b, c, t, eps, w, net)


def _post_bwd(cfg, param, res, gL):
  x, a, b, c, t, eps, w, net = res
  B, D = a.shape
  f32 = jax.ShapeDtypeStruct((B, D), jnp.float32)
  n_bar = jax.ffi.ffi_call('MulanBwdPost', f32)(x, a, b, c, t, eps, net, w, gL,
                                                **_attrs(cfg, param))
  g_shape = (B,) if cfg.unet_type == 'vdm' else (B, D)
  a_bar, b_bar, c_bar = jax.ffi.ffi_call('MulanBwdPre', (f32, f32, f32))(
      x, a, b, c, t, eps, net, jnp.zeros_like(eps), jnp.zeros(g_shape, jnp.float32), gL,
      **_attrs(cfg, param))
  return (None, a_bar, b_bar, c_bar, None, None, None, n_bar)


mulan_post.defvjp(_post_fwd, _post_bwd)
