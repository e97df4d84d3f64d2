"""jax.custom_vjp wrappers over the XLA-FFI handlers of mulan_xla_ffi.cc.

STATUS: NOT executed in this repository's image (JAX is not installable there).  It shows
the reference-side binding: `mulan_pre` / `mulan_post` replace the statements of
VDM.__call__ (ldm/model_mulan_epsilon.py:300-347, ldm/model_mulan_velocity.py:208-260)
around the `self.score_model(...)` call, and stay differentiable under
`jax.value_and_grad` / `pmap` / `scan` (ldm/experiment.py:89-91, 339).

ONE mulan_bwd_pre launch per backward pass, as in the tested PyTorch binding
(mulan_b200/ops.py): the complete cotangents of (a, b, c) -- the paths through z_t, through the
denoiser's noise-level input and through loss_diff -- are produced by mulan_pre's vjp.  What it
needs from the other side of the denoiser travels back as COTANGENTS of two carrier outputs of
mulan_pre that only mulan_post consumes:
    link      [B]     its cotangent, returned by mulan_post's vjp, is gL = d loss / d loss_diff
    link_net  [B, D]  its "cotangent" is the denoiser output `net` itself (mulan_post's residual)
Both carriers are zeros that no kernel reads (XLA removes them); a, b, c enter mulan_post under
stop_gradient, so mulan_post's vjp launches mulan_bwd_post only (37 B/sub-pixel of backward
traffic for the coefficients instead of 74 B with one mulan_bwd_pre per custom_vjp).

c_raw=True hands the kernels the PRE-ACTIVATION of dense_out_c (MULAN_FLAG_C_RAW): replace
`c = 1e-3 + nn.softplus(self.l3_c(h))` (ldm/model_mulan_epsilon.py:537) by `c = self.l3_c(h)`.
"""
import ctypes
from functools import partial

import jax
import jax.numpy as jnp
import numpy as np

_so = ctypes.cdll.LoadLibrary('libmulan_xla_ffi.so')
for _name in ('MulanFwdPre', 'MulanFwdPost', 'MulanBwdPost', 'MulanBwdPre'):
  jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(_so, _name)), platform='CUDA')

MULAN_FLAG_C_RAW = 1


def _attrs(cfg, param, c_raw):
  return dict(vocab=np.int32(cfg.vocab_size), param=np.int32(param),
              gt_mode=np.int32(0 if cfg.unet_type == 'vdm' else 1),
              n_timesteps=np.int32(cfg.sm_n_timesteps),     # every handler: T scales the loss
              flags=np.int32(MULAN_FLAG_C_RAW if c_raw else 0),
              gamma_min=np.float64(cfg.gamma_min), gamma_max=np.float64(cfg.gamma_max))


# ---- mulan_pre: everything before the denoiser ------------------------------------------------
@partial(jax.custom_vjp, nondiff_argnums=(0, 1, 2))
def mulan_pre(cfg, param, c_raw, x, a, b, c, t, eps0, eps):
  """-> z_t, g_net, loss_recon, loss_klz_prior, var_sums, w, link, link_net."""
  return _pre_fwd(cfg, param, c_raw, x, a, b, c, t, eps0, eps)[0]


def _pre_fwd(cfg, param, c_raw, x, a, b, c, t, eps0, eps):
  B, D = a.shape
  f32 = lambda *s: jax.ShapeDtypeStruct(s, jnp.float32)
  g_shape = (B,) if cfg.unet_type == 'vdm' else (B, D)
  z_t, g_net, w, rec, klz, var_sums = jax.ffi.ffi_call(
      'MulanFwdPre', (f32(B, D), f32(*g_shape), f32(B, D), f32(B), f32(B), f32(B, 2)))(
          x, a, b, c, t, eps0, eps, **_attrs(cfg, param, c_raw))
  link, link_net = jnp.zeros((B,), jnp.float32), jnp.zeros((B, D), jnp.float32)
  return (z_t, g_net, rec, klz, var_sums, w, link, link_net), (x, a, b, c, t, eps)


def _pre_bwd(cfg, param, c_raw, res, cts):
  x, a, b, c, t, eps = res
  # recon / prior KL: fixed ends, zero gradient.  gL and net arrive through the carriers.
  z_bar, g_bar, _, _, _, _, gL, net = cts
  B, D = a.shape
  f32 = jax.ShapeDtypeStruct((B, D), jnp.float32)
  a_bar, b_bar, c_bar = jax.ffi.ffi_call('MulanBwdPre', (f32, f32, f32))(
      x, a, b, c, t, eps, net, z_bar, g_bar, gL, **_attrs(cfg, param, c_raw))
  return (None, a_bar, b_bar, c_bar, None, None, None)


mulan_pre.defvjp(_pre_fwd, _pre_bwd)


# ---- mulan_post: the diffusion loss after the denoiser ----------------------------------------
@partial(jax.custom_vjp, nondiff_argnums=(0, 1, 2))
def _post(cfg, param, c_raw, x, a, b, c, t, eps, w, net, link, link_net):
  B = a.shape[0]
  return jax.ffi.ffi_call('MulanFwdPost', jax.ShapeDtypeStruct((B,), jnp.float32))(
      x, a, b, c, t, eps, net, w, **_attrs(cfg, param, c_raw))


def _post_fwd(cfg, param, c_raw, x, a, b, c, t, eps, w, net, link, link_net):
  return (_post(cfg, param, c_raw, x, a, b, c, t, eps, w, net, link, link_net),
          (x, a, b, c, t, eps, w, net))


def _post_bwd(cfg, param, c_raw, res, gL):
  x, a, b, c, t, eps, w, net = res
  B, D = a.shape
  n_bar = jax.ffi.ffi_call('MulanBwdPost', jax.ShapeDtypeStruct((B, D), jnp.float32))(
      x, a, b, c, t, eps, net, w, gL, **_attrs(cfg, param, c_raw))
  # cotangents: x a b c t eps w | net | link <- gL | link_net <- net (carried to mulan_pre's vjp)
  return (None, None, None, None, None, None, None, n_bar, gL, net)


_post.defvjp(_post_fwd, _post_bwd)


def mulan_post(cfg, param, c_raw, x, a, b, c, t, eps, w, net, link, link_net):
  """-> loss_diff[B].  (a, b, c) receive their whole cotangent from mulan_pre's vjp."""
  sg = jax.lax.stop_gradient
  return _post(cfg, param, c_raw, x, sg(a), sg(b), sg(c), t, eps, sg(w), net, link, link_net)
