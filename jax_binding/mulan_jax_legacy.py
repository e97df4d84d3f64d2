"""Binding for the jaxlib the reference pins (jax <= 0.4.23, README.md:26): XLA legacy GPU
custom calls (API_VERSION_STATUS_RETURNING) straight into libmulan_b200.so.

STATUS: this Python file is NOT executed in this repository's image (JAX is not installable
there).  The native side it binds -- the `mulan_xla_*` targets of include/mulan_b200_xla.h --
IS compiled into libmulan_b200.so and tested (tests/test_xla_custom_call.py drives each target
exactly as XLA does: stream, void** buffers = operands then results, opaque bytes).

Usage in the reference (ldm/model_mulan_epsilon.py, VDM.__call__, replacing :307-347; the same
in ldm/model_mulan_velocity.py:215-260): see INTEGRATION.md section 1 -- `mulan_pre` before the
`self.score_model(...)` call, `mulan_post` after it.  Both are jax.custom_vjp's, so
`jax.value_and_grad(loss_fn)` under `pmap` / `scan` (ldm/experiment.py:89-91, 339) works
unchanged; XLA calls each target on the executor thread of the device that owns the shard.
"""
import ctypes
import struct
from functools import partial

import jax
import jax.numpy as jnp
import numpy as np
from jax import core
from jax.interpreters import mlir
from jax.lib import xla_client
from jaxlib.hlo_helpers import custom_call

_so = ctypes.PyDLL('libmulan_b200.so')          # PyDLL: we only take addresses here
_capsule_new = ctypes.pythonapi.PyCapsule_New
_capsule_new.restype = ctypes.py_object
_capsule_new.argtypes = (ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p)

TARGETS = ('mulan_xla_fwd_pre', 'mulan_xla_fwd_post', 'mulan_xla_bwd_post',
           'mulan_xla_fwd_bwd_post', 'mulan_xla_bwd_pre', 'mulan_xla_bpd_reduce',
           'mulan_xla_aux_topk_fwd', 'mulan_xla_aux_topk_bwd')
for _name in TARGETS:
  _addr = ctypes.cast(getattr(_so, _name), ctypes.c_void_p).value
  xla_client.register_custom_call_target(
      _name.encode(), _capsule_new(_addr, b'xla._CUSTOM_CALL_TARGET', None), platform='CUDA')


def _opaque(cfg, param, rows, dim, absent_mask=0, flags=0):
  """bytes of mulan_xla_opaque {mulan_desc (ABI v2: ..., uint32 flags, int32 noise_rows);
  uint32 absent_mask; uint32 reserved}."""
  return struct.pack('<6i2dIi2I', rows, dim, cfg.vocab_size, param,
                     0 if cfg.unet_type == 'vdm' else 1, cfg.sm_n_timesteps,
                     cfg.gamma_min, cfg.gamma_max, flags, 0, absent_mask, 0)


def _row_major(aval):
  return tuple(range(len(aval.shape) - 1, -1, -1))


def _primitive(target, out_avals_fn):
  """One jax primitive per target: abstract eval from shapes, CUDA lowering = custom_call."""
  prim = core.Primitive(target)
  prim.multiple_results = True
  prim.def_impl(partial(jax.interpreters.xla.apply_primitive, prim))
  prim.def_abstract_eval(lambda *avals, **kw: out_avals_fn(*avals, **kw))

  def lowering(ctx, *operands, opaque, **_):
    return custom_call(
        target.encode(),
        result_types=[mlir.aval_to_ir_type(a) for a in ctx.avals_out],
        operands=operands,
        backend_config=opaque,
        operand_layouts=[_row_major(a) for a in ctx.avals_in],
        result_layouts=[_row_major(a) for a in ctx.avals_out],
        api_version=2).results      # 2 = API_VERSION_STATUS_RETURNING
  mlir.register_lowering(prim, lowering, platform='cuda')
  return prim


def _f32(*shape):
  return core.ShapedArray(shape, jnp.float32)


_fwd_pre_p = _primitive(
    'mulan_xla_fwd_pre',
    lambda x, a, *rest, opaque, pixel_gt: (
        _f32(*a.shape), _f32(*a.shape) if pixel_gt else _f32(a.shape[0]), _f32(*a.shape),
        _f32(a.shape[0]), _f32(a.shape[0]), _f32(a.shape[0], 2)))
_fwd_post_p = _primitive('mulan_xla_fwd_post', lambda x, a, *rest, opaque: (_f32(a.shape[0]),))
_bwd_post_p = _primitive('mulan_xla_bwd_post', lambda x, a, *rest, opaque: (_f32(*a.shape),))
_bwd_pre_p = _primitive('mulan_xla_bwd_pre',
                        lambda x, a, *rest, opaque: (_f32(*a.shape),) * 3)


# ONE mulan_bwd_pre launch per backward pass (as mulan_jax.py / the tested PyTorch binding): the
# loss cotangent gL and the denoiser output `net` travel back to mulan_pre's vjp as the
# cotangents of two zero carrier outputs (link [B], link_net [B, D]) that only mulan_post
# consumes; a, b, c enter mulan_post under stop_gradient.
MULAN_FLAG_C_RAW = 1


# ---- mulan_pre: everything before the denoiser (model_mulan_epsilon.py:300-328, 339-343) ----
@partial(jax.custom_vjp, nondiff_argnums=(0, 1, 2))
def mulan_pre(cfg, param, c_raw, x, a, b, c, t, eps0, eps):
  return _pre_fwd(cfg, param, c_raw, x, a, b, c, t, eps0, eps)[0]


def _flags(c_raw):
  return MULAN_FLAG_C_RAW if c_raw else 0


def _pre_fwd(cfg, param, c_raw, x, a, b, c, t, eps0, eps):
  B, D = a.shape
  z_t, g_net, w, rec, klz, var_sums = _fwd_pre_p.bind(
      x, a, b, c, t, eps0, eps, opaque=_opaque(cfg, param, B, D, flags=_flags(c_raw)),
      pixel_gt=cfg.unet_type != 'vdm')
  link, link_net = jnp.zeros((B,), jnp.float32), jnp.zeros((B, D), jnp.float32)
  return (z_t, g_net, rec, klz, var_sums, w, link, link_net), (x, a, b, c, t, eps)


def _pre_bwd(cfg, param, c_raw, res, cts):
  x, a, b, c, t, eps = res
  # recon / prior KL: fixed ends, zero (a,b,c) gradient; gL and net arrive through the carriers
  z_bar, g_bar, gL, net = cts[0], cts[1], cts[6], cts[7]
  B, D = a.shape
  # buffers: x a b c t eps net z_bar g_bar gL : nothing absent, ONE launch for every path
  a_bar, b_bar, c_bar = _bwd_pre_p.bind(
      x, a, b, c, t, eps, net, z_bar, g_bar, gL,
      opaque=_opaque(cfg, param, B, D, flags=_flags(c_raw)))
  return (None, a_bar, b_bar, c_bar, None, None, None)


mulan_pre.defvjp(_pre_fwd, _pre_bwd)


# ---- mulan_post: the diffusion loss after the denoiser (model_mulan_epsilon.py:345-355,
#      model_mulan_velocity.py:243-260) ----
@partial(jax.custom_vjp, nondiff_argnums=(0, 1, 2))
def _post(cfg, param, c_raw, x, a, b, c, t, eps, w, net, link, link_net):
  B, D = a.shape
  return _fwd_post_p.bind(x, a, b, c, t, eps, net, w,
                          opaque=_opaque(cfg, param, B, D, flags=_flags(c_raw)))[0]


def _post_fwd(cfg, param, c_raw, x, a, b, c, t, eps, w, net, link, link_net):
  return (_post(cfg, param, c_raw, x, a, b, c, t, eps, w, net, link, link_net),
          (x, a, b, c, t, eps, w, net))


def _post_bwd(cfg, param, c_raw, res, gL):
  x, a, b, c, t, eps, w, net = res
  B, D = a.shape
  n_bar, = _bwd_post_p.bind(x, a, b, c, t, eps, net, w, gL,
                            opaque=_opaque(cfg, param, B, D, flags=_flags(c_raw)))
  # x a b c t eps w | net | link <- gL | link_net <- net (carried to mulan_pre's vjp)
  return (None, None, None, None, None, None, None, n_bar, gL, net)


_post.defvjp(_post_fwd, _post_bwd)


def mulan_post(cfg, param, c_raw, x, a, b, c, t, eps, w, net, link, link_net):
  sg = jax.lax.stop_gradient
  return _post(cfg, param, c_raw, x, sg(a), sg(b), sg(c), t, eps, sg(w), net, link, link_net)
