"""The documented JAX binding's custom_vjp WIRING (jax_binding/mulan_jax.py), executed on a
torch-backed stand-in for the handful of jax APIs it uses.  JAX cannot be installed in this image,
so the file itself has never run under real JAX; what CAN be checked here is the part that is
this repository's own logic: which handler each forward / backward rule calls with which
operands and attributes, and that the carrier-cotangent trick (gL and the denoiser output travel
back to mulan_pre's vjp as the cotangents of two zero outputs, so ONE mulan_bwd_pre launch
produces the complete a_bar, b_bar, c_bar) yields the gradients of the tested PyTorch binding.

The shim maps  jax.custom_vjp -> torch.autograd.Function,  jax.ffi.ffi_call(name, ...) -> the same
C-ABI entry point the XLA-FFI handler of that name forwards to (jax_binding/mulan_xla_ffi.cc),
jax.lax.stop_gradient -> detach.  Real-JAX semantics the shim cannot vouch for (tracing, XLA's
dead-code elimination of the zero carriers) stay unverified and are labelled so in INTEGRATION.md.
"""
import ctypes
import importlib.util
import math
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import mulan_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAUNCHES = []


def _make_jax_shim():
  from mulan_b200 import ops
  jax = types.ModuleType('jax')
  jnp = types.ModuleType('jax.numpy')
  jnp.float32 = torch.float32
  jnp.zeros = lambda shape, dtype=torch.float32: torch.zeros(shape, dtype=dtype, device='cuda')
  jax.numpy = jnp

  class ShapeDtypeStruct:
    def __init__(self, shape, dtype):
      self.shape, self.dtype = tuple(shape), dtype
  jax.ShapeDtypeStruct = ShapeDtypeStruct
  lax = types.ModuleType('jax.lax')
  lax.stop_gradient = lambda v: v.detach()
  jax.lax = lax

  def desc_of(attrs):
    flags = int(attrs['flags'])
    return ops.Desc(vocab=int(attrs['vocab']), param=int(attrs['param']),
                    gt_mode=int(attrs['gt_mode']), n_timesteps=int(attrs['n_timesteps']),
                    gamma_min=float(attrs['gamma_min']), gamma_max=float(attrs['gamma_max']),
                    c_raw=bool(flags & 1), pdl=bool(flags & 2))

  # what each XLA-FFI handler of jax_binding/mulan_xla_ffi.cc forwards to
  def fwd_pre(x, a, b, c, t, eps0, eps, **at):
    r = ops.fwd_pre(desc_of(at), x, a, b, c, t, eps0, eps, save_w=True)
    return (r['z_t'], r['g_net'], r['w'], r['loss_recon'], r['loss_klz_prior'], r['var_sums'])

  def fwd_post(x, a, b, c, t, eps, net, w, **at):
    d = desc_of(at)
    return ops.fwd_post(d, x, a, b, c, t, eps, net, w if ops.saves_w(d) else None)

  def bwd_post(x, a, b, c, t, eps, net, w, gL, **at):
    d = desc_of(at)
    return ops.bwd_post(d, x, a, b, c, t, eps, net, w if ops.saves_w(d) else None, gL)

  def bwd_pre(x, a, b, c, t, eps, net, z_bar, g_bar, gL, **at):
    return ops.bwd_pre(desc_of(at), x, a, b, c, t, eps, net, z_bar, g_bar, gL)
  handlers = {'MulanFwdPre': fwd_pre, 'MulanFwdPost': fwd_post, 'MulanBwdPost': bwd_post,
              'MulanBwdPre': bwd_pre}
  ffi = types.ModuleType('jax.ffi')
  ffi.register_ffi_target = lambda *a, **k: None
  ffi.pycapsule = lambda fn: fn

  def ffi_call(name, out_types):
    def call(*args, **attrs):
      LAUNCHES.append(name)
      args = [v.contiguous() for v in args]
      out = handlers[name](*args, **attrs)
      want = out_types if isinstance(out_types, (tuple, list)) else (out_types,)
      got = out if isinstance(out, tuple) else (out,)
      assert len(got) == len(want)
      for g_, w_ in zip(got, want):
        assert tuple(g_.shape) == w_.shape, (name, g_.shape, w_.shape)
      return out
    return call
  ffi.ffi_call = ffi_call
  jax.ffi = ffi

  class custom_vjp:
    """jax.custom_vjp with nondiff_argnums, on torch autograd: cotangents of outputs nobody
    used arrive as zeros (JAX) instead of None (torch)."""

    def __init__(self, fun, nondiff_argnums=()):
      self.fun, self.nondiff = fun, tuple(nondiff_argnums)

    def defvjp(self, fwd, bwd):
      self.fwd, self.bwd = fwd, bwd

    def __call__(self, *args):
      if not torch.is_grad_enabled():
        # inside a forward rule (or with differentiation off) the wrapper evaluates the primal
        # body, as jax does when `f_fwd` calls `f` itself
        return self.fun(*args)
      outer = self
      nd = [args[i] for i in self.nondiff]
      diff_idx = [i for i in range(len(args)) if i not in self.nondiff]

      class Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, *dargs):
          full = list(args)
          for i, v in zip(diff_idx, dargs):
            full[i] = v
          with torch.no_grad():
            out, res = outer.fwd(*full)
          ctx.res = res
          ctx.is_tuple = isinstance(out, tuple)
          outs = out if ctx.is_tuple else (out,)
          ctx.protos = [(o.shape, o.dtype) for o in outs]
          return tuple(outs)

        @staticmethod
        def backward(ctx, *cts):
          cts = [torch.zeros(s, dtype=d, device='cuda') if c is None else c.contiguous()
                 for c, (s, d) in zip(cts, ctx.protos)]
          with torch.no_grad():
            grads = outer.bwd(*nd, ctx.res, tuple(cts) if ctx.is_tuple else cts[0])
          assert len(grads) == len(diff_idx)
          return tuple(grads)
      dargs = [args[i] for i in diff_idx]
      out = Fn.apply(*dargs)
      res = out if len(out) > 1 else out[0]
      return res
  jax.custom_vjp = lambda fun=None, nondiff_argnums=(): (
      custom_vjp(fun, nondiff_argnums) if fun is not None
      else (lambda f: custom_vjp(f, nondiff_argnums)))
  return jax, jnp


def _import_binding(monkeypatch):
  jax, jnp = _make_jax_shim()
  monkeypatch.setitem(sys.modules, 'jax', jax)
  monkeypatch.setitem(sys.modules, 'jax.numpy', jnp)

  class _Cdll:
    def LoadLibrary(self, name):
      return types.SimpleNamespace(MulanFwdPre=1, MulanFwdPost=2, MulanBwdPost=3, MulanBwdPre=4)
  monkeypatch.setattr(ctypes, 'cdll', _Cdll())
  spec = importlib.util.spec_from_file_location(
      'mulan_jax_under_shim', os.path.join(ROOT, 'jax_binding', 'mulan_jax.py'))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod


@pytest.mark.parametrize('param', [O.MODE_EPS, O.MODE_VEL, O.MODE_VEL_FROM_EPS])
@pytest.mark.parametrize('unet_type,c_raw', [('vdm', False), ('ldm', False), ('vdm', True)])
def test_jax_binding_wiring_matches_torch_binding(cuda_device, monkeypatch, param, unet_type,
                                                  c_raw):
  from mulan_b200 import ops
  mj = _import_binding(monkeypatch)
  dev = cuda_device
  B, D = 6, 3072
  inp = O.synth_inputs(B, 90 + param)
  g = {k: v.to(dev).contiguous() for k, v in inp.items()}
  if c_raw:
    g['c'] = torch.randn(B, D, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
  cfg = types.SimpleNamespace(vocab_size=256, unet_type=unet_type, sm_n_timesteps=0,
                              gamma_min=-13.3, gamma_max=5.0)
  pixel = unet_type != 'vdm'

  def score(z_t, g_net, w1, w2):
    gg = g_net if pixel else g_net.reshape(-1, 1)
    return w1 * z_t + w2 * gg + g['net']

  def leaves():
    mk = lambda v: v.clone().requires_grad_(True)
    return (mk(g['a']), mk(g['b']), mk(g['c']), torch.tensor(0.7, device=dev, requires_grad=True),
            torch.tensor(0.05, device=dev, requires_grad=True))
  scale = 1.0 / (D * math.log(2.0))
  # ---- the JAX binding's composition, as INTEGRATION.md patches VDM.__call__
  a, b, c, w1, w2 = leaves()
  LAUNCHES.clear()
  z_t, g_net, rec, klz, var_sums, w, link, link_net = mj.mulan_pre(
      cfg, param, c_raw, g['x'], a, b, c, g['t'], g['eps_0'], g['eps'])
  net = score(z_t, g_net, w1, w2)
  diff = mj.mulan_post(cfg, param, c_raw, g['x'], a, b, c, g['t'], g['eps'], w, net, link, link_net)
  bpd = (rec.mean() + klz.mean() + diff.mean()) * scale
  got = torch.autograd.grad(bpd, [a, b, c, w1, w2])
  assert LAUNCHES.count('MulanBwdPre') == 1, LAUNCHES      # ONE launch for every path
  assert LAUNCHES.count('MulanBwdPost') == 1 and LAUNCHES.count('MulanFwdPre') == 1
  # ---- the tested PyTorch binding on the same inputs
  a2, b2, c2, v1, v2 = leaves()
  desc = ops.Desc(param=param, gt_mode=1 if pixel else 0, c_raw=c_raw)
  tape = ops.ElboTape(desc)
  z2, g2, rec2, klz2, vs2, link2 = ops.mulan_pre(tape, g['x'], a2, b2, c2, g['t'], g['eps_0'],
                                                g['eps'])
  diff2 = ops.mulan_post(tape, score(z2, g2, v1, v2), link2)
  bpd2 = (rec2.mean() + klz2.mean() + diff2.mean()) * scale
  want = torch.autograd.grad(bpd2, [a2, b2, c2, v1, v2])
  assert abs(bpd.item() - bpd2.item()) < 1e-6 * abs(bpd2.item())
  for name, x_, y_ in zip('a b c w1 w2'.split(), got, want):
    err = ((x_ - y_).norm() / y_.norm().clamp_min(1e-30)).item()
    assert err < 1e-6, (name, err)
