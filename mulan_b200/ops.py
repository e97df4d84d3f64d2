"""PyTorch bindings of the C ABI: raw launches + the autograd pair used by the model.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); every number is
produced by libmulan_b200.so.  All ops require contiguous CUDA tensors and raise otherwise:
there is no CPU path.

Autograd structure (mirrors what ``jax.value_and_grad`` does around the reference's
``VDM.__call__``, ldm/experiment.py:339): ``mulan_pre`` runs before the denoiser,
``mulan_post`` after it.  ``mulan_post.backward`` only produces the denoiser cotangent
``n_bar``; the loss cotangent ``gL`` travels back to ``mulan_pre.backward`` through a
``link`` tensor, so that ONE kernel (``mulan_bwd_pre``) produces the complete
``a_bar, b_bar, c_bar`` (paths through z_t, through the denoiser's noise-level input and
through loss_diff) instead of two kernels plus an add.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import _lib
from ._lib import (MULAN_FLAG_C_RAW, MULAN_FLAG_PDL, MULAN_GT_MEAN, MULAN_GT_PIXEL,
                   MULAN_PARAM_EPS, MULAN_PARAM_VEL, MULAN_PARAM_VEL_FROM_EPS, make_desc)


def _p(t: Optional[torch.Tensor]):
  return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
  return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t: torch.Tensor, dtype, shape, name: str) -> torch.Tensor:
  if not isinstance(t, torch.Tensor) or not t.is_cuda:
    raise TypeError(f'{name}: expected a CUDA tensor (libmulan_b200 has no CPU path)')
  if t.dtype != dtype:
    raise TypeError(f'{name}: expected dtype {dtype}, got {t.dtype}')
  if tuple(t.shape) != tuple(shape):
    raise ValueError(f'{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}')
  if not t.is_contiguous():
    raise ValueError(f'{name}: must be contiguous')
  return t


def _opt(t, dtype, shape, name):
  return None if t is None else _req(t, dtype, shape, name)


class Desc:
  """Python view of mulan_desc (per launch; rows filled from the tensors).

  c_raw: the `c` tensors are the PRE-ACTIVATION of dense_out_c; the kernels apply
  1e-3 + softplus (ldm/model_mulan_epsilon.py:537) and bwd_pre returns its cotangent.
  pdl: programmatic dependent launch.  noise_rows: eps0 / eps are [noise_rows, D], broadcast
  over the batch by row % noise_rows (dense-VLB evaluation)."""

  def __init__(self, dim=3072, vocab=256, param=MULAN_PARAM_EPS, gt_mode=MULAN_GT_MEAN,
               n_timesteps=0, gamma_min=-13.3, gamma_max=5.0, c_raw=False, pdl=False,
               noise_rows=0):
    self.dim, self.vocab, self.param, self.gt_mode = dim, vocab, param, gt_mode
    self.n_timesteps, self.gamma_min, self.gamma_max = n_timesteps, gamma_min, gamma_max
    self.c_raw, self.pdl, self.noise_rows = bool(c_raw), bool(pdl), int(noise_rows)

  @property
  def flags(self) -> int:
    return (MULAN_FLAG_C_RAW if self.c_raw else 0) | (MULAN_FLAG_PDL if self.pdl else 0)

  def replace(self, **kw) -> 'Desc':
    d = Desc.__new__(Desc)
    d.__dict__.update(self.__dict__)
    d.__dict__.update(kw)
    return d

  def c(self, rows: int):
    return make_desc(rows, self.dim, self.vocab, self.param, self.gt_mode, self.n_timesteps,
                     self.gamma_min, self.gamma_max, self.flags, self.noise_rows)

  def noise_shape(self, rows: int):
    return (self.noise_rows if self.noise_rows > 0 else rows, self.dim)


def reduce_workspace(rows: int, device) -> torch.Tensor:
  """Zeroed scratch of the fixed-order loss-scalar reduction (mulan_reduce_ws_bytes); every
  call leaves it zero, so one buffer serves consecutive calls on a stream."""
  n = int(_lib.load().mulan_reduce_ws_bytes(int(rows)))
  return torch.zeros((n + 3) // 4, dtype=torch.int32, device=device)


def saves_w(desc: 'Desc') -> bool:
  """Whether mulan_fwd_pre should save the loss weight w for the post kernels of this model."""
  return _lib.kernel_param(desc.param) == MULAN_PARAM_EPS


# ------------------------------------------------------------------------------------
# Raw launches (no autograd)
# ------------------------------------------------------------------------------------

def fwd_pre(desc: Desc, x, a, b, c, t, eps0, eps, save_w: bool = True, end_consts=None):
  """mulan_fwd_pre. Returns dict(z_t, g_net, w, loss_recon, loss_klz_prior, var_sums).
  end_consts (_lib.MulanEndConsts, optional): the fixed-end constants as the caller's framework
  rounds them (mulan_fwd_pre_consts) instead of the host-computed, correctly rounded ones."""
  B, D = a.shape
  _req(x, torch.uint8, (B, D), 'x')
  for n, v in (('a', a), ('b', b), ('c', c)):
    _req(v, torch.float32, (B, D), n)
  for n, v in (('eps0', eps0), ('eps', eps)):
    _req(v, torch.float32, desc.noise_shape(B), n)
  _req(t, torch.float32, (B,), 't')
  dev = a.device
  z_t = torch.empty((B, D), dtype=torch.float32, device=dev)
  g_net = torch.empty((B,) if desc.gt_mode == MULAN_GT_MEAN else (B, D),
                      dtype=torch.float32, device=dev)
  w = torch.empty((B, D), dtype=torch.float32, device=dev) if save_w else None
  rec = torch.empty((B,), dtype=torch.float32, device=dev)
  klz = torch.empty((B,), dtype=torch.float32, device=dev)
  vs = torch.empty((B, 2), dtype=torch.float32, device=dev)
  d = desc.c(B)
  args = (_p(x), _p(a), _p(b), _p(c), _p(t), _p(eps0), _p(eps),
          _p(z_t), _p(g_net), _p(w), _p(rec), _p(klz), _p(vs), _stream())
  if end_consts is None:
    _lib.check(_lib.load().mulan_fwd_pre(C.byref(d), *args))
  else:
    _lib.check(_lib.load().mulan_fwd_pre_consts(C.byref(d), C.byref(end_consts), *args))
  return dict(z_t=z_t, g_net=g_net, w=w, loss_recon=rec, loss_klz_prior=klz, var_sums=vs)


def fwd_pre_keyed(desc: Desc, key_eps0, key_eps, x, a, b, c, t, save_w: bool = True,
                  want_eps: bool = True, want_eps0: bool = False):
  """mulan_fwd_pre_keyed: eps_0 / eps drawn inside the kernel from their (k0, k1) threefry keys.
  Returns fwd_pre's dict + eps (and eps_0) when asked for."""
  B, D = a.shape
  _req(x, torch.uint8, (B, D), 'x')
  for n, v in (('a', a), ('b', b), ('c', c)):
    _req(v, torch.float32, (B, D), n)
  _req(t, torch.float32, (B,), 't')
  dev = a.device
  f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
  z_t = f(B, D)
  g_net = f(B) if desc.gt_mode == MULAN_GT_MEAN else f(B, D)
  w = f(B, D) if save_w else None
  eps = f(B, D) if want_eps else None
  eps0 = f(B, D) if want_eps0 else None
  rec, klz, vs = f(B), f(B), f(B, 2)
  k0 = (C.c_uint32 * 2)(int(key_eps0[0]), int(key_eps0[1]))
  k1 = (C.c_uint32 * 2)(int(key_eps[0]), int(key_eps[1]))
  d = desc.c(B)
  _lib.check(_lib.load().mulan_fwd_pre_keyed(
      C.byref(d), k0, k1, _p(x), _p(a), _p(b), _p(c), _p(t), _p(z_t), _p(g_net), _p(w), _p(eps0),
      _p(eps), _p(rec), _p(klz), _p(vs), _stream()))
  return dict(z_t=z_t, g_net=g_net, w=w, loss_recon=rec, loss_klz_prior=klz, var_sums=vs, eps=eps,
              eps_0=eps0)


def fwd_post(desc: Desc, x, a, b, c, t, eps, net, w=None):
  """mulan_fwd_post -> loss_diff[B]."""
  B, D = net.shape
  _req(net, torch.float32, (B, D), 'net')
  _req(eps, torch.float32, desc.noise_shape(B), 'eps')
  out = torch.empty((B,), dtype=torch.float32, device=eps.device)
  d = desc.c(B)
  _lib.check(_lib.load().mulan_fwd_post(
      C.byref(d), _p(x), _p(a), _p(b), _p(c), _p(t), _p(eps), _p(net), _p(w), _p(out),
      _stream()))
  return out


def bwd_post(desc: Desc, x, a, b, c, t, eps, net, w, gL):
  """mulan_bwd_post -> n_bar[B,D]."""
  B, D = net.shape
  _req(gL, torch.float32, (B,), 'gL')
  out = torch.empty((B, D), dtype=torch.float32, device=eps.device)
  d = desc.c(B)
  _lib.check(_lib.load().mulan_bwd_post(
      C.byref(d), _p(x), _p(a), _p(b), _p(c), _p(t), _p(eps), _p(net), _p(w), _p(gL), _p(out),
      _stream()))
  return out


def fwd_bwd_post(desc: Desc, x, a, b, c, t, eps, net, w, gL):
  """mulan_fwd_bwd_post -> (loss_diff[B], n_bar[B,D]) in one pass, for a known cotangent gL."""
  B, D = net.shape
  _req(net, torch.float32, (B, D), 'net')
  _req(eps, torch.float32, desc.noise_shape(B), 'eps')
  _req(gL, torch.float32, (B,), 'gL')
  diff = torch.empty((B,), dtype=torch.float32, device=eps.device)
  n_bar = torch.empty((B, D), dtype=torch.float32, device=eps.device)
  d = desc.c(B)
  _lib.check(_lib.load().mulan_fwd_bwd_post(
      C.byref(d), _p(x), _p(a), _p(b), _p(c), _p(t), _p(eps), _p(net), _p(w), _p(gL), _p(diff),
      _p(n_bar), _stream()))
  return diff, n_bar


def scale_rows(v, num, den):
  """mulan_scale_rows: v[b,:] *= num[b]/den[b] in place for rows where they differ."""
  B, D = v.shape
  _req(v, torch.float32, (B, D), 'v')
  _req(num, torch.float32, (B,), 'num')
  _req(den, torch.float32, (B,), 'den')
  _lib.check(_lib.load().mulan_scale_rows(B, D, _p(v), _p(num), _p(den), _stream()))
  return v


def bwd_pre(desc: Desc, x, a, b, c, t, eps, net, z_bar, g_bar, gL):
  """mulan_bwd_pre -> (a_bar, b_bar, c_bar)."""
  B, D = a.shape
  _opt(z_bar, torch.float32, (B, D), 'z_bar')
  _opt(g_bar, torch.float32, (B,) if desc.gt_mode == MULAN_GT_MEAN else (B, D), 'g_bar')
  _opt(gL, torch.float32, (B,), 'gL')
  ab, bb, cb = (torch.empty((B, D), dtype=torch.float32, device=a.device) for _ in range(3))
  d = desc.c(B)
  _lib.check(_lib.load().mulan_bwd_pre(
      C.byref(d), _p(x), _p(a), _p(b), _p(c), _p(t), _p(eps), _p(net), _p(z_bar), _p(g_bar),
      _p(gL), _p(ab), _p(bb), _p(cb), _stream()))
  return ab, bb, cb


def aux_topk_fwd(logits, gamma_draw, k: int):
  """mulan_aux_topk_fwd -> (embedding[B,L], kl_z[B])."""
  B, L = logits.shape
  _req(logits, torch.float32, (B, L), 'logits')
  _opt(gamma_draw, torch.float32, (10, B, L), 'gamma_draw')
  emb = torch.empty((B, L), dtype=torch.float32, device=logits.device)
  kl = torch.empty((B,), dtype=torch.float32, device=logits.device)
  _lib.check(_lib.load().mulan_aux_topk_fwd(B, L, k, _p(logits), _p(gamma_draw), _p(emb),
                                            _p(kl), _stream()))
  return emb, kl


def aux_topk_bwd(logits, gamma_draw, k: int, emb_bar, klz_bar):
  """mulan_aux_topk_bwd -> logits_bar[B,L]."""
  B, L = logits.shape
  _opt(emb_bar, torch.float32, (B, L), 'emb_bar')
  _opt(klz_bar, torch.float32, (B,), 'klz_bar')
  out = torch.empty((B, L), dtype=torch.float32, device=logits.device)
  _lib.check(_lib.load().mulan_aux_topk_bwd(B, L, k, _p(logits), _p(gamma_draw), _p(emb_bar),
                                            _p(klz_bar), _p(out), _stream()))
  return out


def bpd_reduce(desc: Desc, loss_recon, loss_klz_prior, kl_z, loss_diff, var_sums,
               want_klz_total: bool = False, ws='auto'):
  """mulan_bpd_reduce -> scalars[6] = bpd, bpd_latent, bpd_recon, bpd_diff, var0, var1.
  ws: reduce_workspace(rows) tensor, 'auto' (allocate one) or None (single-CTA form)."""
  B = loss_recon.shape[0]
  sc = torch.empty((6,), dtype=torch.float32, device=loss_recon.device)
  tot = torch.empty((B,), dtype=torch.float32, device=loss_recon.device) if want_klz_total else None
  if isinstance(ws, str):
    ws = reduce_workspace(B, loss_recon.device)
  d = desc.c(B)
  _lib.check(_lib.load().mulan_bpd_reduce(
      C.byref(d), _p(loss_recon), _p(loss_klz_prior), _p(kl_z), _p(loss_diff), _p(var_sums),
      _p(sc), _p(tot), _p(ws), _stream()))
  return (sc, tot) if want_klz_total else sc


def post_bpd(desc: Desc, x, a, b, c, t, eps, net, w, gL, loss_recon, loss_klz_prior, kl_z,
             var_sums, want_klz_total: bool = False, ws=None):
  """mulan_post_bpd: the post kernel (value only when gL is None, value-and-grad otherwise) with
  the loss-scalar reduction fused into its epilogue.
  -> dict(loss_diff[B], n_bar[B,D] | None, scalars[6], loss_klz_total[B] | None)."""
  B, D = net.shape
  _req(net, torch.float32, (B, D), 'net')
  _req(eps, torch.float32, desc.noise_shape(B), 'eps')
  _opt(gL, torch.float32, (B,), 'gL')
  dev = net.device
  diff = torch.empty((B,), dtype=torch.float32, device=dev)
  n_bar = torch.empty((B, D), dtype=torch.float32, device=dev) if gL is not None else None
  sc = torch.empty((6,), dtype=torch.float32, device=dev)
  tot = torch.empty((B,), dtype=torch.float32, device=dev) if want_klz_total else None
  if ws is None:
    ws = reduce_workspace(B, dev)
  d = desc.c(B)
  _lib.check(_lib.load().mulan_post_bpd(
      C.byref(d), _p(x), _p(a), _p(b), _p(c), _p(t), _p(eps), _p(net), _p(w), _p(gL),
      _p(loss_recon), _p(loss_klz_prior), _p(kl_z), _p(var_sums), _p(diff), _p(n_bar), _p(sc),
      _p(tot), _p(ws), _stream()))
  return dict(loss_diff=diff, n_bar=n_bar, scalars=sc, loss_klz_total=tot)


def sample_gamma(desc: Desc, a, b, c, t):
  """mulan_sample_gamma -> the denoiser's noise-level input for a sampler step ([B] or [B,D]).
  a, b, c: [B,D] or [1,D] (one coefficient row broadcast over the batch)."""
  B = t.shape[0]
  D = a.shape[1]
  rows_abc = a.shape[0]
  for n, v in (('a', a), ('b', b), ('c', c)):
    _req(v, torch.float32, (rows_abc, D), n)
  _req(t, torch.float32, (B,), 't')
  out = torch.empty((B,) if desc.gt_mode == MULAN_GT_MEAN else (B, D), dtype=torch.float32,
                    device=a.device)
  d = desc.c(B)
  _lib.check(_lib.load().mulan_sample_gamma(C.byref(d), rows_abc, _p(a), _p(b), _p(c), _p(t),
                                            _p(out), _stream()))
  return out


def sample_step(desc: Desc, a, b, c, t, s, z_t, net, eps, out=None):
  """mulan_sample_step -> z_s[B,D] (one ancestral step; VDM.sample after the denoiser call)."""
  B, D = z_t.shape
  rows_abc = a.shape[0]
  for n, v in (('a', a), ('b', b), ('c', c)):
    _req(v, torch.float32, (rows_abc, D), n)
  for n, v in (('z_t', z_t), ('net', net), ('eps', eps)):
    _req(v, torch.float32, (B, D), n)
  _req(t, torch.float32, (B,), 't')
  _req(s, torch.float32, (B,), 's')
  if out is None:
    out = torch.empty((B, D), dtype=torch.float32, device=z_t.device)
  d = desc.c(B)
  _lib.check(_lib.load().mulan_sample_step(C.byref(d), rows_abc, _p(a), _p(b), _p(c), _p(t),
                                           _p(s), _p(z_t), _p(net), _p(eps), _p(out), _stream()))
  return out


def generate_x(desc: Desc, z_0):
  """mulan_generate_x -> x[B,D] uint8 (argmax decode of z_0; VDM.generate_x)."""
  B, D = z_0.shape
  _req(z_0, torch.float32, (B, D), 'z_0')
  out = torch.empty((B, D), dtype=torch.uint8, device=z_0.device)
  d = desc.c(B)
  _lib.check(_lib.load().mulan_generate_x(C.byref(d), _p(z_0), _p(out), _stream()))
  return out


def ode_drift(desc: Desc, a, b, c, t, x_t, eps_hat, v=None, high_precision: bool = False,
              out=None):
  """mulan_ode_drift -> drift[B,D] (v is None) or (drift, net_bar[B,D], div_direct[B]);
  `out` (optional) receives the drift."""
  B, D = x_t.shape
  rows_abc = a.shape[0]
  for n, q in (('a', a), ('b', b), ('c', c)):
    _req(q, torch.float32, (rows_abc, D), n)
  for n, q in (('x_t', x_t), ('eps_hat', eps_hat)):
    _req(q, torch.float32, (B, D), n)
  _opt(v, torch.float32, (B, D), 'v')
  _req(t, torch.float32, (B,), 't')
  drift = (torch.empty((B, D), dtype=torch.float32, device=x_t.device) if out is None
           else _req(out, torch.float32, (B, D), 'out'))
  nb = torch.empty_like(drift) if v is not None else None
  dd = torch.empty((B,), dtype=torch.float32, device=x_t.device) if v is not None else None
  d = desc.c(B)
  _lib.check(_lib.load().mulan_ode_drift(
      C.byref(d), rows_abc, _p(a), _p(b), _p(c), _p(t), _p(x_t), _p(eps_hat), _p(v),
      1 if high_precision else 0, _p(drift), _p(nb), _p(dd), _stream()))
  return drift if v is None else (drift, nb, dd)


def row_dot(u, v, add=None, out=None):
  """mulan_row_dot -> out[b] = <u[b], v[b]> (+ add[b])."""
  B, D = u.shape
  _req(u, torch.float32, (B, D), 'u')
  _req(v, torch.float32, (B, D), 'v')
  _opt(add, torch.float32, (B,), 'add')
  out = (torch.empty((B,), dtype=torch.float32, device=u.device) if out is None
         else _req(out, torch.float32, (B,), 'out'))
  _lib.check(_lib.load().mulan_row_dot(B, D, _p(u), _p(v), _p(add), _p(out), _stream()))
  return out


def _rk45_K(K, n_k: int, n: int):
  """K: [>= n_k, k_stride >= n] float32 stage derivatives (rows may be padded for alignment)."""
  if K.dim() != 2 or K.shape[0] < n_k or K.shape[1] < n:
    raise ValueError(f'K: expected [>= {n_k}, >= {n}], got {tuple(K.shape)}')
  _req(K, torch.float32, tuple(K.shape), 'K')


def rk45_stage(n_k: int, coef, h: float, y, K, y_stage=None, y_out=None):
  """mulan_rk45_stage: v = y + (sum_{j<n_k} coef[j] K[j]) h -> y_stage (float32) / y_out (float64)."""
  n = y.numel()
  _req(y, torch.float64, (n,), 'y')
  _rk45_K(K, n_k, n)
  if y_stage is not None:
    _req(y_stage, torch.float32, (n,), 'y_stage')
  if y_out is not None:
    _req(y_out, torch.float64, (n,), 'y_out')
  cf = (C.c_double * 7)(*[float(v) for v in coef[:n_k]])
  _lib.check(_lib.load().mulan_rk45_stage(n, n_k, cf, float(h), _p(y), _p(K), K.shape[1], _p(y_stage),
                                          _p(y_out), _stream()))


def rk45_norm(n_k: int, coef, h: float, rtol: float, atol: float, y, y_new, K, of_y: bool,
              scratch, out):
  """mulan_rk45_norm: out[0] = sum ((y | (sum coef K) h) / (atol + rtol max(|y|,|y_new|)))^2."""
  n = y.numel()
  _req(y, torch.float64, (n,), 'y')
  if y_new is not None:
    _req(y_new, torch.float64, (n,), 'y_new')
  _rk45_K(K, n_k, n)
  _req(scratch, torch.float64, (_lib.MULAN_RK45_SCRATCH,), 'scratch')
  _req(out, torch.float64, (1,), 'out')
  cf = (C.c_double * 7)(*[float(v) for v in coef[:n_k]])
  _lib.check(_lib.load().mulan_rk45_norm(n, n_k, cf, float(h), float(rtol), float(atol), _p(y),
                                         _p(y_new), _p(K), K.shape[1], 1 if of_y else 0, _p(scratch),
                                         _p(out), _stream()))


class ElboWorkspace:
  """Preallocated outputs of the whole path for a fixed shard size: every launch writes into
  the same buffers, so a step makes no allocation and can be captured in a CUDA graph
  (the reference's `pmap(scan(train_step))` keeps the 1000 sub-steps on the device the same
  way, ldm/experiment.py:89-91)."""

  def __init__(self, desc: Desc, rows: int, device, save_w: Optional[bool] = None):
    self.desc, self.rows = desc, rows
    self.save_w = saves_w(desc) if save_w is None else save_w
    D = desc.dim
    f = lambda *s: torch.empty(s, dtype=torch.float32, device=device)
    self.z_t = f(rows, D)
    self.g_net = f(rows) if desc.gt_mode == MULAN_GT_MEAN else f(rows, D)
    self.w = f(rows, D) if self.save_w else None
    self.loss_recon, self.loss_klz_prior, self.loss_diff = f(rows), f(rows), f(rows)
    self.var_sums, self.scalars, self.loss_klz = f(rows, 2), f(6), f(rows)
    self.n_bar, self.a_bar, self.b_bar, self.c_bar = f(rows, D), f(rows, D), f(rows, D), f(rows, D)
    self.reduce_ws = reduce_workspace(rows, device)
    self._d = desc.c(rows)
    self._lib = _lib.load()

  def fwd_pre(self, x, a, b, c, t, eps0, eps):
    _lib.check(self._lib.mulan_fwd_pre(
        C.byref(self._d), _p(x), _p(a), _p(b), _p(c), _p(t), _p(eps0), _p(eps), _p(self.z_t),
        _p(self.g_net), _p(self.w), _p(self.loss_recon), _p(self.loss_klz_prior),
        _p(self.var_sums), _stream()))

  def fwd_post(self, x, a, b, c, t, eps, net):
    _lib.check(self._lib.mulan_fwd_post(
        C.byref(self._d), _p(x), _p(a), _p(b), _p(c), _p(t), _p(eps), _p(net), _p(self.w),
        _p(self.loss_diff), _stream()))

  def fwd_bwd_post(self, x, a, b, c, t, eps, net, gL):
    _lib.check(self._lib.mulan_fwd_bwd_post(
        C.byref(self._d), _p(x), _p(a), _p(b), _p(c), _p(t), _p(eps), _p(net), _p(self.w), _p(gL),
        _p(self.loss_diff), _p(self.n_bar), _stream()))

  def bpd_reduce(self, kl_z=None, parallel: bool = True):
    _lib.check(self._lib.mulan_bpd_reduce(
        C.byref(self._d), _p(self.loss_recon), _p(self.loss_klz_prior), _p(kl_z),
        _p(self.loss_diff), _p(self.var_sums), _p(self.scalars), _p(self.loss_klz),
        _p(self.reduce_ws if parallel else None), _stream()))

  def post_bpd(self, x, a, b, c, t, eps, net, gL=None, kl_z=None, board=None):
    """post kernel (value-and-grad when gL is given) + the six scalars, one launch.  board
    (peer.ScalarBoard): also publish the scalars to every rank over NVLink peer memory."""
    _lib.check(self._lib.mulan_post_bpd_peer(
        C.byref(self._d), _p(x), _p(a), _p(b), _p(c), _p(t), _p(eps), _p(net), _p(self.w), _p(gL),
        _p(self.loss_recon), _p(self.loss_klz_prior), _p(kl_z), _p(self.var_sums),
        _p(self.loss_diff), _p(self.n_bar if gL is not None else None), _p(self.scalars),
        _p(self.loss_klz), _p(self.reduce_ws), board.byref() if board is not None else None,
        _stream()))

  def bwd_post(self, x, a, b, c, t, eps, net, gL):
    _lib.check(self._lib.mulan_bwd_post(
        C.byref(self._d), _p(x), _p(a), _p(b), _p(c), _p(t), _p(eps), _p(net), _p(self.w), _p(gL),
        _p(self.n_bar), _stream()))

  def bwd_pre(self, x, a, b, c, t, eps, net, z_bar, g_bar, gL):
    _lib.check(self._lib.mulan_bwd_pre(
        C.byref(self._d), _p(x), _p(a), _p(b), _p(c), _p(t), _p(eps), _p(net), _p(z_bar),
        _p(g_bar), _p(gL), _p(self.a_bar), _p(self.b_bar), _p(self.c_bar), _stream()))


# ------------------------------------------------------------------------------------
# Autograd pair
# ------------------------------------------------------------------------------------

class ElboTape:
  """Residuals shared by mulan_pre / mulan_post of ONE forward pass."""

  def __init__(self, desc: Desc, save_w: Optional[bool] = None):
    self.desc = desc
    # Saving w pays for the epsilon form (post kernels read 12 B instead of 20 B) -- which
    # velocity_from_epsilon also runs (mulan_kernel_param); the plain velocity model
    # recomputes gamma_t anyway.
    self.save_w = saves_w(desc) if save_w is None else save_w
    self.x = self.a = self.b = self.c = self.t = self.eps = self.w = self.net = None


class _MulanPre(torch.autograd.Function):

  @staticmethod
  def forward(ctx, tape: ElboTape, x, a, b, c, t, eps0, eps):
    a, b, c = a.contiguous(), b.contiguous(), c.contiguous()
    out = fwd_pre(tape.desc, x, a, b, c, t, eps0, eps, save_w=tape.save_w)
    tape.x, tape.a, tape.b, tape.c, tape.t, tape.eps, tape.w = x, a, b, c, t, eps, out['w']
    ctx.tape = tape
    link = torch.zeros_like(t)
    ctx.mark_non_differentiable(out['loss_recon'], out['loss_klz_prior'], out['var_sums'])
    return (out['z_t'], out['g_net'], out['loss_recon'], out['loss_klz_prior'],
            out['var_sums'], link)

  @staticmethod
  def backward(ctx, z_bar, g_bar, _r, _k, _v, link_bar):
    tp = ctx.tape
    gL = link_bar
    if gL is not None and tp.net is None:
      gL = None
    cont = lambda v: None if v is None else v.contiguous()
    ab, bb, cb = bwd_pre(tp.desc, tp.x, tp.a, tp.b, tp.c, tp.t, tp.eps, tp.net,
                         cont(z_bar), cont(g_bar), cont(gL))
    return None, None, ab, bb, cb, None, None, None


class _MulanPost(torch.autograd.Function):

  @staticmethod
  def forward(ctx, tape: ElboTape, net, link):
    net_c = net.contiguous()
    tape.net = net_c.detach()
    ctx.tape = tape
    return fwd_post(tape.desc, tape.x, tape.a, tape.b, tape.c, tape.t, tape.eps, net_c, tape.w)

  @staticmethod
  def backward(ctx, gL):
    tp = ctx.tape
    gL = gL.contiguous()
    n_bar = bwd_post(tp.desc, tp.x, tp.a, tp.b, tp.c, tp.t, tp.eps, tp.net, tp.w, gL)
    return None, n_bar, gL


class _MulanPostFused(torch.autograd.Function):
  """loss_diff AND the denoiser cotangent in the forward pass, for the cotangent `gL_hint` the
  caller expects (the mean-of-losses scaling of loss_fn).  backward() hands out the
  precomputed n_bar; a row whose true cotangent differs from the hint is rescaled in place
  by mulan_scale_rows (no host sync, no traffic for matching rows)."""

  @staticmethod
  def forward(ctx, tape: ElboTape, net, link, gL_hint):
    net_c = net.contiguous()
    tape.net = net_c.detach()
    ctx.tape = tape
    diff, n_bar = fwd_bwd_post(tape.desc, tape.x, tape.a, tape.b, tape.c, tape.t, tape.eps,
                               net_c, tape.w, gL_hint)
    ctx.n_bar, ctx.hint = n_bar, gL_hint
    return diff

  @staticmethod
  def backward(ctx, gL):
    gL = gL.contiguous()
    n_bar = scale_rows(ctx.n_bar, gL, ctx.hint)
    ctx.n_bar = None
    return None, n_bar, gL, None


def mulan_post_fused(tape: ElboTape, net, link, gL_hint):
  """-> loss_diff[B]; the backward pass costs no extra kernel when d loss / d loss_diff_b
  equals gL_hint[b] (single backward only)."""
  return _MulanPostFused.apply(tape, net, link, gL_hint)


def mulan_pre(tape: ElboTape, x, a, b, c, t, eps0, eps):
  """-> z_t[B,D], g_net, loss_recon[B], loss_klz_prior[B], var_sums[B,2], link[B]."""
  return _MulanPre.apply(tape, x, a, b, c, t, eps0, eps)


def mulan_post(tape: ElboTape, net, link):
  """-> loss_diff[B]."""
  return _MulanPost.apply(tape, net, link)


class _AuxTopK(torch.autograd.Function):

  @staticmethod
  def forward(ctx, logits, gamma_draw, k: int):
    logits = logits.contiguous()
    emb, kl = aux_topk_fwd(logits, gamma_draw, k)
    ctx.save_for_backward(logits, gamma_draw)
    ctx.k = k
    return emb, kl

  @staticmethod
  def backward(ctx, emb_bar, kl_bar):
    logits, gamma_draw = ctx.saved_tensors
    cont = lambda v: None if v is None else v.contiguous()
    return aux_topk_bwd(logits, gamma_draw, ctx.k, cont(emb_bar), cont(kl_bar)), None, None


def aux_topk(logits, gamma_draw, k: int):
  """_topk_embedding_and_loss (ldm/model_mulan_epsilon.py:233-252) -> (embedding, kl_z)."""
  return _AuxTopK.apply(logits, gamma_draw, k)


# ------------------------------------------------------------------------------------
# The other auxiliary-latent variants (ldm/model_mulan_epsilon.py:195-219, 238-239, 264-270)
# ------------------------------------------------------------------------------------

class _AuxTopKAdd(torch.autograd.Function):
  """top-k with additive noise [B,L] (topk_noise_type == 'gumbel')."""

  @staticmethod
  def forward(ctx, logits, noise, k: int):
    logits = logits.contiguous()
    B, L = logits.shape
    _req(logits, torch.float32, (B, L), 'logits')
    _opt(noise, torch.float32, (B, L), 'noise')
    emb = torch.empty_like(logits)
    kl = torch.empty((B,), dtype=torch.float32, device=logits.device)
    _lib.check(_lib.load().mulan_aux_topk_add_fwd(B, L, k, _p(logits), _p(noise), _p(emb), _p(kl),
                                                  _stream()))
    ctx.save_for_backward(logits, noise)
    ctx.k = k
    return emb, kl

  @staticmethod
  def backward(ctx, emb_bar, kl_bar):
    logits, noise = ctx.saved_tensors
    B, L = logits.shape
    cont = lambda v: None if v is None else v.contiguous()
    out = torch.empty_like(logits)
    _lib.check(_lib.load().mulan_aux_topk_add_bwd(B, L, ctx.k, _p(logits), _p(noise),
                                                  _p(cont(emb_bar)), _p(cont(kl_bar)), _p(out),
                                                  _stream()))
    return out, None, None


class _AuxGumbel(torch.autograd.Function):
  """latent_type == 'gumbel' (_gumbel_embedding_and_loss)."""

  @staticmethod
  def forward(ctx, logits, noise, tau: float):
    logits = logits.contiguous()
    B, L = logits.shape
    _req(logits, torch.float32, (B, L), 'logits')
    _opt(noise, torch.float32, (B, L), 'gumbel_noise')
    emb = torch.empty_like(logits)
    kl = torch.empty((B,), dtype=torch.float32, device=logits.device)
    _lib.check(_lib.load().mulan_aux_gumbel_fwd(B, L, float(tau), _p(logits), _p(noise), _p(emb),
                                                _p(kl), _stream()))
    ctx.save_for_backward(logits, noise)
    ctx.tau = float(tau)
    return emb, kl

  @staticmethod
  def backward(ctx, emb_bar, kl_bar):
    logits, noise = ctx.saved_tensors
    B, L = logits.shape
    cont = lambda v: None if v is None else v.contiguous()
    out = torch.empty_like(logits)
    _lib.check(_lib.load().mulan_aux_gumbel_bwd(B, L, ctx.tau, _p(logits), _p(noise),
                                                _p(cont(emb_bar)), _p(cont(kl_bar)), _p(out),
                                                _stream()))
    return out, None, None


class _AuxGaussian(torch.autograd.Function):
  """latent_type == 'gaussian' (:264-270)."""

  @staticmethod
  def forward(ctx, mu, var, eps_z):
    mu, var = mu.contiguous(), var.contiguous()
    B, L = mu.shape
    for n, v in (('mu', mu), ('var', var), ('eps_z', eps_z)):
      _req(v, torch.float32, (B, L), n)
    emb = torch.empty_like(mu)
    kl = torch.empty((B,), dtype=torch.float32, device=mu.device)
    _lib.check(_lib.load().mulan_aux_gaussian_fwd(B, L, _p(mu), _p(var), _p(eps_z), _p(emb), _p(kl),
                                                  _stream()))
    ctx.save_for_backward(mu, var, eps_z)
    return emb, kl

  @staticmethod
  def backward(ctx, emb_bar, kl_bar):
    mu, var, eps_z = ctx.saved_tensors
    B, L = mu.shape
    cont = lambda v: None if v is None else v.contiguous()
    mb, vb = torch.empty_like(mu), torch.empty_like(mu)
    _lib.check(_lib.load().mulan_aux_gaussian_bwd(B, L, _p(mu), _p(var), _p(eps_z),
                                                  _p(cont(emb_bar)), _p(cont(kl_bar)), _p(mb),
                                                  _p(vb), _stream()))
    return mb, vb, None


def aux_topk_add(logits, noise, k: int):
  return _AuxTopKAdd.apply(logits, noise, k)


def aux_gumbel(logits, gumbel_noise, tau: float):
  return _AuxGumbel.apply(logits, gumbel_noise, tau)


def aux_gaussian(mu, var, eps_z):
  return _AuxGaussian.apply(mu, var, eps_z)


# ------------------------------------------------------------------------------------
# JAX-compatible threefry draws ("next" row 2): key = (uint32, uint32) of ONE draw
# ------------------------------------------------------------------------------------

def _rng_out(shape, dtype, device, out):
  n = int(math.prod(shape))
  if out is None:
    return torch.empty(shape, dtype=dtype, device=device), n
  return _req(out, dtype, tuple(shape), 'out'), n


def rng_bits(key, shape, device='cuda', out=None):
  """jax.random.bits(key, shape) as int64-free uint32 stored in an int32 tensor's bits."""
  out, n = _rng_out(tuple(shape), torch.int32, device, out)
  _lib.check(_lib.load().mulan_rng_bits(int(key[0]), int(key[1]), n, _p(out), _stream()))
  return out


def rng_uniform(key, shape, minval: float = 0.0, maxval: float = 1.0, device='cuda', out=None):
  """jax.random.uniform(key, shape, float32, minval, maxval)."""
  out, n = _rng_out(tuple(shape), torch.float32, device, out)
  _lib.check(_lib.load().mulan_rng_uniform(int(key[0]), int(key[1]), n, float(minval),
                                           float(maxval), _p(out), _stream()))
  return out


def rng_normal(key, shape, device='cuda', out=None):
  """jax.random.normal(key, shape, float32)."""
  out, n = _rng_out(tuple(shape), torch.float32, device, out)
  _lib.check(_lib.load().mulan_rng_normal(int(key[0]), int(key[1]), n, _p(out), _stream()))
  return out
