"""Probability-flow ODE: exact-likelihood evaluation and ODE sampling with the integrator state
resident in HBM (SURVEY.md 8f "next" row 4).

Mirrors, name for name, the reference's notebook driver:
  get_ode_likelihood_fn / likelihood_fn   ldm/notebook_utils.py:264-373
  get_sample_fn / sample_fn               ldm/notebook_utils.py:376-433
  _get_bpd_offset, eval_bpd_ode           ldm/notebook_utils.py:436-531
  Hutchinson, _prior_logp, logits_to_embeddings   :219-256, :548-551

The reference integrates with scipy.integrate.solve_ivp(method='RK45') on a float64 numpy
vector on the HOST: each of the ~100-400 function evaluations converts the whole state to
float32, ships it to the devices, runs value_div_fn there and ships the derivative back.  Here
`solve_ivp_rk45` keeps y (float64) and the seven stage derivatives (float32, exactly what the
float32 drift kernel produced) on the device (csrc/mulan_rk45.cu); per step ATTEMPT one double -
the squared error norm - comes to the host, where the accept / reject / next-step scalar
arithmetic below is the same as scipy's (Dormand-Prince 5(4); Hairer, Norsett & Wanner II.4).

Multi-GPU: rows are sharded (one process per GPU); the error norm is the one data-path
exchange - the reference's solver sees the concatenated state of all devices - so with
`group` set the two scalars (sum of squares, count) are all-reduced; the step sequence is then
identical on every rank and identical to the single-process one.

There is no CPU path: tensors must be CUDA tensors and libmulan_b200.so must be built.
"""
from __future__ import annotations

import math
from typing import Callable, NamedTuple, Optional

import numpy as np
import torch

from . import _lib, ops

# Dormand & Prince (1980) 5(4) pair; rows of A are the stage weights, E = B_5th - B_4th.
RK45_C = (0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0)
RK45_A = ((),
          (1 / 5,),
          (3 / 40, 9 / 40),
          (44 / 45, -56 / 15, 32 / 9),
          (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
          (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656))
RK45_B = (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84)
RK45_E = (-71 / 57600, 0.0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40)
SAFETY, MIN_FACTOR, MAX_FACTOR = 0.9, 0.2, 10.0
ERROR_EXPONENT = -1.0 / 5.0


class OdeSolution(NamedTuple):
  t: float
  y: torch.Tensor        # [n] float64, on the device
  nfev: int
  n_steps: int
  n_rejected: int
  status: int            # 0: reached t_bound; -1: required step fell below float spacing
  ts: list


class DeviceState:
  """y, y_new (float64), K[7] (float32) and the float32 stage vector, plus the norm plumbing.
  The only host-visible method results are Python floats (error norms)."""

  def __init__(self, y0: torch.Tensor, group=None):
    if not y0.is_cuda:
      raise TypeError('solve_ivp_rk45: y0 must be a CUDA tensor (no CPU path)')
    n = y0.numel()
    dev = y0.device
    self.n = n
    self.y = y0.detach().reshape(-1).to(torch.float64).clone()
    self.y_new = torch.empty_like(self.y)
    # rows padded to 16 B so each K[j] can be the float4 output of the drift kernel
    self.K = torch.zeros((7, (n + 3) // 4 * 4), dtype=torch.float32, device=dev)
    self.y32 = torch.empty((n,), dtype=torch.float32, device=dev)
    self._scratch = torch.empty((_lib.MULAN_RK45_SCRATCH,), dtype=torch.float64, device=dev)
    self._out = torch.empty((1,), dtype=torch.float64, device=dev)
    self.group = group
    self.n_total = n
    if group is not None:
      import torch.distributed as dist
      cnt = torch.tensor([float(n)], dtype=torch.float64, device=dev)
      dist.all_reduce(cnt, group=group)
      self.n_total = int(cnt.item())

  def stage(self, n_k: int, coef, h: float, want_new: bool = False):
    """y32 <- float32(y + (sum coef_j K_j) h)  (and y_new <- the float64 value)."""
    ops.rk45_stage(n_k, coef, h, self.y, self.K, y_stage=self.y32,
                   y_out=self.y_new if want_new else None)

  def rms(self, n_k: int, coef, h: float, rtol: float, atol: float, of_y: bool = False,
          with_new: bool = False) -> float:
    ops.rk45_norm(n_k, coef, h, rtol, atol, self.y, self.y_new if with_new else None, self.K,
                  of_y, self._scratch, self._out)
    if self.group is not None:
      import torch.distributed as dist
      dist.all_reduce(self._out, group=self.group)
    return math.sqrt(self._out.item() / self.n_total)

  def accept(self):
    self.y, self.y_new = self.y_new, self.y
    self.K[0].copy_(self.K[6])                 # first-same-as-last

  def k_row(self, j: int) -> torch.Tensor:
    return self.K[j, :self.n]


def _initial_step(fun, st, t0, t_bound, direction, rtol, atol):
  """Hairer-Norsett-Wanner II.4 starting step; K[0] holds f0, K[1] receives f1."""
  interval = abs(t_bound - t0)
  if interval == 0.0:
    return 0.0
  d0 = st.rms(0, (), 0.0, rtol, atol, of_y=True)
  d1 = st.rms(1, (1.0,), 1.0, rtol, atol)
  h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
  h0 = min(h0, interval)
  st.stage(1, (1.0,), h0 * direction)
  fun(t0 + h0 * direction, st.k_row(1))
  d2 = st.rms(2, (-1.0, 1.0), 1.0, rtol, atol) / h0
  if d1 <= 1e-15 and d2 <= 1e-15:
    h1 = max(1e-6, h0 * 1e-3)
  else:
    h1 = (0.01 / max(d1, d2)) ** (1.0 / 5.0)
  return min(100 * h0, h1, interval)


def solve_ivp_rk45(fun: Callable, t_span, y0: torch.Tensor, rtol: float = 1e-3,
                   atol: float = 1e-6, group=None, _state=None) -> OdeSolution:
  """scipy.integrate.solve_ivp(fun, t_span, y0, method='RK45', rtol, atol) with the state on
  the device.  fun(t: float, y32: Tensor[n] float32, out: Tensor[n] float32) must WRITE dy/dt
  into `out` (a row of K), enqueued on the current stream.
  (_state: the CPU test suite substitutes a stand-in for DeviceState to exercise this host
  control flow without a GPU; nothing in the package passes it.)"""
  t, t_bound = float(t_span[0]), float(t_span[1])
  rtol = max(float(rtol), 100 * np.finfo(float).eps)
  atol = float(atol)
  direction = (1.0 if t_bound > t else -1.0) if t_bound != t else 1.0
  st = DeviceState(y0, group) if _state is None else _state
  nfev = 0

  def f(tt, out_row):
    nonlocal nfev
    nfev += 1
    fun(tt, st.y32, out_row)

  st.stage(0, (), 0.0)
  f(t, st.k_row(0))
  h_abs = _initial_step(f, st, t, t_bound, direction, rtol, atol)
  n_steps = n_rej = 0
  ts = [t]
  status = 0
  while t != t_bound:
    min_step = 10 * abs(float(np.nextafter(t, direction * np.inf)) - t)
    h_abs = max(h_abs, min_step)
    rejected = False
    while True:
      if h_abs < min_step:
        status = -1
        break
      h = h_abs * direction
      t_new = t + h
      if direction * (t_new - t_bound) > 0:
        t_new = t_bound
      h = t_new - t
      h_abs = abs(h)
      for s in range(1, 6):
        st.stage(s, RK45_A[s], h)
        f(t + RK45_C[s] * h, st.k_row(s))
      st.stage(6, RK45_B, h, want_new=True)
      f(t + h, st.k_row(6))
      err = st.rms(7, RK45_E, h, rtol, atol, with_new=True)
      if err < 1:
        factor = MAX_FACTOR if err == 0 else min(MAX_FACTOR, SAFETY * err ** ERROR_EXPONENT)
        if rejected:
          factor = min(1.0, factor)
        h_abs *= factor
        break
      h_abs *= max(MIN_FACTOR, SAFETY * err ** ERROR_EXPONENT)
      rejected = True
      n_rej += 1
    if status != 0:
      break
    st.accept()
    t = t_new
    n_steps += 1
    ts.append(t)
  return OdeSolution(t, st.y, nfev, n_steps, n_rej, status, ts)


# ---------------------------------------------------------------------------------------------
# notebook_utils.py drivers
# ---------------------------------------------------------------------------------------------

def _prior_logp(z: torch.Tensor) -> torch.Tensor:
  """notebook_utils.py:219-222: standard-normal log-density per row; z [B, D] float32."""
  B, D = z.shape
  return ops.row_dot(z, z).mul_(-0.5).add_(-0.5 * D * math.log(2 * math.pi))


def logits_to_embeddings(logits: torch.Tensor, k: int = 15) -> torch.Tensor:
  """notebook_utils.py:548-551: the hard top-k mask of raw logits."""
  top = torch.topk(logits, k, dim=1).values
  return (logits >= top[:, -1][:, None]).to(torch.float32)


class Hutchinson:
  """notebook_utils.py:227-256; torch.Generator instead of a jax PRNGKey."""

  def __init__(self, hutchinson_type: str, shape, generator, deterministic: bool = False,
               device='cuda'):
    if hutchinson_type not in ('Gaussian', 'Rademacher'):
      raise ValueError(f'unknown hutchinson_type {hutchinson_type!r}')
    self.hutchinson_type, self.shape = hutchinson_type, tuple(shape)
    self.generator, self.deterministic, self.device = generator, deterministic, device
    if deterministic:
      self.det_noise = self._sample_noise()

  def noise(self):
    return self.det_noise if self.deterministic else self._sample_noise()

  def _sample_noise(self):
    if self.hutchinson_type == 'Gaussian':
      return torch.randn(self.shape, generator=self.generator, device=self.device)
    bits = torch.randint(0, 2, self.shape, generator=self.generator, device=self.device)
    return bits.to(torch.float32) * 2 - 1


def _drift_fun(model, embeddings, hutchinson, B, high_precision, with_div):
  """ode_func (notebook_utils.py:350-358 / :417-421) writing straight into a row of K."""
  from .model import value_div_fn
  D = 32 * 32 * 3
  with torch.no_grad():
    coeffs = tuple(q.contiguous() for q in model.gamma._compute_coefficients(embeddings))

  def fun(t, y32, out):
    x = y32[:B * D].reshape(B, 32, 32, 3)
    if with_div:
      value_div_fn(model, x, embeddings, t, hutchinson.noise(), high_precision, coeffs=coeffs,
                   out=(out[:B * D].reshape(B, D), out[B * D:]))
    else:
      value_div_fn(model, x, embeddings, t, None, high_precision, coeffs=coeffs,
                   out=(out.reshape(B, D), None))

  return fun


def get_ode_likelihood_fn(model, hutchinson_type: str = 'Rademacher', rtol: float = 1e-5,
                          atol: float = 1e-5, method: str = 'RK45',
                          dequantization: str = 'uniform', high_precision: bool = False,
                          group=None):
  """notebook_utils.py:264-373.  `likelihood_fn(generator, data[B,32,32,3], deterministic_noise,
  draws=None)` -> (log_p[B], log_q_eps[B] | None, aux_loss[B]); the last OdeSolution is left in
  `likelihood_fn.last_solution`.  `draws` = {'u': ..., 'hutchinson': ...} overrides the two
  random draws (tests).  Reference quirk kept visible: with dequantization='uniform' the
  reference has no log_q_eps (it would fail on `None.reshape`, :371); here it is None."""
  if method != 'RK45':
    raise NotImplementedError('only method="RK45" (the reference\'s default) is on the device')
  if dequantization not in ('uniform', 'tn'):
    raise ValueError(f'unknown dequantization {dequantization!r}')
  D = 32 * 32 * 3

  @torch.no_grad()
  def likelihood_fn(generator, data, deterministic_noise: bool = False, draws=None):
    dev = data.device
    shape = tuple(data.shape)
    B = shape[0]
    data = model.encdec.encode(data.to(torch.float32).round())                   # :313
    draws = draws or {}
    if dequantization == 'uniform':
      u = draws['u'] if 'u' in draws else torch.rand(shape, generator=generator, device=dev)
      u = 2 * (u - 0.5) / 256
      log_q_eps = None
    else:
      gt = -13.3
      if 'u' in draws:
        u = draws['u']
      else:
        u = torch.empty(shape, device=dev)
        torch.nn.init.trunc_normal_(u, 0.0, 1.0, -3.0, 3.0, generator=generator)
      log_q_eps = _prior_logp(u.reshape(B, D).contiguous()) - D * math.log(0.9974613)
      u = u * math.exp(0.5 * gt)
    data = data + u
    logits = model.apply_encoder(torch.clip(128 * (data + 1) - 0.5, 0, 255).round())
    _, aux = ops.aux_topk_add(logits.contiguous(), None, model.config.latent_k)  # _gumbel_kl_loss
    embeddings = logits_to_embeddings(logits, model.config.latent_k)
    if 'hutchinson' in draws:
      hutch = Hutchinson(hutchinson_type, shape, None, deterministic=False, device=dev)
      hutch.deterministic, hutch.det_noise = True, draws['hutchinson']
    else:
      hutch = Hutchinson(hutchinson_type, shape, generator, deterministic_noise, dev)
    fun = _drift_fun(model, embeddings, hutch, B, high_precision, with_div=True)
    init = torch.cat([data.reshape(-1).to(torch.float64),
                      torch.zeros((B,), dtype=torch.float64, device=dev)])
    sol = solve_ivp_rk45(fun, (0.0, 1.0), init, rtol=rtol, atol=atol, group=group)
    likelihood_fn.last_solution = sol
    zp = sol.y.to(torch.float32)                                                   # :363
    z = zp[:B * D].reshape(B, D).contiguous()
    delta_logp = zp[B * D:]
    log_p = _prior_logp(z) + delta_logp
    return log_p, log_q_eps, aux

  likelihood_fn.last_solution = None
  return likelihood_fn


def get_sample_fn(model, hutchinson_type: str = 'Rademacher', rtol: float = 1e-5,
                  atol: float = 1e-5, method: str = 'RK45', high_precision: bool = False,
                  group=None):
  """notebook_utils.py:376-433: integrate the drift from t=1 to t=0 starting at N(0, I), with
  embeddings = hard top-k of random normal logits.  -> sample_fn(generator, sample_size, ...)
  -> (z[B,32,32,3], nfev)."""
  if method != 'RK45':
    raise NotImplementedError('only method="RK45" is on the device')

  @torch.no_grad()
  def sample_fn(generator, deterministic_noise: bool = False, sample_size: int = 32,
                device='cuda', draws=None):
    draws = draws or {}
    L, k = model.config.latent_size, model.config.latent_k
    logits = draws['logits'] if 'logits' in draws else torch.randn(
        (sample_size, L), generator=generator, device=device)
    embeddings = logits_to_embeddings(logits, k)
    B = sample_size
    fun = _drift_fun(model, embeddings, None, B, high_precision, with_div=False)
    prior = draws['prior'] if 'prior' in draws else torch.randn(
        (B, 32, 32, 3), generator=generator, device=device)
    sol = solve_ivp_rk45(fun, (1.0, 0.0), prior.reshape(-1), rtol=rtol, atol=atol, group=group)
    return sol.y.to(torch.float32).reshape(B, 32, 32, 3), sol.nfev

  return sample_fn


def _get_bpd_offset(dequantization: str, num_is: int) -> float:
  """notebook_utils.py:436-448."""
  if dequantization == 'uniform':
    return math.log2(128)
  if dequantization != 'tn':
    raise ValueError(f'unknown dequantization {dequantization!r}')
  gt = -13.3
  log_sigma = 0.5 * (gt - math.log1p(math.exp(gt)))
  extra_terms = 0.5 * (1 + math.log(2 * math.pi)) - 0.01522 if num_is == 1 else 0.0
  return -(extra_terms + log_sigma) / math.log(2)


def eval_bpd_ode(model, batches, deterministic_noise: bool, hutchinson_type: str,
                 dequantization: str = 'tn', num_is: int = 1, rtol: float = 1e-5,
                 atol: float = 1e-5, generator: Optional[torch.Generator] = None, group=None):
  """_eval_bpd_ode (notebook_utils.py:484-531) over an iterable of uint8/int image batches
  [B,32,32,3] (this rank's shard): -> mean bits/dim over the batches."""
  likelihood_function = get_ode_likelihood_fn(
      model, rtol=rtol, atol=atol, hutchinson_type=hutchinson_type,
      dequantization=dequantization, group=group)
  offset = _get_bpd_offset(dequantization, num_is)
  bpds = []
  for data in batches:
    log_ps, log_qs = [], []
    for _ in range(num_is):
      log_p, log_q, aux_loss = likelihood_function(generator, data,
                                                   deterministic_noise=deterministic_noise)
      log_ps.append(log_p)
      log_qs.append(log_q)
    if num_is == 1:
      iws = log_ps[0]
    else:
      iws = torch.logsumexp(torch.stack(log_ps) - torch.stack(log_qs), dim=0) - math.log(num_is)
    bpd = torch.mean(-iws + aux_loss) / (32 * 32 * 3 * math.log(2)) + offset
    if group is not None:
      import torch.distributed as dist
      dist.all_reduce(bpd, group=group)
      bpd = bpd / dist.get_world_size(group)
    bpds.append(bpd.item())
  return float(np.mean(bpds))
