"""TEST INFRASTRUCTURE ONLY - CPU restatement of the adaptive RK45 integration that drives the
reference's exact-likelihood evaluation and ODE sampler.  Only tests/, __graft_entry__.smoke()
and bench.py's CPU legs may import this file; the product (mulan_b200/ode.py) never does.

What it restates
  * scipy.integrate.solve_ivp(method='RK45') as the reference calls it
    (ldm/notebook_utils.py:353, :427): an UN-VENDORED third-party dependency (scipy, unpinned in
    the reference's requirements.txt; scipy 1.18.1 is installed in this image).  Published
    algorithm: Dormand & Prince, "A family of embedded Runge-Kutta formulae" (1980), with the
    initial-step and step-size control of Hairer, Norsett & Wanner, "Solving Ordinary
    Differential Equations I", Sec. II.4 (safety 0.9, factors in [0.2, 10], RMS error norm).
  * likelihood_fn / sample_fn around it (ldm/notebook_utils.py:264-373, :376-433) with every
    random draw an INPUT.

Pinning: tests/test_ode.py checks `solve_rk45` step for step against scipy itself (same
accepted times, same nfev, same final state to 1e-13) and the tableau against scipy's class
attributes, so parity for the integrator is pinned to the real dependency.
"""
from __future__ import annotations

import math
from typing import Callable, NamedTuple

import numpy as np

# Dormand-Prince 5(4) tableau (Dormand & Prince 1980, Table 2).
C = np.array([0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1])
A = np.array([
    [0, 0, 0, 0, 0],
    [1 / 5, 0, 0, 0, 0],
    [3 / 40, 9 / 40, 0, 0, 0],
    [44 / 45, -56 / 15, 32 / 9, 0, 0],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729, 0],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
])
B = np.array([35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84])
E = np.array([-71 / 57600, 0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40])

SAFETY, MIN_FACTOR, MAX_FACTOR = 0.9, 0.2, 10.0
ERROR_EXPONENT = -1.0 / 5.0      # -1 / (error_estimator_order + 1), order 4


class OdeResult(NamedTuple):
  t: float
  y: np.ndarray
  nfev: int
  n_steps: int
  n_rejected: int
  status: int            # 0 reached t_bound, -1 step size fell below the float spacing
  ts: list


def rms(x):
  return np.linalg.norm(x) / x.size ** 0.5


def initial_step(fun, t0, y0, t_bound, f0, direction, rtol, atol):
  """Hairer II.4 starting step (order 4 error estimator); one extra evaluation of fun."""
  interval = abs(t_bound - t0)
  if interval == 0.0:
    return 0.0
  scale = atol + np.abs(y0) * rtol
  d0, d1 = rms(y0 / scale), rms(f0 / scale)
  h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
  h0 = min(h0, interval)
  f1 = fun(t0 + h0 * direction, y0 + h0 * direction * f0)
  d2 = rms((f1 - f0) / scale) / h0
  if d1 <= 1e-15 and d2 <= 1e-15:
    h1 = max(1e-6, h0 * 1e-3)
  else:
    h1 = (0.01 / max(d1, d2)) ** (1.0 / 5.0)
  return min(100 * h0, h1, interval)


def solve_rk45(fun: Callable, t_span, y0, rtol=1e-3, atol=1e-6) -> OdeResult:
  """fun(t, y[n] float64) -> dy/dt [n].  Same control flow as solve_ivp(method='RK45')."""
  t, t_bound = float(t_span[0]), float(t_span[1])
  y = np.asarray(y0).astype(np.float64)
  rtol = max(rtol, 100 * np.finfo(float).eps)
  direction = float(np.sign(t_bound - t)) if t_bound != t else 1.0
  nfev = 0

  def f_counted(tt, yy):
    nonlocal nfev
    nfev += 1
    return np.asarray(fun(tt, yy), dtype=np.float64)

  f = f_counted(t, y)
  h_abs = initial_step(f_counted, t, y, t_bound, f, direction, rtol, atol)
  K = np.empty((7, y.size))
  n_steps = n_rej = 0
  ts = [t]
  status = 0
  while t != t_bound:
    min_step = 10 * abs(np.nextafter(t, direction * np.inf) - t)
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
      K[0] = f
      for s in range(1, 6):
        dy = np.dot(K[:s].T, A[s, :s]) * h
        K[s] = f_counted(t + C[s] * h, y + dy)
      y_new = y + h * np.dot(K[:6].T, B)
      f_new = f_counted(t + h, y_new)
      K[6] = f_new
      scale = atol + np.maximum(np.abs(y), np.abs(y_new)) * rtol
      err = rms(np.dot(K.T, E) * h / scale)
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
    t, y, f = t_new, y_new, f_new
    n_steps += 1
    ts.append(t)
  return OdeResult(t, y, nfev, n_steps, n_rej, status, ts)


# ---------------------------------------------------------------------------------------------
# likelihood_fn / sample_fn  (ldm/notebook_utils.py:264-373, :376-433)
# ---------------------------------------------------------------------------------------------

def prior_logp(z):
  """_prior_logp (notebook_utils.py:219-222); z [B, ...] torch tensor."""
  import torch
  n = int(np.prod(z.shape[1:]))
  return -0.5 * n * math.log(2 * math.pi) - 0.5 * torch.sum(z ** 2, dim=tuple(range(1, z.dim())))


def logits_to_embeddings(logits, k: int = 15):
  """notebook_utils.py:548-551."""
  import torch
  top = torch.topk(logits, k, dim=1).values
  return (logits >= top[:, -1][:, None]).to(logits.dtype)


def bpd_offset(dequantization: str, num_is: int):
  """_get_bpd_offset (notebook_utils.py:436-448)."""
  if dequantization == 'uniform':
    return math.log2(128)
  gt = -13.3
  log_sigma = 0.5 * (gt - math.log1p(math.exp(gt)))
  extra = 0.5 * (1 + math.log(2 * math.pi)) - 0.01522 if num_is == 1 else 0.0
  return -(extra + log_sigma) / math.log(2)


def likelihood(data_u8, u, hutchinson_noise, encoder_fn: Callable, value_div_fn: Callable,
               dequantization: str = 'tn', rtol=1e-5, atol=1e-5):
  """likelihood_fn (notebook_utils.py:303-371) on ONE device's batch, deterministic noise.

  data_u8 [B,32,32,3]; u = the dequantisation draw (uniform in [0,1) for 'uniform', truncated
  normal in [-3,3] for 'tn'); hutchinson_noise [B,32,32,3];
  encoder_fn(images_int float [B,32,32,3]) -> logits [B,50];
  value_div_fn(x[B,32,32,3] f32, embeddings, t float, noise) -> (drift, logp_grad[B]).
  Returns (log_p[B], log_q_eps[B] | None, aux_loss[B], OdeResult).
  """
  import torch
  from oracle import mulan_oracle as O
  B = data_u8.shape[0]
  shape = tuple(data_u8.shape)
  data = 2 * ((data_u8.to(torch.float32).round() + .5) / 256) - 1            # :313
  if dequantization == 'uniform':
    du = 2 * (u - 0.5) / 256
    log_q_eps = None
  else:
    gt = -13.3
    log_q_eps = prior_logp(u) - (32 * 32 * 3) * math.log(0.9974613)         # :331
    du = u * math.exp(0.5 * gt)
  data = data + du
  logits = encoder_fn(torch.clip(128 * (data + 1) - 0.5, 0, 255).round())    # :341
  aux = O.gumbel_kl_loss(logits, logits.shape[-1])
  emb = logits_to_embeddings(logits)

  def ode_func(t, x):                                                        # :350-358
    xt = torch.from_numpy(np.asarray(x[:-B], dtype=np.float32)).reshape(shape)
    drift, logp_grad = value_div_fn(xt, emb, t, hutchinson_noise)
    return np.concatenate([drift.detach().double().numpy().reshape(-1),
                           logp_grad.detach().double().numpy().reshape(-1)])

  init = np.concatenate([data.double().numpy().reshape(-1), np.zeros(B)])
  sol = solve_rk45(ode_func, (0.0, 1.0), init, rtol=rtol, atol=atol)
  zp = torch.from_numpy(sol.y.astype(np.float32))                            # :363
  z = zp[:-B].reshape(shape)
  delta_logp = zp[-B:]
  log_p = prior_logp(z) + delta_logp
  return log_p, log_q_eps, aux, sol
