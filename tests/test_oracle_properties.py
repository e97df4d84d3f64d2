"""Property tests (hypothesis) of the oracle's schedule and loss algebra over the coefficient
ranges SURVEY.md 8c asks for: a == 0 (the reference's zero-initialised head), |a|, |b| up to 30,
c down to its floor 1e-3, t at and near the ends.  float64, so the properties are exact up to
rounding; they bound what every kernel test then compares against."""
import math

import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import mulan_oracle as O

CFG = O.OracleConfig()
coef = st.floats(min_value=-30.0, max_value=30.0, allow_nan=False, allow_infinity=False)
craw = st.floats(min_value=-12.0, max_value=12.0, allow_nan=False)
tval = st.one_of(st.sampled_from([0.0, 1.0, 1e-6, 1.0 - 1e-6, 0.5]),
                 st.floats(min_value=0.0, max_value=1.0, allow_nan=False))


def _abc(a, b, cr):
  t64 = lambda v: torch.tensor([[v]], dtype=torch.float64)
  return t64(a), t64(b), O.coefficients_from_raw(t64(cr))


@settings(max_examples=200, deadline=None)
@given(a=st.one_of(st.just(0.0), coef), b=coef, cr=craw, t=tval)
def test_gamma_is_monotone_between_fixed_ends(a, b, cr, t):
  A, B, C = _abc(a, b, cr)
  assert C.item() >= 1e-3                                  # c = 1e-3 + softplus(.)
  T = torch.tensor([[t]], dtype=torch.float64)
  g = O.eval_polynomial(A, B, C, T, CFG).item()
  g0 = O.eval_polynomial(A, B, C, torch.zeros(1, 1, dtype=torch.float64), CFG).item()
  g1 = O.eval_polynomial(A, B, C, torch.ones(1, 1, dtype=torch.float64), CFG).item()
  assert g0 == CFG.gamma_min                               # P(0) == 0 exactly
  assert abs(g1 - CFG.gamma_max) < 1e-12 * 20              # (Delta*S)/S
  assert g0 - 1e-9 <= g <= g1 + 1e-9
  w = O.eval_polynomial_dt(A, B, C, T, CFG).item()
  assert w >= -1e-9 * (1 + abs(w))                         # d gamma / dt = Delta q^2 / S >= 0


@settings(max_examples=60, deadline=None)
@given(a=coef, b=coef, cr=craw)
def test_dgamma_dt_integrates_to_the_gamma_range(a, b, cr):
  """int_0^1 d gamma/dt dt == gamma(1) - gamma(0) (Gauss-Legendre, exact for the quartic q^2)."""
  A, B, C = _abc(a, b, cr)
  xs, ws = np.polynomial.legendre.leggauss(8)
  ts = torch.tensor(0.5 * (xs + 1.0), dtype=torch.float64).reshape(-1, 1)
  w = O.eval_polynomial_dt(A.expand(8, 1), B.expand(8, 1), C.expand(8, 1), ts, CFG).reshape(-1)
  integral = float((w * torch.tensor(0.5 * ws)).sum())
  assert abs(integral - (CFG.gamma_max - CFG.gamma_min)) < 1e-9 * (CFG.gamma_max - CFG.gamma_min)


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(min_value=0, max_value=2 ** 31 - 1), B=st.sampled_from([1, 2, 3, 8]))
def test_loss_algebra(seed, B):
  """All three terms are non-negative; velocity_from_epsilon equals the epsilon loss (the
  identity the kernels use, include/mulan_b200.h: mulan_kernel_param); the plain velocity loss
  with the exact velocity as network output is zero."""
  inp = O.synth_inputs(B, seed % 10_000, D=48, dtype=torch.float64)
  args = (inp['x'], inp['a'], inp['b'], inp['c'], inp['t'], inp['eps_0'], inp['eps'])
  e, aux = O.elbo_terms(*args, lambda z, g: inp['net'], O.MODE_EPS, CFG, dtype=torch.float64,
                        return_aux=True)
  v = O.elbo_terms(*args, lambda z, g: inp['net'], O.MODE_VEL_FROM_EPS, CFG, dtype=torch.float64)
  assert (e.loss_recon >= 0).all() and (e.loss_klz >= 0).all() and (e.loss_diff >= 0).all()
  assert ((e.loss_diff - v.loss_diff).abs() <= 1e-10 * e.loss_diff.abs() + 1e-12).all()
  # plain velocity model fed the exact velocity target: zero diffusion loss
  var_t = O.sigmoid(aux['g_t'])
  v_target = torch.sqrt(1 - var_t) * inp['eps'] - torch.sqrt(var_t) * aux['orig_f']
  z = O.elbo_terms(*args, lambda z_, g: v_target, O.MODE_VEL, CFG, dtype=torch.float64)
  assert z.loss_diff.abs().max().item() < 1e-18


@settings(max_examples=100, deadline=None)
@given(t0=st.floats(min_value=0.0, max_value=1.0, exclude_max=True, allow_nan=False),
       B=st.sampled_from([1, 2, 8, 127, 128]), T=st.sampled_from([0, 10, 1000]))
def test_sample_t_is_a_stratified_cover(t0, B, T):
  """ldm/model_mulan_epsilon.py:287-297: one t per stratum of width 1/B; discretised times are
  multiples of 1/T in (0, 1]."""
  cfg = O.OracleConfig(sm_n_timesteps=T)
  t = O.sample_t(t0, B, cfg, torch.float64).numpy()
  assert t.shape == (B,) and (t >= 0).all() and (t <= 1).all()
  if T == 0:
    assert (t < 1).all()
    strata = np.sort(np.floor(t * B + 1e-9).astype(int) % B)
    assert np.array_equal(strata, np.arange(B)) or B == 1
  else:
    k = t * T
    assert np.allclose(k, np.round(k), atol=1e-9)
