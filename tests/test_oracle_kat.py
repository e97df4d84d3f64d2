"""Analytic known-answer tests for the CPU oracle: properties the
reference's algorithm must satisfy regardless of implementation."""
import math

import numpy as np
import pytest
import torch

from oracle import mulan_oracle as O

CFG = O.OracleConfig()


def _abc(B=6, seed=0, dtype=torch.float64):
  i = O.synth_inputs(B, seed, dtype=dtype)
  return i


def test_fixed_ends():
  """gamma(0) == gamma_min exactly; gamma(1) == gamma_max within an ulp of float32."""
  i = _abc(dtype=torch.float32)
  z = torch.zeros(6, 1)
  g0 = O.eval_polynomial(i['a'], i['b'], i['c'], z, CFG)
  g1 = O.eval_polynomial(i['a'], i['b'], i['c'], z + 1, CFG)
  assert torch.all(g0 == torch.tensor(-13.3, dtype=torch.float32))
  # (Delta*S)/S is Delta +- 1 ulp(18.3) = 1.9e-6, then one more rounding at 5
  # (f32(-13.3) + f32(18.3) is already 0.95e-6 below 5)
  assert (g1 - 5.0).abs().max().item() <= 4e-6
  assert torch.unique(g1).numel() <= 3
  assert torch.unique(O.sigmoid(g1)).numel() == 1   # what the kernels' constant path relies on


def test_gamma_monotone_and_derivative():
  i = _abc()
  ts = torch.linspace(0, 1, 33, dtype=torch.float64)
  gs = torch.stack([O.eval_polynomial(i['a'], i['b'], i['c'], torch.full((6, 1), float(t),
                    dtype=torch.float64), CFG) for t in ts])
  assert torch.all(gs[1:] >= gs[:-1] - 1e-12)
  t = torch.full((6, 1), 0.37, dtype=torch.float64)
  h = 1e-6
  fd = (O.eval_polynomial(i['a'], i['b'], i['c'], t + h, CFG)
        - O.eval_polynomial(i['a'], i['b'], i['c'], t - h, CFG)) / (2 * h)
  dg = O.eval_polynomial_dt(i['a'], i['b'], i['c'], t, CFG)
  assert torch.all(dg >= 0)
  assert ((fd - dg).abs() / (dg.abs() + 1e-3 * dg.abs().max())).max().item() < 1e-6
  # closed form the kernels use: Delta (a t^2 + b t + c)^2 / S
  q = i['a'] * t * t + i['b'] * t + i['c']
  S = (i['a'] ** 2 / 5 + (i['b'] ** 2 + 2 * i['a'] * i['c']) / 3 + i['a'] * i['b'] / 2
       + i['b'] * i['c'] + i['c'] ** 2)
  assert ((18.3 * q * q / S - dg).abs() / (dg.abs() + 1e-3 * dg.abs().max())).max().item() < 1e-10


def test_integer_pow_matches_multiply_chain():
  t = torch.rand(100, dtype=torch.float32)
  assert torch.equal(O.integer_pow(t, 5), t * ((t * t) * (t * t)))
  assert torch.equal(O.integer_pow(t, 3), t * (t * t))
  assert torch.equal(O.integer_pow(t, 4), (t * t) * (t * t))


def test_decode_is_a_distribution():
  z = torch.tensor([[-0.9, 0.0, 0.31, 1.2]], dtype=torch.float64)
  for g0 in (-13.3, -6.0, 0.0):
    lp = O.decode(z, torch.full_like(z, g0), 256)
    assert torch.allclose(torch.exp(lp).sum(-1), torch.ones_like(z), atol=1e-12)


def test_encode_range_and_bins():
  x = torch.arange(256)
  f = O.encode(x, 256)
  assert f.min().item() == -1 + 1 / 256 and f.max().item() == 1 - 1 / 256
  assert torch.all((f[1:] - f[:-1]) == 2 / 256)


def test_loss_terms_nonnegative_and_kl_uniform_zero():
  i = _abc(dtype=torch.float32)
  out = O.elbo_terms(i['x'], i['a'], i['b'], i['c'], i['t'], i['eps_0'], i['eps'],
                     lambda z, g: i['net'], O.MODE_EPS, CFG)
  assert torch.all(out.loss_recon >= 0) and torch.all(out.loss_klz >= 0)
  assert torch.all(out.loss_diff >= 0)
  assert O.gumbel_kl_loss(torch.zeros(3, 50), 50).abs().max().item() < 1e-6
  lg = torch.randn(5, 50, dtype=torch.float64)
  assert torch.all(O.gumbel_kl_loss(lg, 50) >= 0)


def test_v_from_eps_equals_eps_loss():
  """(v_target - v_hat) = (eps - net)/alpha  =>  the v-from-eps loss equals the eps loss."""
  i = _abc(dtype=torch.float64)
  args = (i['x'], i['a'], i['b'], i['c'], i['t'], i['eps_0'], i['eps'], lambda z, g: i['net'])
  e = O.elbo_terms(*args, O.MODE_EPS, CFG, dtype=torch.float64)
  v = O.elbo_terms(*args, O.MODE_VEL_FROM_EPS, CFG, dtype=torch.float64)
  assert ((e.loss_diff - v.loss_diff).abs() / e.loss_diff).max().item() < 1e-10


def test_prior_kl_closed_form():
  i = _abc(dtype=torch.float64)
  out, aux = O.elbo_terms(i['x'], i['a'], i['b'], i['c'], i['t'], i['eps_0'], i['eps'],
                          lambda z, g: i['net'], O.MODE_EPS, CFG, dtype=torch.float64,
                          return_aux=True)
  v1 = 1 / (1 + math.exp(-5.0))
  f = aux['orig_f'].reshape(6, -1)
  want = 0.5 * ((1 - v1) * f ** 2 + v1 - math.log(v1) - 1).sum(1)
  assert ((aux['loss_klz_prior'] - want).abs() / want).max().item() < 1e-9


def test_recon_window_claim():
  """At gamma_0 = -13.3 a +-1-bin window reproduces the 256-bin log-softmax (what the CUDA
  kernel evaluates)."""
  rng = np.random.default_rng(0)
  x = torch.from_numpy(rng.integers(0, 256, (4, 3072), dtype=np.uint8))
  e0 = torch.from_numpy(rng.standard_normal((4, 3072))).double() * 1.5
  f = O.encode(x, 256, torch.float64)
  g0 = torch.full_like(f, -13.3)
  z = f + torch.exp(0.5 * g0) * e0
  full = O.logprob(x, z, g0, 256)
  inv = math.exp(6.65)
  k = torch.clamp(torch.round((z + 1) * 128 - 0.5), 0, 255)
  def logit(kk):
    return -0.5 * ((z - (2 * (kk + .5) / 256 - 1)) * inv) ** 2
  lc = logit(k)
  lm = torch.where(k > 0, logit(k - 1), torch.full_like(z, -math.inf))
  lp = torch.where(k < 255, logit(k + 1), torch.full_like(z, -math.inf))
  lse = lc + torch.log(1 + torch.exp(lm - lc) + torch.exp(lp - lc))
  win = (logit(x.double()) - lse).sum(1)
  assert ((win - full).abs() / full.abs()).max().item() < 1e-12


def test_sample_t_antithetic():
  t = O.sample_t(0.73, 8, CFG)
  assert t.shape == (8,)
  d = torch.sort(t).values
  assert torch.allclose(d[1:] - d[:-1], torch.full((7,), 0.125), atol=1e-6)
  assert torch.all((t >= 0) & (t < 1))
  cfgT = O.OracleConfig(sm_n_timesteps=10)
  tt = O.sample_t(0.73, 8, cfgT)
  assert torch.allclose(tt * 10, torch.round(tt * 10))


def test_topk_embedding_properties():
  rng = np.random.default_rng(1)
  lg = torch.from_numpy(rng.standard_normal((7, 50)))
  G = torch.from_numpy(rng.gamma(1 / 15, size=(10, 7, 50)))
  emb, kl = O.topk_embedding_and_loss(lg, G, 15, 50)
  assert torch.all((emb > 0.5).sum(1) == 15)
  assert torch.allclose(emb, (emb > 0.5).double(), atol=1e-12)   # straight-through value is hard
  assert torch.all(kl >= 0)


def test_bpd_assembly():
  out = O.VDMOutput(torch.tensor([1.0, 3.0]), torch.tensor([2.0, 2.0]), torch.tensor([10., 30.]),
                    torch.tensor(0.1), torch.tensor(0.9))
  bpd, sc = O.loss_fn_bpd(out)
  assert abs(bpd.item() - (2 + 2 + 20) / (3072 * math.log(2))) < 1e-9
  assert set(sc) == {'bpd', 'bpd_latent', 'bpd_recon', 'bpd_diff', 'var0', 'var'}
