"""Ancestral sampler step + decode: oracle and CUDA kernels against the golden vectors
produced by executing the reference's own VDM.sample / VDM.generate_x
(tests/golden/make_golden_sampler.py)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import mulan_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
import golden_inputs as GI  # noqa: E402

G = np.load(os.path.join(HERE, 'golden', 'sampler.npz'))
SEED, B, T = int(G['seed']), int(G['B']), int(G['T'])
STEPS = [int(i) for i in G['steps']]


def sampler_inputs():
  r = np.random.default_rng(SEED)
  f32 = np.float32
  return dict(z_t=r.standard_normal((B, 32, 32, 3)).astype(f32),
              eps=r.standard_normal((len(STEPS), B, 32, 32, 3)).astype(f32),
              noise=(0.3 * r.standard_normal((B, 32, 32, 3))).astype(f32),
              z_0=(r.uniform(-1.1, 1.1, (B, 32, 32, 3))).astype(f32))


@pytest.mark.parametrize('kind', ['eps', 'vel'])
@pytest.mark.parametrize('tag', ['f32', 'f64'])
def test_oracle_sampler_matches_reference_source(kind, tag):
  dtype = torch.float32 if tag == 'f32' else torch.float64
  cfg = O.OracleConfig()
  inp = sampler_inputs()
  tt = lambda v: torch.from_numpy(np.asarray(v)).to(dtype)
  W = {k: tt(v) for k, v in GI.mlp_weights(SEED + 1000).items()}
  emb = O.deterministic_embedding(B, cfg, dtype)
  a, b, c = O.compute_coefficients(W, emb)
  z_t, noise = tt(inp['z_t']), tt(inp['noise'])
  mode = O.MODE_EPS if kind == 'eps' else O.MODE_VEL
  for n, i in enumerate(STEPS):
    t = torch.full((B, 1), (T - i) / T, dtype=dtype)
    s = torch.full((B, 1), (T - i - 1) / T, dtype=dtype)
    g_t = O.eval_polynomial(a, b, c, t, cfg).reshape(B, 32, 32, 3)
    g_s = O.eval_polynomial(a, b, c, s, cfg).reshape(B, 32, 32, 3)
    g_net = O.score_model_gt(g_t, cfg)
    net = 0.7 * z_t + 0.05 * g_net.reshape(-1, 1, 1, 1) + noise
    z_s = O.sample_step(z_t, g_t, g_s, net, tt(inp['eps'][n]), mode)
    want = G[f'{kind}_{tag}_z_s_{i}']
    # The reference's float32 sampler yields NaN in pixels where gamma is locally flat:
    # rounding makes g_s >= g_t, c = -expm1(g_s - g_t) <= 0 and sqrt((1-a) c) = NaN (4 of 6144
    # here; none in float64).  The oracle reproduces exactly those NaNs.
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(z_s.numpy()), nan) and nan.sum() <= 8
    tol = 2e-5 if tag == 'f32' else 1e-10     # f32: t**n by pow vs multiply chain, amplified
    assert np.abs(z_s.numpy() - want)[~nan].max() < tol * max(1.0, np.abs(want[~nan]).max()), i
    assert np.abs(g_net.numpy() - G[f'{kind}_{tag}_g_net_{i}']).max() < (1e-5 if tag == 'f32' else 1e-11)
  g0 = O.eval_polynomial(a, b, c, torch.zeros((B, 1), dtype=dtype), cfg).reshape(B, 32, 32, 3)
  x = O.generate_x(tt(inp['z_0']), g0, 256)
  assert np.array_equal(x.numpy(), G[f'{kind}_{tag}_x'])


@pytest.mark.gpu
@pytest.mark.parametrize('kind', ['eps', 'vel'])
def test_cuda_sampler_matches_reference_source(cuda_device, kind):
  from mulan_b200 import model as M
  dev = cuda_device
  inp = sampler_inputs()
  tt = lambda v: torch.from_numpy(np.asarray(v)).to(dev)
  noise = tt(inp['noise'])
  cap = {}

  def score(z, g, cond, det):
    cap['g_net'] = g
    return 0.7 * z + 0.05 * g.reshape(-1, 1, 1, 1) + noise
  cfg = M.VDMConfig(vdm_type='mulan_epsilon' if kind == 'eps' else 'mulan_velocity')
  vdm = M.VDM(cfg, lambda f, d: None, score).to(dev)
  vdm.gamma.load_flax(GI.mlp_weights(SEED + 1000))
  z_t = tt(inp['z_t'])
  for n, i in enumerate(STEPS):
    z_s = M.sample(vdm, i, T, z_t, eps=tt(inp['eps'][n]))
    want32, want64 = G[f'{kind}_f32_z_s_{i}'], G[f'{kind}_f64_z_s_{i}']
    got = z_s.cpu().numpy()
    # elementwise against the reference's f32 result; and no further from exact arithmetic
    # than a few times the reference's own float32 error
    # (where the reference's own float32 evaluation is NaN -- see the oracle test -- the kernel
    # must return the finite value exact arithmetic gives: g_s - g_t is formed from factored
    # power differences and c is clamped at 0)
    ok = ~np.isnan(want32)
    assert np.isfinite(got).all()
    ref = np.abs(want32.astype(np.float64) - want64)[ok].max()
    # per pixel: within 3e-5 of the reference's float32 value, plus that value's own distance
    # from exact arithmetic (pixels where gamma is nearly flat have c ~ 1e-7 with O(1) relative
    # float32 error in the reference; the kernel's c is accurate there)
    own = np.abs(want32.astype(np.float64) - want64)
    lim = 3e-5 * max(1.0, np.abs(want32[ok]).max()) + 2 * own
    assert (np.abs(got - want32)[ok] <= lim[ok]).all(), (i, ref)
    assert np.abs(got - want64)[ok].max() < 4 * ref + 3e-6, (i, ref)
    assert np.abs(got - want64)[~ok].max() < 1e-3 if (~ok).any() else True
    assert np.abs(cap['g_net'].cpu().numpy() - G[f'{kind}_f32_g_net_{i}']).max() < 2e-5
  x = M.generate_x(vdm, tt(inp['z_0']))
  assert x.dtype == torch.uint8
  assert np.array_equal(x.cpu().numpy().astype(np.int64), G[f'{kind}_f64_x'])


@pytest.mark.gpu
def test_sampler_per_row_coefficients_and_loop(cuda_device):
  """abc_rows == rows (conditional sampling) equals the broadcast form when all rows share the
  coefficients; a short sample_fn loop runs and returns uint8 images."""
  from mulan_b200 import model as M, ops
  dev = cuda_device
  cfg = M.VDMConfig()
  vdm = M.VDM(cfg, lambda f, d: None, lambda z, g, c, d: 0.9 * z).to(dev)
  vdm.gamma.load_flax(GI.mlp_weights(9))
  a, b, c = (v.contiguous() for v in
             vdm.gamma._compute_coefficients(M._deterministic_embedding(vdm, 1, dev)))
  Bn = 5
  z = torch.randn(Bn, 3072, device=dev)
  net, eps = torch.randn_like(z), torch.randn_like(z)
  t = torch.full((Bn,), 0.4, device=dev)
  s = torch.full((Bn,), 0.399, device=dev)
  one = ops.sample_step(vdm.desc, a, b, c, t, s, z, net, eps)
  rep = ops.sample_step(vdm.desc, a.repeat(Bn, 1), b.repeat(Bn, 1), c.repeat(Bn, 1), t, s, z, net,
                        eps)
  assert torch.equal(one, rep)
  assert torch.equal(ops.sample_gamma(vdm.desc, a, b, c, t),
                     ops.sample_gamma(vdm.desc, a.repeat(Bn, 1), b.repeat(Bn, 1), c.repeat(Bn, 1), t))
  x = M.sample_fn(vdm, 3, T=8, generator=torch.Generator(device=dev).manual_seed(0))
  assert x.shape == (3, 32, 32, 3) and x.dtype == torch.uint8


ODE_T = [0.03, 0.6]


@pytest.mark.parametrize('tag', ['f32', 'f64'])
@pytest.mark.parametrize('hp', [False, True])
def test_oracle_reverse_ode_matches_reference_source(tag, hp):
  dtype = torch.float32 if tag == 'f32' else torch.float64
  cfg = O.OracleConfig()
  inp = sampler_inputs()
  tt = lambda v: torch.from_numpy(np.asarray(v)).to(dtype)
  W = {k: tt(v) for k, v in GI.mlp_weights(SEED + 1000).items()}
  a, b, c = O.compute_coefficients(W, O.deterministic_embedding(B, cfg, dtype))
  t = tt(np.asarray(ODE_T, np.float32)).reshape(B, 1)
  g_t = O.eval_polynomial(a, b, c, t, cfg).reshape(B, 32, 32, 3)
  w = O.eval_polynomial_dt(a, b, c, t, cfg).reshape(B, 32, 32, 3)
  x = tt(inp['z_t'])
  net = 0.7 * x + 0.05 * O.score_model_gt(g_t, cfg).reshape(-1, 1, 1, 1) + tt(inp['noise'])
  drift = O.reverse_ode(x, g_t, w, net, hp)
  want = G[f'eps_{tag}_ode_drift_hp{int(hp)}']
  tol = (2e-5 if tag == 'f32' else 1e-10) * np.abs(want).max()
  assert np.abs(drift.numpy() - want).max() < tol


@pytest.mark.gpu
@pytest.mark.parametrize('hp', [False, True])
def test_cuda_reverse_ode_and_divergence(cuda_device, hp):
  """Drift against the golden from the reference's reverse_ode; Hutchinson divergence
  (drift pieces + autograd through a torch denoiser) against the oracle's autograd."""
  from mulan_b200 import model as M
  dev = cuda_device
  inp = sampler_inputs()
  tt = lambda v: torch.from_numpy(np.asarray(v)).to(dev)
  noise = tt(inp['noise'])
  w1 = torch.tensor(0.7, device=dev)
  conv = torch.nn.Conv2d(3, 3, 3, padding=1).to(dev)

  def score_lin(z, g, cond, det):
    return w1 * z + 0.05 * g.reshape(-1, 1, 1, 1) + noise
  cfg = M.VDMConfig()
  vdm = M.VDM(cfg, lambda f, d: None, score_lin).to(dev)
  vdm.gamma.load_flax(GI.mlp_weights(SEED + 1000))
  emb = M._deterministic_embedding(vdm, B, dev)
  t = tt(np.asarray(ODE_T, np.float32))
  x = tt(inp['z_t'])
  rng = np.random.default_rng(4)
  v = tt((rng.integers(0, 2, (B, 32, 32, 3)) * 2 - 1).astype(np.float32))   # Rademacher
  drift, div = M.value_div_fn(vdm, x, emb, t, v, high_precision=hp)
  want = G[f'eps_f32_ode_drift_hp{int(hp)}']
  assert np.abs(drift.cpu().numpy() - want).max() < 3e-5 * np.abs(want).max()
  # divergence with a non-trivial denoiser Jacobian (a conv), against oracle autograd in f64
  vdm.score_model = lambda z, g, cond, det: (
      w1 * z + 0.3 * conv(z.permute(0, 3, 1, 2)).permute(0, 2, 3, 1) + noise)
  drift, div = M.value_div_fn(vdm, x, emb, t, v, high_precision=hp)
  ocfg = O.OracleConfig()
  Wd = {k: torch.from_numpy(val).double() for k, val in GI.mlp_weights(SEED + 1000).items()}
  a, b, c = O.compute_coefficients(Wd, O.deterministic_embedding(B, ocfg, torch.float64))
  td = t.cpu().double().reshape(B, 1)
  g_t = O.eval_polynomial(a, b, c, td, ocfg).reshape(B, 32, 32, 3)
  wd = O.eval_polynomial_dt(a, b, c, td, ocfg).reshape(B, 32, 32, 3)
  conv64 = torch.nn.Conv2d(3, 3, 3, padding=1).double()
  conv64.load_state_dict({k: val.cpu().double() for k, val in conv.state_dict().items()})
  score64 = lambda z: (0.7 * z + 0.3 * conv64(z.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
                       + noise.cpu().double())
  f64, div64 = O.value_div(x.cpu().double(), g_t, wd, score64, v.cpu().double(), hp)
  assert np.abs(drift.cpu().numpy() - f64.numpy()).max() < 1e-4 * f64.abs().max().item()
  assert np.abs(div.cpu().numpy() - div64.numpy()).max() < 1e-4 * div64.abs().max().item()


@pytest.mark.gpu
def test_apply_gamma(cuda_device):
  """VDM.apply_gamma (ldm/model_mulan_epsilon.py:182-193): per-pixel gamma(embedding, t);
  fixed ends at t = 0 / 1, oracle in between."""
  from mulan_b200 import model as M
  dev = cuda_device
  vdm = M.VDM(M.VDMConfig(), lambda f, d: None, lambda *a: None).to(dev)
  W = GI.mlp_weights(12)
  vdm.gamma.load_flax(W)
  t = torch.tensor([0.0, 1.0, 0.37], device=dev)
  emb = M._deterministic_embedding(vdm, 3, dev)
  g = vdm.apply_gamma(t, emb).cpu()
  assert g.shape == (3, 3072)
  assert torch.all(g[0] == torch.tensor(-13.3, dtype=torch.float32))
  assert (g[1] - 5.0).abs().max().item() < 5e-6
  ocfg = O.OracleConfig()
  Wd = {k: torch.from_numpy(v).double() for k, v in W.items()}
  a, b, c = O.compute_coefficients(Wd, O.deterministic_embedding(3, ocfg, torch.float64))
  want = O.eval_polynomial(a, b, c, t.cpu().double().reshape(3, 1), ocfg)
  assert (g.double() - want).abs().max().item() < 1e-4


@pytest.mark.gpu
def test_cuda_sampler_ieee_variant(cuda_device):
  """MULAN_SAMPLER_IEEE=1 selects the op-for-op form of the ancestral step (IEEE division /
  sqrt / expm1); it must pass the same golden test as the default fast form.  The switch is
  read once per process, hence the child process."""
  import subprocess
  root = os.path.dirname(HERE)
  out = subprocess.run(
      [sys.executable, '-m', 'pytest', os.path.join(HERE, 'test_sampler.py'), '-q', '-m', 'gpu',
       '-k', 'test_cuda_sampler_matches_reference_source or per_row_coefficients'],
      cwd=root, env=dict(os.environ, MULAN_SAMPLER_IEEE='1'), capture_output=True, text=True,
      timeout=900)
  assert out.returncode == 0 and '3 passed' in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize('kind', ['eps', 'vel'])
def test_sampler_broadcast_factor_table(cuda_device, kind):
  """abc_rows == 1 runs a persistent kernel that caches the step's factors per CTA while
  consecutive rows share (t, s).  With more rows than resident CTAs every path is taken (fill on
  the first row of a run, table hits, rows whose times differ); each must be bit-identical to
  the per-example-coefficient kernel."""
  from mulan_b200 import model as M, ops
  dev = cuda_device
  cfg = M.VDMConfig(vdm_type='mulan_epsilon' if kind == 'eps' else 'mulan_velocity')
  vdm = M.VDM(cfg, lambda f, d: None, lambda z, g, c, d: z).to(dev)
  vdm.gamma.load_flax(GI.mlp_weights(9))
  a, b, c = (v.contiguous() for v in
             vdm.gamma._compute_coefficients(M._deterministic_embedding(vdm, 1, dev)))
  Bn = 3001
  gen = torch.Generator(device=dev).manual_seed(1)
  z, net, eps = (torch.randn(Bn, 3072, device=dev, generator=gen) for _ in range(3))
  A, Bc, Cc = a.repeat(Bn, 1), b.repeat(Bn, 1), c.repeat(Bn, 1)
  uniform = torch.full((Bn,), 0.4, device=dev)
  halves = torch.where(torch.arange(Bn, device=dev) < 1700, 0.4, 0.7).float()
  random_t = torch.rand(Bn, device=dev, generator=gen) * 0.9 + 0.05
  for t in (uniform, halves, random_t):
    s = t - 1.0 / 1000
    one = ops.sample_step(vdm.desc, a, b, c, t, s, z, net, eps)
    rep = ops.sample_step(vdm.desc, A, Bc, Cc, t, s, z, net, eps)
    assert torch.equal(one, rep)
    assert torch.isfinite(one).all()


@pytest.mark.gpu
def test_conditional_sample_equals_sample_for_the_deterministic_embedding(cuda_device):
  """VDM.conditional_sample with every example's embedding set to the deterministic one is
  VDM.sample (ldm/model_mulan_epsilon.py:377-438): per-example coefficients vs one broadcast row."""
  from mulan_b200 import model as M
  dev = cuda_device
  for kind in ('mulan_epsilon', 'mulan_velocity'):
    cfg = M.VDMConfig(vdm_type=kind)
    vdm = M.VDM(cfg, lambda f, d: None,
                lambda z, g, c, d: 0.8 * z + 0.01 * g.reshape(-1, 1, 1, 1) + 0.001 * c.sum(1).reshape(-1, 1, 1, 1)).to(dev)
    vdm.gamma.load_flax(GI.mlp_weights(12))
    Bn = 7
    gen = torch.Generator(device=dev).manual_seed(3)
    z = torch.randn(Bn, 32, 32, 3, device=dev, generator=gen)
    eps = torch.randn(Bn, 32, 32, 3, device=dev, generator=gen)
    emb = M._deterministic_embedding(vdm, Bn, dev)
    c1 = tuple(v.contiguous() for v in
               vdm.gamma._compute_coefficients(M._deterministic_embedding(vdm, 1, dev)))
    one = M.sample(vdm, 400, 1000, z, eps=eps, coeffs=c1)
    two = M.conditional_sample(vdm, 400, 1000, z, emb, eps=eps,
                               coeffs=tuple(v.repeat(Bn, 1) for v in c1))
    assert torch.equal(one, two)
    # coefficients from the batch-7 GEMM instead of the batch-1 GEMM: cuBLAS's summation order
    # differs, and gamma amplifies a 1e-7 change of (a, b, c) where P/S cancels
    three = M.conditional_sample(vdm, 400, 1000, z, emb, eps=eps)
    assert (three - one).norm().item() < 1e-4 * one.norm().item()
    # a different embedding changes the schedule, hence the step
    emb2 = emb.clone()
    emb2[:, :5] = 0.0
    assert not torch.allclose(M.conditional_sample(vdm, 400, 1000, z, emb2, eps=eps), one, rtol=1e-3)
