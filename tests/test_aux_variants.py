"""The auxiliary-latent variants the shipped configs do not select:
latent_type 'gumbel' / 'gaussian' and topk_noise_type 'gumbel'
(ldm/model_mulan_epsilon.py:195-219, :238-239, :264-270).

CPU: the oracle against tests/golden/latent.npz, which was produced by executing the reference's
own `_get_embedding_and_kl_z` (tests/golden/make_golden_latent.py).
GPU: the CUDA kernels, through the C ABI, against the same goldens and against the oracle's
autograd on larger seeded inputs.
"""
import math
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mulan_oracle as O  # noqa: E402

GOLD = np.load(os.path.join(ROOT, 'tests', 'golden', 'latent.npz'))
SEED, B, L, K = 401, 6, 50, 15
CASES = ('topk_gumbel', 'gumbel_step0', 'gumbel_step30000', 'gaussian')


def latent_inputs():
  """Same draws as tests/golden/make_golden_latent.py:latent_inputs."""
  r = np.random.default_rng(SEED)
  f32 = np.float32
  return dict(logits=(2.0 * r.standard_normal((B, L))).astype(f32),
              gumbel=r.gumbel(size=(B, L)).astype(f32),
              mu=r.standard_normal((B, L)).astype(f32),
              var=np.logaddexp(r.standard_normal((B, L)), 0.0).astype(f32),
              eps_z=r.standard_normal((B, L)).astype(f32),
              emb_bar=r.standard_normal((B, L)).astype(f32),
              kl_bar=r.standard_normal((B,)).astype(f32))


def tau_of(case):
  return max(0.5, math.exp(-0.00001 * (30000.0 if case.endswith('30000') else 0.0)))


def run_case(case, fns, dtype, device='cpu'):
  """fns = (topk_add, gumbel, gaussian) callables; returns emb, kl and the leaf gradients."""
  inp = {k: torch.from_numpy(v).to(dtype).to(device) for k, v in latent_inputs().items()}
  topk_add, gumbel, gaussian = fns
  if case == 'gaussian':
    leaves = {'mu': inp['mu'].requires_grad_(True), 'var': inp['var'].requires_grad_(True)}
    emb, kl = gaussian(leaves['mu'], leaves['var'], inp['eps_z'])
  else:
    leaves = {'logits': inp['logits'].requires_grad_(True)}
    if case == 'topk_gumbel':
      emb, kl = topk_add(leaves['logits'], inp['gumbel'], K)
    else:
      emb, kl = gumbel(leaves['logits'], inp['gumbel'], tau_of(case))
  loss = (emb * inp['emb_bar']).sum() + (kl * inp['kl_bar']).sum()
  grads = torch.autograd.grad(loss, list(leaves.values()))
  out = {'emb': emb.detach(), 'kl': kl.detach()}
  for n, g in zip(leaves, grads):
    out['grad_' + n] = g
  return {k: v.cpu().numpy() for k, v in out.items()}


ORACLE_FNS = (lambda l, n, k: O.topk_add_embedding_and_loss(l, n, k, L),
              lambda l, n, tau: O.gumbel_embedding_and_loss(l, n, tau, L),
              O.gaussian_embedding_and_loss)


def rel(got, want):
  return float(np.max(np.abs(got.astype(np.float64) - want) / (np.abs(want) + 1e-3)))


@pytest.mark.parametrize('case', CASES)
def test_oracle_matches_reference_latent_golden(case):
  """f64 oracle == f64 reference to rounding; f32 oracle within f32 rounding of it."""
  for dtype, tag, tol in ((torch.float64, 'f64', 1e-12), (torch.float32, 'f32', 2e-5)):
    got = run_case(case, ORACLE_FNS, dtype)
    for k, v in got.items():
      want = GOLD[f'{case}_{tag}_{k}']
      assert v.shape == want.shape
      assert rel(v, want.astype(np.float64)) < tol, (case, tag, k)


def test_latent_golden_hard_part_is_the_forward_value():
  """Straight-through: the embedding's VALUE is the hard one-hot / k-hot vector."""
  e = GOLD['gumbel_step0_f64_emb']
  assert np.allclose(e.sum(1), 1.0) and np.allclose(np.sort(e, 1)[:, -1], 1.0)
  e = GOLD['topk_gumbel_f64_emb']
  assert np.allclose(e.sum(1), K)


def test_oracle_deterministic_embedding_variants():
  cfg = O.OracleConfig()
  assert O.deterministic_embedding(2, cfg).sum().item() == 2 * cfg.latent_k
  g = O.deterministic_embedding(2, cfg, latent_type='gumbel')
  assert g.sum().item() == 2 and g[0, 1].item() == 1.0
  assert O.deterministic_embedding(2, cfg, latent_type='gaussian').abs().sum().item() == 0.0


# ------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------

def cuda_fns():
  from mulan_b200 import ops
  return (ops.aux_topk_add, ops.aux_gumbel, ops.aux_gaussian)


@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES)
def test_cuda_latent_variants_match_reference_golden(case, cuda_device):
  got = run_case(case, cuda_fns(), torch.float32, 'cuda')
  for k, v in got.items():
    want64 = GOLD[f'{case}_f64_{k}']
    ref_err = rel(GOLD[f'{case}_f32_{k}'], want64)
    # 1e-5 rel on values, 1e-4 rel on gradients (north_star), or the reference's own f32 error
    tol = max(1e-4 if k.startswith('grad') else 1e-5, 4 * ref_err)
    assert rel(v, want64) < tol, (case, k, rel(v, want64), ref_err)


@pytest.mark.gpu
@pytest.mark.parametrize('rows', [1, 37, 1024])
def test_cuda_latent_variants_match_oracle_autograd(rows, cuda_device):
  g = torch.Generator().manual_seed(7 + rows)
  logits = 3.0 * torch.randn(rows, L, generator=g)
  noise = -torch.log(-torch.log(torch.rand(rows, L, generator=g).clamp_min(1e-20)))
  mu = torch.randn(rows, L, generator=g)
  var = torch.nn.functional.softplus(torch.randn(rows, L, generator=g))
  eps_z = torch.randn(rows, L, generator=g)
  eb = torch.randn(rows, L, generator=g)
  kb = torch.randn(rows, generator=g)
  topk_add, gumbel, gaussian = cuda_fns()

  def both(fn_cuda, fn_orc, leaves, consts):
    outs = []
    for fn, dev, dt in ((fn_cuda, 'cuda', torch.float32), (fn_orc, 'cpu', torch.float64)):
      lv = [v.to(dt).to(dev).requires_grad_(True) for v in leaves]
      cs = [v.to(dt).to(dev) if torch.is_tensor(v) else v for v in consts]
      emb, kl = fn(*lv, *cs)
      loss = (emb * eb.to(dt).to(dev)).sum() + (kl * kb.to(dt).to(dev)).sum()
      gr = torch.autograd.grad(loss, lv)
      outs.append([emb.detach().cpu().double(), kl.detach().cpu().double()] +
                  [v.cpu().double() for v in gr])
    return outs

  cases = [
      (topk_add, ORACLE_FNS[0], [logits], [noise, K]),
      (gumbel, ORACLE_FNS[1], [logits], [noise, 0.5]),
      (gumbel, ORACLE_FNS[1], [logits], [noise, 0.8187]),
      (gaussian, ORACLE_FNS[2], [mu, var], [eps_z]),
  ]
  for fc, fo, leaves, consts in cases:
    got, want = both(fc, fo, leaves, consts)
    for i, (gv, wv) in enumerate(zip(got, want)):
      tol = 1e-5 if i < 2 else 1e-4
      # elementwise, relative to the element or (where a sum of O(1) terms cancels) to 1% of
      # the tensor's largest element: the float32 rounding of the terms is the floor
      err = ((gv - wv).abs() / (wv.abs() + 0.01 * wv.abs().max())).max().item()
      assert err < tol, (fc.__name__, i, err)


@pytest.mark.gpu
def test_cuda_latent_variants_none_cotangents(cuda_device):
  """Only one of (embedding, kl_z) reaching the loss must still give the right gradient."""
  topk_add, gumbel, gaussian = cuda_fns()
  g = torch.Generator().manual_seed(3)
  logits = torch.randn(9, L, generator=g)
  noise = torch.randn(9, L, generator=g)
  for fc, fo, extra in ((topk_add, ORACLE_FNS[0], K), (gumbel, ORACLE_FNS[1], 0.7)):
    for pick in (0, 1):
      lc = logits.cuda().requires_grad_(True)
      lo = logits.double().requires_grad_(True)
      fc(lc, noise.cuda(), extra)[pick].sum().backward()
      fo(lo, noise.double(), extra)[pick].sum().backward()
      assert torch.allclose(lc.grad.cpu().double(), lo.grad, rtol=1e-4, atol=1e-6), (fc, pick)


@pytest.mark.gpu
@pytest.mark.parametrize('latent_type,noise_type', [('gumbel', 'gamma'), ('gaussian', 'gamma'),
                                                    ('topk', 'gumbel')])
def test_vdm_call_other_latent_types(latent_type, noise_type, cuda_device):
  """VDM.__call__ with the other latent types == the oracle's vdm_call (f64) with the matching
  latent_fn; encoder / denoiser are small closed-form stand-ins shared by both sides."""
  from mulan_b200.model import VDM, VDMConfig, loss_fn
  n, step = 8, 20000
  g = torch.Generator().manual_seed(11)
  We = 0.05 * torch.randn(3072, L, generator=g)
  We2 = 0.05 * torch.randn(3072, L, generator=g)
  Wn = 0.02 * torch.randn(L, 3072, generator=g)

  def make(dev, dt):
    we, we2, wn = (v.to(dt).to(dev).requires_grad_(True) for v in (We, We2, Wn))

    def encoder(orig_f, deterministic=True):
      h = orig_f.reshape(orig_f.shape[0], -1)
      if latent_type == 'gaussian':
        return h @ we, torch.nn.functional.softplus(h @ we2)
      return h @ we

    def score(z, g_t, cond, deterministic=True):
      return 0.5 * z + (cond @ wn).reshape(z.shape) + 0.01 * g_t.reshape(-1, 1, 1, 1)

    return encoder, score, (we, we2, wn)

  enc, score, leaves = make('cuda', torch.float32)
  model = VDM(VDMConfig(latent_type=latent_type, topk_noise_type=noise_type), enc, score).cuda()
  torch.manual_seed(3)
  for name in ('dense_out_a', 'dense_out_b'):        # zero-init in the reference; make a, b live
    torch.nn.init.normal_(getattr(model.gamma, name).weight, std=0.05)
  images = torch.randint(0, 256, (n, 32, 32, 3), dtype=torch.int32, device='cuda')
  draws = model.make_draws(n, 'cuda', torch.Generator(device='cuda').manual_seed(5))
  assert draws['G'].shape == (n, L)
  out = model(images, step=step, deterministic=False, draws=draws)
  bpd, _ = loss_fn(model, {'images': images}, step=step, is_train=True, draws=draws)
  used = [v for v in leaves if latent_type == 'gaussian' or v is not leaves[1]]
  grads = torch.autograd.grad(bpd, used)

  head = model.gamma.state_dict()
  tau = max(0.5, math.exp(-0.00001 * step))

  def oracle(dt):
    enc_o, score_o, leaves_o = make('cpu', dt)
    params = {k.replace('.weight', '/kernel').replace('.bias', '/bias'):
              (v.T if v.dim() == 2 else v).detach().cpu().to(dt) for k, v in head.items()}
    latent_fn = {
        'gumbel': lambda f, G: O.gumbel_embedding_and_loss(enc_o(f), G, tau, L),
        'gaussian': lambda f, G: O.gaussian_embedding_and_loss(*enc_o(f), G),
        'topk': lambda f, G: O.topk_add_embedding_and_loss(enc_o(f), G, K, L),
    }[latent_type]
    d = {k: v.detach().cpu().to(dt) for k, v in draws.items()}
    res = O.vdm_call(images.cpu(), d, lambda e: O.compute_coefficients(params, e), None,
                     score_o, O.MODE_EPS, O.OracleConfig(), dtype=dt, latent_fn=latent_fn)
    bpd_o, _ = O.loss_fn_bpd(res)
    used_o = [v for v in leaves_o if latent_type == 'gaussian' or v is not leaves_o[1]]
    return res, bpd_o, torch.autograd.grad(bpd_o, used_o)

  want, bpd64, grads64 = oracle(torch.float64)
  ref, _, grads32 = oracle(torch.float32)            # what the reference's dtype itself gives
  relmax = lambda u, v: ((u.detach().double() - v.detach()).abs() / v.detach().abs()).max().item()
  for name in ('loss_recon', 'loss_klz', 'loss_diff'):
    wv = getattr(want, name)
    ref_err = relmax(getattr(ref, name), wv)
    assert relmax(getattr(out, name).cpu(), wv) < 1e-5 + ref_err, (name, ref_err)
  assert abs(bpd.item() - bpd64.item()) < 1e-4
  for gv, wv, rv in zip(grads, grads64, grads32):
    ref_err = ((rv.double() - wv).norm() / wv.norm()).item()
    err = ((gv.cpu().double() - wv).norm() / wv.norm()).item()
    assert err < 1e-4 + ref_err, (err, ref_err)
