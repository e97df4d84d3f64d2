"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs.  Tolerances are the north_star's: per-example loss terms 1e-5 relative, gradients
1e-4 relative (per-example L2), BPD 1e-4 bits/dim, all float32.
"""
import math

import numpy as np
import pytest
import torch

from oracle import mulan_oracle as O

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-4
BPD_ATOL = 1e-4

MODES = {'eps': O.MODE_EPS, 'vel': O.MODE_VEL, 'vel_from_eps': O.MODE_VEL_FROM_EPS}


def _ops():
  from mulan_b200 import ops
  return ops


def _dev(d, device):
  return {k: v.to(device).contiguous() for k, v in d.items()}


def _rel(got, want):
  got, want = got.double().cpu(), want.double().cpu()
  return ((got - want).abs() / want.abs().clamp_min(1e-30)).max().item()


def _rel_l2_rows(got, want):
  got, want = got.double().cpu(), want.double().cpu()
  num = (got - want).flatten(1).norm(dim=1)
  den = want.flatten(1).norm(dim=1).clamp_min(1e-30)
  return (num / den).max().item()


def _oracle(inp, mode, cfg=None, dtype=torch.float32, gt='vdm'):
  cfg = cfg or O.OracleConfig(unet_type=gt)
  cast = lambda v: v.to(dtype) if v.is_floating_point() else v
  i = {k: cast(v) for k, v in inp.items()}
  return O.elbo_terms(i['x'], i['a'], i['b'], i['c'], i['t'], i['eps_0'], i['eps'],
                      lambda z, g: i['net'], mode, cfg, dtype=dtype, return_aux=True)


@pytest.mark.parametrize('B,seed', [(8, 0), (1, 1), (2, 2), (127, 3), (128, 4)])
def test_fwd_pre_parity(cuda_device, B, seed):
  ops = _ops()
  inp = O.synth_inputs(B, seed)
  out, aux = _oracle(inp, O.MODE_EPS)
  g = _dev(inp, cuda_device)
  desc = ops.Desc()
  r = ops.fwd_pre(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'], g['eps'])
  assert _rel(r['loss_recon'], out.loss_recon) < LOSS_RTOL
  assert _rel(r['loss_klz_prior'], aux['loss_klz_prior']) < LOSS_RTOL
  # z_t elementwise: float32 cancellation in P/S makes single pixels of the float32 reference
  # itself differ from exact arithmetic by ~1e-5..1e-4; bound the CUDA error by the oracle's.
  _, aux64 = _oracle(inp, O.MODE_EPS, dtype=torch.float64)
  z64 = aux64['z_t'].reshape(B, -1)
  err_ref = (aux['z_t'].reshape(B, -1).double() - z64).abs().max().item()
  err_got = (r['z_t'].cpu().double() - z64).abs().max().item()
  assert err_got <= 4 * err_ref + 2e-6, (err_got, err_ref)
  assert _rel_l2_rows(r['z_t'], aux['z_t'].reshape(B, -1)) < 1e-5
  g_mean = aux['g_t'].reshape(B, -1).mean(dim=1)
  assert (r['g_net'].cpu() - g_mean).abs().max().item() < 2e-5
  # d gamma/dt: q^2 form vs the reference's expanded jvp -> compare per-example sums
  w_ref = aux['g_t_grad'].reshape(B, -1)
  assert _rel(r['w'].sum(dim=1), w_ref.sum(dim=1)) < 1e-5
  assert _rel_l2_rows(r['w'], w_ref) < 1e-5
  var0 = r['var_sums'][:, 0].sum().item() / (B * 3072)
  var1 = r['var_sums'][:, 1].sum().item() / (B * 3072)
  assert abs(var0 - out.var_0.item()) < 1e-6 * out.var_0.item() + 1e-12
  assert abs(var1 - out.var_1.item()) < 1e-6


@pytest.mark.parametrize('mode', list(MODES))
@pytest.mark.parametrize('B,seed', [(8, 0), (128, 5)])
def test_fwd_post_parity(cuda_device, mode, B, seed):
  ops = _ops()
  inp = O.synth_inputs(B, seed)
  out, aux = _oracle(inp, MODES[mode])
  g = _dev(inp, cuda_device)
  desc = ops.Desc(param=MODES[mode])
  got = ops.fwd_post(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'], None)
  assert _rel(got, out.loss_diff) < LOSS_RTOL
  if mode == 'eps':
    pre = ops.fwd_pre(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'], g['eps'])
    got_w = ops.fwd_post(desc, None, None, None, None, None, g['eps'], g['net'], pre['w'])
    assert _rel(got_w, out.loss_diff) < LOSS_RTOL


def test_fwd_gt_pixel(cuda_device):
  ops = _ops()
  B = 4
  inp = O.synth_inputs(B, 11)
  out, aux = _oracle(inp, O.MODE_EPS, gt='ldm')
  g = _dev(inp, cuda_device)
  desc = ops.Desc(gt_mode=1)
  r = ops.fwd_pre(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'], g['eps'],
                  save_w=False)
  assert r['w'] is None
  assert r['g_net'].shape == (B, 3072)
  # per-pixel gamma_t: bound the error against float64 by the float32 oracle's own error
  _, aux64 = _oracle(inp, O.MODE_EPS, dtype=torch.float64, gt='ldm')
  g64 = aux64['g_t'].reshape(B, -1)
  err_ref = (aux['g_t'].reshape(B, -1).double() - g64).abs().max().item()
  err_got = (r['g_net'].cpu().double() - g64).abs().max().item()
  assert err_got <= 4 * err_ref + 4e-6, (err_got, err_ref)
  assert _rel_l2_rows(r['g_net'], aux['g_t'].reshape(B, -1)) < 1e-5
  assert _rel(r['loss_recon'], out.loss_recon) < LOSS_RTOL


def _oracle_grads(inp, mode, gL, zbar, gbar, dtype, gt='vdm'):
  """Cotangents of (a,b,c,net) for L = sum gL*loss_diff + <zbar, z_t> + <gbar, g_net>."""
  cfg = O.OracleConfig(unet_type=gt)
  cast = lambda v: v.to(dtype) if v.is_floating_point() else v
  i = {k: cast(v) for k, v in inp.items()}
  a, b, c, net = (i[k].clone().requires_grad_(True) for k in ('a', 'b', 'c', 'net'))
  out, aux = O.elbo_terms(i['x'], a, b, c, i['t'], i['eps_0'], i['eps'], lambda z, g: net,
                          mode, cfg, dtype=dtype, return_aux=True)
  B = a.shape[0]
  L = (gL.to(dtype) * out.loss_diff).sum()
  L = L + (zbar.to(dtype) * aux['z_t'].reshape(B, -1)).sum()
  L = L + (gbar.to(dtype) * O.score_model_gt(aux['g_t'], cfg).reshape(gbar.shape)).sum()
  # recon and prior KL are part of the loss too; their (a,b,c) gradient is rounding noise
  L = L + (out.loss_recon + out.loss_klz).sum() * float(gL.mean())
  return torch.autograd.grad(L, [a, b, c, net])


@pytest.mark.parametrize('mode', list(MODES))
@pytest.mark.parametrize('gt', ['vdm', 'ldm'])
def test_backward_parity(cuda_device, mode, gt):
  ops = _ops()
  B = 8
  inp = O.synth_inputs(B, 21)
  rng = np.random.default_rng(99)
  gL = torch.from_numpy(rng.uniform(0.5, 1.5, B).astype(np.float32)) / (B * 3072 * math.log(2))
  zbar = torch.from_numpy(rng.standard_normal((B, 3072)).astype(np.float32)) * 1e-4
  gbar = torch.from_numpy(
      rng.standard_normal((B,) if gt == 'vdm' else (B, 3072)).astype(np.float32)) * 1e-3
  want = _oracle_grads(inp, MODES[mode], gL, zbar, gbar, torch.float64, gt)
  want32 = _oracle_grads(inp, MODES[mode], gL, zbar, gbar, torch.float32, gt)
  g = _dev(inp, cuda_device)
  desc = ops.Desc(param=MODES[mode], gt_mode=0 if gt == 'vdm' else 1)
  gLd, zbd, gbd = gL.to(cuda_device), zbar.to(cuda_device), gbar.to(cuda_device)
  n_bar = ops.bwd_post(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'], None,
                       gLd)
  ab, bb, cb = ops.bwd_pre(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'],
                           zbd, gbd, gLd)
  for got, w64, w32, name in ((ab, want[0], want32[0], 'a'), (bb, want[1], want32[1], 'b'),
                              (cb, want[2], want32[2], 'c'), (n_bar, want[3], want32[3], 'n')):
    err = _rel_l2_rows(got, w64)
    err32 = _rel_l2_rows(w32, w64)
    assert err < GRAD_RTOL, f'{name}_bar: rel l2 {err:.3e} (f32 oracle itself {err32:.3e})'
    assert _rel_l2_rows(got, w32) < GRAD_RTOL
  if mode == 'eps':
    pre = ops.fwd_pre(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'], g['eps'])
    n2 = ops.bwd_post(desc, None, None, None, None, None, g['eps'], g['net'], pre['w'], gLd)
    assert _rel_l2_rows(n2, want[3]) < GRAD_RTOL


def test_bwd_pre_optional_inputs(cuda_device):
  """NULL z_bar / g_bar / gL mean zero cotangents."""
  ops = _ops()
  B = 4
  inp = O.synth_inputs(B, 31)
  g = _dev(inp, cuda_device)
  desc = ops.Desc()
  z = torch.zeros((B, 3072), device=cuda_device)
  zero_g = torch.zeros((B,), device=cuda_device)
  gL = torch.full((B,), 1e-3, device=cuda_device)
  full = ops.bwd_pre(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'], z,
                     zero_g, gL)
  part = ops.bwd_pre(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'], None,
                     None, gL)
  for u, v in zip(full, part):
    assert torch.equal(u, v)
  none = ops.bwd_pre(desc, None, g['a'], g['b'], g['c'], g['t'], None, None, None, None, None)
  for u in none:
    assert torch.count_nonzero(u).item() == 0


def test_edge_cases(cuda_device):
  """t at the ends, a == 0 (the reference's zero-init head), large |a|,|b|, c -> 1e-3,
  x at the vocabulary edges with large eps_0."""
  ops = _ops()
  B = 8
  inp = O.synth_inputs(B, 41)
  inp['t'] = torch.tensor([0.0, 1.0, 1e-6, 1 - 1e-6, 0.5, 0.25, 0.999, 0.001])
  inp['a'][0] = 0.0
  inp['a'][1] = 0.0
  inp['a'][2] *= 30.0
  inp['b'][3] *= 30.0
  inp['c'][4] = 1e-3
  inp['x'][5] = 0
  inp['x'][6] = 255
  inp['eps_0'][5] = inp['eps_0'][5] * 3.0
  inp['eps_0'][6] = inp['eps_0'][6] * 3.0
  for mode in MODES:
    out, aux = _oracle(inp, MODES[mode])
    g = _dev(inp, cuda_device)
    desc = ops.Desc(param=MODES[mode])
    r = ops.fwd_pre(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'], g['eps'])
    assert _rel(r['loss_recon'], out.loss_recon) < LOSS_RTOL
    assert _rel(r['loss_klz_prior'], aux['loss_klz_prior']) < LOSS_RTOL
    got = ops.fwd_post(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'], None)
    # rows with t == 0 have loss_diff ~ c^2-weighted; all rows must match
    assert _rel(got, out.loss_diff) < LOSS_RTOL, mode
    assert torch.isfinite(r['z_t']).all()


def test_rows_zero_and_errors(cuda_device):
  ops = _ops()
  from mulan_b200._lib import MulanError
  desc = ops.Desc()
  e = lambda *s, dt=torch.float32: torch.empty(s, dtype=dt, device=cuda_device)
  r = ops.fwd_pre(desc, e(0, 3072, dt=torch.uint8), e(0, 3072), e(0, 3072), e(0, 3072), e(0),
                  e(0, 3072), e(0, 3072))
  assert r['loss_recon'].shape == (0,)
  with pytest.raises(TypeError):
    ops.fwd_pre(desc, torch.zeros(2, 3072, dtype=torch.uint8), torch.zeros(2, 3072),
                torch.zeros(2, 3072), torch.zeros(2, 3072), torch.zeros(2), torch.zeros(2, 3072),
                torch.zeros(2, 3072))
  bad = ops.Desc(dim=3070)
  with pytest.raises(MulanError) as ei:
    ops.fwd_pre(bad, e(2, 3070, dt=torch.uint8), e(2, 3070), e(2, 3070), e(2, 3070), e(2),
                e(2, 3070), e(2, 3070))
  assert ei.value.status == -2
  with pytest.raises(MulanError) as ei:
    ops.fwd_pre(ops.Desc(n_timesteps=10, param=1), e(2, 3072, dt=torch.uint8), e(2, 3072),
                e(2, 3072), e(2, 3072), e(2), e(2, 3072), e(2, 3072))
  assert ei.value.status == -3
  with pytest.raises(MulanError) as ei:   # discrete time needs the saved weight
    ops.fwd_pre(ops.Desc(n_timesteps=10), e(2, 3072, dt=torch.uint8), e(2, 3072), e(2, 3072),
                e(2, 3072), e(2), e(2, 3072), e(2, 3072), save_w=False)
  assert ei.value.status == -1
  # misaligned float pointer
  buf = e(2 * 3072 + 1)
  mis = buf[1:].view(2, 3072)
  with pytest.raises(MulanError) as ei:
    ops.fwd_pre(desc, e(2, 3072, dt=torch.uint8), mis, e(2, 3072), e(2, 3072), e(2), e(2, 3072),
                e(2, 3072))
  assert ei.value.status == -2


@pytest.mark.parametrize('T', [10, 1000])
def test_discrete_time_epsilon(cuda_device, T):
  """sm_n_timesteps > 0 (ldm/model_mulan_epsilon.py:348-355): loss and gradients."""
  ops = _ops()
  B = 8
  cfg = O.OracleConfig(sm_n_timesteps=T)
  inp = O.synth_inputs(B, 71)
  inp['t'] = O.sample_t(0.321, B, cfg)
  rng = np.random.default_rng(5)
  gL = torch.from_numpy(rng.uniform(0.5, 1.5, B).astype(np.float32)) / (B * 3072 * math.log(2))
  res = {}
  for dtype in (torch.float32, torch.float64):
    cast = lambda v: v.to(dtype) if v.is_floating_point() else v
    i = {k: cast(v) for k, v in inp.items()}
    a, b, c, net = (i[k].clone().requires_grad_(True) for k in ('a', 'b', 'c', 'net'))
    out = O.elbo_terms(i['x'], a, b, c, i['t'], i['eps_0'], i['eps'], lambda z, g: net,
                       O.MODE_EPS, cfg, dtype=dtype)
    grads = torch.autograd.grad((gL.to(dtype) * out.loss_diff).sum(), [a, b, c, net])
    res[dtype] = (out.loss_diff.detach(), grads)
  g = _dev(inp, cuda_device)
  desc = ops.Desc(n_timesteps=T)
  gLd = gL.to(cuda_device)
  pre = ops.fwd_pre(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'], g['eps'])
  diff = ops.fwd_post(desc, None, None, None, None, None, g['eps'], g['net'], pre['w'])
  # float32 cancellation in g_t - g_s: hold the CUDA loss to the float32 oracle at 1e-5 when
  # the oracle itself is that close to float64, else to a multiple of the oracle's own error
  l32, l64 = res[torch.float32][0], res[torch.float64][0]
  ref_err = _rel(l32, l64)
  assert _rel(diff, l64) < max(LOSS_RTOL, 4 * ref_err), (_rel(diff, l64), ref_err)
  n_bar = ops.bwd_post(desc, None, None, None, None, None, g['eps'], g['net'], pre['w'], gLd)
  ab, bb, cb = ops.bwd_pre(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'],
                           None, None, gLd)
  for got, k in ((ab, 0), (bb, 1), (cb, 2), (n_bar, 3)):
    w64, w32 = res[torch.float64][1][k], res[torch.float32][1][k]
    ref = _rel_l2_rows(w32, w64)
    assert _rel_l2_rows(got, w64) < max(GRAD_RTOL, 4 * ref), (k, _rel_l2_rows(got, w64), ref)


@pytest.mark.parametrize('B,seed', [(8, 0), (3, 1), (128, 2)])
def test_aux_topk_parity(cuda_device, B, seed):
  ops = _ops()
  rng = np.random.default_rng(seed)
  L, k = 50, 15
  logits = torch.from_numpy(rng.standard_normal((B, L)).astype(np.float32)) * 2.0
  G = torch.from_numpy(rng.gamma(1.0 / k, size=(10, B, L)).astype(np.float32))
  lg = logits.clone().requires_grad_(True)
  emb, kl = O.topk_embedding_and_loss(lg, G, k, L)
  eb = torch.from_numpy(rng.standard_normal((B, L)).astype(np.float32))
  kb = torch.from_numpy(rng.standard_normal((B,)).astype(np.float32))
  (want_lb,) = torch.autograd.grad((emb * eb).sum() + (kl * kb).sum(), [lg])
  d = lambda v: v.to(cuda_device).contiguous()
  got_emb, got_kl = ops.aux_topk_fwd(d(logits), d(G), k)
  assert torch.equal((got_emb.cpu() > 0.5), (emb.detach() > 0.5))
  assert (got_emb.cpu().sum(dim=1) - emb.detach().sum(dim=1)).abs().max() < 1e-5
  assert (got_emb.cpu() - emb.detach()).abs().max().item() < 1e-6
  assert (got_kl.cpu() - kl.detach()).abs().max().item() < 1e-5 * kl.detach().abs().max().item() + 1e-7
  got_lb = ops.aux_topk_bwd(d(logits), d(G), k, d(eb), d(kb))
  assert _rel_l2_rows(got_lb, want_lb) < GRAD_RTOL
  # KAT: uniform logits -> KL == 0 (no noise -> hard mask is all ones by ties)
  z = torch.zeros((4, L), device=cuda_device)
  e0, k0 = ops.aux_topk_fwd(z, None, k)
  assert k0.abs().max().item() < 1e-6


def test_bpd_reduce_and_loss_fn(cuda_device):
  ops = _ops()
  B = 16
  inp = O.synth_inputs(B, 7)
  rng = np.random.default_rng(5)
  kl_z = torch.from_numpy(rng.uniform(0, 2, B).astype(np.float32))
  out, aux = _oracle(inp, O.MODE_EPS)
  out = O.VDMOutput(out.loss_recon, kl_z + aux['loss_klz_prior'], out.loss_diff, out.var_0,
                    out.var_1)
  bpd, sc = O.loss_fn_bpd(out)
  g = _dev(inp, cuda_device)
  desc = ops.Desc()
  r = ops.fwd_pre(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'], g['eps'])
  diff = ops.fwd_post(desc, None, None, None, None, None, g['eps'], g['net'], r['w'])
  got, tot = ops.bpd_reduce(desc, r['loss_recon'], r['loss_klz_prior'], kl_z.to(cuda_device), diff,
                            r['var_sums'], want_klz_total=True)
  got = got.cpu()
  assert abs(got[0].item() - bpd.item()) < BPD_ATOL
  for i, key in enumerate(['bpd', 'bpd_latent', 'bpd_recon', 'bpd_diff']):
    assert abs(got[i].item() - sc[key].item()) < 1e-5 * abs(sc[key].item()) + 1e-7, key
  assert abs(got[4].item() - sc['var0'].item()) < 1e-9
  assert abs(got[5].item() - sc['var'].item()) < 1e-6
  assert _rel(tot, out.loss_klz) < LOSS_RTOL


@pytest.mark.parametrize('mode', list(MODES))
def test_autograd_pair_with_denoiser(cuda_device, mode):
  """End to end through mulan_pre -> torch denoiser -> mulan_post, gradients w.r.t. the
  coefficient heads and the denoiser weights, against oracle autograd."""
  ops = _ops()
  B = 8
  inp = O.synth_inputs(B, 51)
  w1 = torch.tensor(0.7)
  w2 = torch.tensor(0.05)

  def net_fn(z, g, w1, w2, noise):
    return w1 * z.reshape(z.shape[0], -1) + w2 * g.reshape(-1, 1) + noise

  for dtype, store in ((torch.float64, 'w64'), (torch.float32, 'w32')):
    cast = lambda v: v.to(dtype) if v.is_floating_point() else v
    i = {k: cast(v) for k, v in inp.items()}
    a, b, c = (i[k].clone().requires_grad_(True) for k in ('a', 'b', 'c'))
    p1 = w1.detach().clone().to(dtype).requires_grad_(True)
    p2 = w2.detach().clone().to(dtype).requires_grad_(True)
    out = O.elbo_terms(i['x'], a, b, c, i['t'], i['eps_0'], i['eps'],
                       lambda z, g: net_fn(z, g, p1, p2, 0.3 * i['net']), MODES[mode],
                       O.OracleConfig(), dtype=dtype)
    bpd, _ = O.loss_fn_bpd(out)
    grads = torch.autograd.grad(bpd, [a, b, c, p1, p2])
    if store == 'w64':
      want, want_bpd = grads, bpd.item()
  g = _dev(inp, cuda_device)
  a, b, c = (g[k].clone().requires_grad_(True) for k in ('a', 'b', 'c'))
  p1 = w1.detach().clone().to(cuda_device).requires_grad_(True)
  p2 = w2.detach().clone().to(cuda_device).requires_grad_(True)
  desc = ops.Desc(param=MODES[mode])
  tape = ops.ElboTape(desc)
  z_t, g_net, rec, klz, vs, link = ops.mulan_pre(tape, g['x'], a, b, c, g['t'], g['eps_0'],
                                                 g['eps'])
  net = net_fn(z_t, g_net, p1, p2, 0.3 * g['net'])
  diff = ops.mulan_post(tape, net, link)
  bpd = (rec.mean() + klz.mean() + diff.mean()) / (3072 * math.log(2.0))
  assert abs(bpd.item() - want_bpd) < BPD_ATOL
  bpd.backward()
  assert _rel_l2_rows(a.grad, want[0]) < GRAD_RTOL
  assert _rel_l2_rows(b.grad, want[1]) < GRAD_RTOL
  assert _rel_l2_rows(c.grad, want[2]) < GRAD_RTOL
  assert abs(p1.grad.item() - want[3].item()) < GRAD_RTOL * abs(want[3].item())
  assert abs(p2.grad.item() - want[4].item()) < GRAD_RTOL * abs(want[4].item()) + 1e-9


def test_full_size_properties(cuda_device):
  """At BASELINE's full per-GPU sizes (and beyond L2): size-independent properties.
  - determinism (bitwise identical reruns),
  - row independence (a row's outputs do not depend on its neighbours / batch size),
  - v-from-eps loss == eps loss (algebraic identity),
  - gamma monotone: w >= 0, and the prior KL / recon terms are >= 0.
  """
  ops = _ops()
  B = 2048
  inp = O.synth_inputs(B, 61, group=128)
  g = _dev(inp, cuda_device)
  desc = ops.Desc()
  r1 = ops.fwd_pre(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'], g['eps'])
  r2 = ops.fwd_pre(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'], g['eps'])
  for k in ('z_t', 'g_net', 'w', 'loss_recon', 'loss_klz_prior', 'var_sums'):
    assert torch.equal(r1[k], r2[k]), k
  sl = slice(700, 716)
  sub = {k: v[sl].contiguous() for k, v in g.items()}
  rs = ops.fwd_pre(desc, sub['x'], sub['a'], sub['b'], sub['c'], sub['t'], sub['eps_0'],
                   sub['eps'])
  # per-pixel outputs are independent of the launch; per-row sums depend on the launch's SHAPE
  # (threads per row is a function of the row count: 768 up to one row per SM, 256 up to four,
  # 128 beyond) only through the float32 summation order
  for k in ('z_t', 'w'):
    assert torch.equal(r1[k][sl], rs[k]), k
  for k in ('g_net', 'loss_recon', 'loss_klz_prior'):
    assert _rel(rs[k], r1[k][sl]) < 2e-6, k
  # same shape (two launches of more than 4 rows per SM): bitwise row independence
  big = slice(300, 1100)
  sub2 = {k: v[big].contiguous() for k, v in g.items()}
  rb = ops.fwd_pre(desc, sub2['x'], sub2['a'], sub2['b'], sub2['c'], sub2['t'], sub2['eps_0'],
                   sub2['eps'])
  for k in ('z_t', 'g_net', 'w', 'loss_recon', 'loss_klz_prior'):
    assert torch.equal(r1[k][big], rb[k]), k
  assert (r1['w'] >= 0).all()
  assert (r1['loss_recon'] >= 0).all() and (r1['loss_klz_prior'] >= 0).all()
  d_eps = ops.fwd_post(ops.Desc(param=0), g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'],
                       g['net'], None)
  d_vfe = ops.fwd_post(ops.Desc(param=2), g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'],
                       g['net'], None)
  assert _rel(d_vfe, d_eps) < 2e-5
  # spot-check 16 rows of the big batch against the oracle
  osub = {k: v[sl] for k, v in inp.items()}
  out, aux = _oracle(osub, O.MODE_EPS)
  assert _rel(r1['loss_recon'][sl], out.loss_recon) < LOSS_RTOL
  assert _rel(d_eps[sl], out.loss_diff) < LOSS_RTOL


@pytest.mark.parametrize('mode', ['eps', 'vel_from_eps'])
def test_host_entry_matches_device_api(cuda_device, mode):
  """mulan_elbo_host (host buffers, chunked 3-stream pipeline) == the device-pointer API, row
  for row, including a batch that spans several chunks with a ragged tail."""
  ops = _ops()
  from mulan_b200 import host
  B = 2500                        # 1024-row chunks: 1024 + 1024 + 452
  inp = O.synth_inputs(B, 81, group=128)
  g = _dev(inp, cuda_device)
  desc = ops.Desc(param=MODES[mode])
  ws = ops.ElboWorkspace(desc, B, cuda_device)
  gL = torch.full((B,), 1.0 / (B * 3072 * math.log(2.0)), device=cuda_device)
  args = (g['x'], g['a'], g['b'], g['c'], g['t'])
  ws.fwd_pre(*args, g['eps_0'], g['eps'])
  ws.fwd_post(*args, g['eps'], g['net'])
  ws.bpd_reduce(None)
  ws.bwd_post(*args, g['eps'], g['net'], gL)
  ws.bwd_pre(*args, g['eps'], g['net'], None, None, gL)
  torch.cuda.synchronize()
  npy = {k: v.numpy() for k, v in inp.items()}
  r = host.elbo_host(npy['x'], npy['a'], npy['b'], npy['c'], npy['t'], npy['eps_0'], npy['eps'],
                     npy['net'], param=MODES[mode], want_grad=True)
  same = lambda a, b: np.array_equal(np.asarray(a), b.cpu().numpy())
  assert same(r['loss_recon'], ws.loss_recon) and same(r['loss_klz_prior'], ws.loss_klz_prior)
  assert same(r['loss_diff'], ws.loss_diff)
  assert same(r['a_bar'], ws.a_bar) and same(r['b_bar'], ws.b_bar) and same(r['c_bar'], ws.c_bar)
  assert same(r['n_bar'], ws.n_bar)
  assert np.allclose(r['scalars'], ws.scalars.cpu().numpy(), rtol=1e-6)
  # and against the oracle on a few rows
  sl = slice(1020, 1030)
  out, _ = _oracle({k: v[sl] for k, v in inp.items()}, MODES[mode])
  assert _rel(torch.from_numpy(r['loss_diff'][sl].copy()), out.loss_diff) < LOSS_RTOL
  assert _rel(torch.from_numpy(r['loss_recon'][sl].copy()), out.loss_recon) < LOSS_RTOL


def test_host_entry_denoiser_callback(cuda_device):
  """The denoiser callback is invoked once per chunk, in row order, with that chunk's device
  pointers and the compute stream."""
  from mulan_b200 import host
  import ctypes as C
  B = 2100
  inp = O.synth_inputs(B, 82, group=128)
  npy = {k: v.numpy() for k, v in inp.items()}
  net_dev = inp['net'].to(cuda_device).contiguous()
  torch.cuda.synchronize()
  cudart = C.CDLL('libcudart.so.12')      # the runtime torch already loaded
  cudart.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
  calls = []

  def denoiser(rows, z_ptr, g_ptr, n_ptr, stream):
    # stand-in network: copy this chunk's slice of a resident tensor into the chunk's output
    r0 = sum(calls)
    calls.append(rows)
    src = net_dev.data_ptr() + r0 * 3072 * 4
    return cudart.cudaMemcpyAsync(n_ptr, src, rows * 3072 * 4, 3, stream)   # 3 = D2D

  r = host.elbo_host(npy['x'], npy['a'], npy['b'], npy['c'], npy['t'], npy['eps_0'], npy['eps'],
                     None, want_grad=False, denoiser=denoiser)
  assert calls == [1024, 1024, 52]
  ref = host.elbo_host(npy['x'], npy['a'], npy['b'], npy['c'], npy['t'], npy['eps_0'], npy['eps'],
                       npy['net'], want_grad=False)
  assert np.array_equal(r['loss_diff'], ref['loss_diff'])


# ------------------------------------------------------------------------------------------
# Paths the shipped configuration never takes: other gamma ranges (wider reconstruction window,
# prior-KL constants that are NOT uniform over the three roundings of gamma_1), a vocabulary that
# is not a power of two, and sub-pixels whose scale S is outside the fast reciprocal's range.
# ------------------------------------------------------------------------------------------
def _generic_case(cuda_device, cfg, inp, modes=('eps', 'vel', 'vel_from_eps'), gtol=GRAD_RTOL):
  ops = _ops()
  B = inp['a'].shape[0]
  g = _dev(inp, cuda_device)
  rng = np.random.default_rng(7)
  gL = torch.from_numpy(rng.uniform(0.5, 1.5, B).astype(np.float32)) / (B * 3072 * math.log(2))
  zbar = torch.from_numpy(rng.standard_normal((B, 3072)).astype(np.float32)) * 1e-4
  gbar = torch.from_numpy(rng.standard_normal(B).astype(np.float32)) * 1e-3
  for mode in modes:
    out, aux = _oracle(inp, MODES[mode], cfg=cfg)
    desc = ops.Desc(param=MODES[mode], vocab=cfg.vocab_size, gamma_min=cfg.gamma_min,
                    gamma_max=cfg.gamma_max)
    r = ops.fwd_pre(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'], g['eps'])
    assert _rel(r['loss_recon'], out.loss_recon) < LOSS_RTOL, mode
    assert _rel(r['loss_klz_prior'], aux['loss_klz_prior']) < LOSS_RTOL, mode
    assert _rel_l2_rows(r['z_t'], aux['z_t'].reshape(B, -1)) < 1e-5
    var0 = r['var_sums'][:, 0].sum().item() / (B * 3072)
    var1 = r['var_sums'][:, 1].sum().item() / (B * 3072)
    assert abs(var0 - out.var_0.item()) < 2e-6 * out.var_0.item() + 1e-12
    assert abs(var1 - out.var_1.item()) < 2e-6
    got = ops.fwd_post(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'], None)
    assert _rel(got, out.loss_diff) < LOSS_RTOL, mode
    # backward against float64 autograd of the oracle under the same config
    cast = lambda v: v.double() if v.is_floating_point() else v
    i = {k: cast(v) for k, v in inp.items()}
    a, b, c, net = (i[k].clone().requires_grad_(True) for k in ('a', 'b', 'c', 'net'))
    o64, aux64 = O.elbo_terms(i['x'], a, b, c, i['t'], i['eps_0'], i['eps'], lambda z, gg: net,
                              MODES[mode], cfg, dtype=torch.float64, return_aux=True)
    L = (gL.double() * o64.loss_diff).sum() + (zbar.double() * aux64['z_t'].reshape(B, -1)).sum()
    L = L + (gbar.double() * O.score_model_gt(aux64['g_t'], cfg).reshape(B)).sum()
    want = torch.autograd.grad(L, [a, b, c, net])
    n_bar = ops.bwd_post(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'], None,
                         gL.to(cuda_device))
    ab, bb, cb = ops.bwd_pre(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'],
                             zbar.to(cuda_device), gbar.to(cuda_device), gL.to(cuda_device))
    for got_, w_, name in ((ab, want[0], 'a'), (bb, want[1], 'b'), (cb, want[2], 'c'),
                           (n_bar, want[3], 'n')):
      assert _rel_l2_rows(got_, w_.reshape(B, -1)) < gtol, (mode, name)


@pytest.mark.parametrize('gmin,gmax', [(-6.0, 3.0), (-2.0, 0.3), (-9.5, 7.25)])
def test_other_gamma_ranges(cuda_device, gmin, gmax):
  """gamma_0 = -6: bins are 0.16 decoder-sigmas apart -> the windowed generic log-softmax;
  (-2, 0.3): sigmoid(gamma_1) depends on how (Delta S)/S rounds -> per-pixel prior KL / var_1."""
  cfg = O.OracleConfig(gamma_min=gmin, gamma_max=gmax)
  _generic_case(cuda_device, cfg, O.synth_inputs(6, 51))


def test_vocab_not_a_power_of_two(cuda_device):
  # gamma_0 = -8 puts the 100 bins 1.09 decoder-sigmas apart: a non-trivial reconstruction term
  cfg = O.OracleConfig(vocab_size=100, gamma_min=-8.0)
  inp = O.synth_inputs(5, 52)
  inp['x'] = (inp['x'].to(torch.int32) % 100).to(torch.uint8)
  _generic_case(cuda_device, cfg, inp)


def test_scale_outside_fast_range(cuda_device):
  """S = c^2 = 1e-32 (a = b = 0, c = 1e-16) is below the fast reciprocal's range: those
  sub-pixels take the IEEE per-pixel path and must still match the reference; S == 0 (all
  coefficients zero) is NaN in the reference and must be NaN here, in that row only."""
  ops = _ops()
  B = 4
  inp = O.synth_inputs(B, 53)
  for k in ('a', 'b'):
    inp[k][1] = 0.0
    inp[k][2, ::3] = 0.0
  inp['c'][1] = 1e-16
  inp['c'][2, ::3] = 1e-16
  out, aux = _oracle(inp, O.MODE_EPS)
  g = _dev(inp, cuda_device)
  desc = ops.Desc()
  r = ops.fwd_pre(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'], g['eps'])
  assert torch.isfinite(out.loss_recon).all()
  assert _rel(r['loss_recon'], out.loss_recon) < LOSS_RTOL
  assert _rel(r['loss_klz_prior'], aux['loss_klz_prior']) < LOSS_RTOL
  assert _rel_l2_rows(r['z_t'], aux['z_t'].reshape(B, -1)) < 1e-5
  assert _rel_l2_rows(r['w'], aux['g_t_grad'].reshape(B, -1)) < 1e-5
  var1 = r['var_sums'][:, 1].sum().item() / (B * 3072)
  assert abs(var1 - out.var_1.item()) < 2e-6
  got = ops.fwd_post(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'], None)
  assert _rel(got, out.loss_diff) < LOSS_RTOL
  # S == 0
  for k in ('a', 'b', 'c'):
    inp[k][3, 5] = 0.0
  out, aux = _oracle(inp, O.MODE_EPS)
  assert torch.isnan(out.loss_klz[3]) and torch.isfinite(out.loss_klz[:3]).all()
  g = _dev(inp, cuda_device)
  r = ops.fwd_pre(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'], g['eps'])
  assert torch.isnan(r['loss_klz_prior'][3]) and torch.isnan(r['var_sums'][3]).all()
  assert torch.isfinite(r['loss_klz_prior'][:3]).all() and torch.isfinite(r['var_sums'][:3]).all()
  assert _rel(r['loss_klz_prior'][:3], aux['loss_klz_prior'][:3]) < LOSS_RTOL


def test_host_entry_keyed_draws(cuda_device):
  """mulan_elbo_host_keyed draws eps_0 / eps on the device from raw threefry keys; it must equal
  mulan_elbo_host fed with the same draws materialised by mulan_rng_normal (bit for bit), over a
  batch that spans several chunks."""
  ops = _ops()
  from mulan_b200 import host
  B = 2100
  inp = O.synth_inputs(B, 83, group=128)
  k0, k1 = (11, 22), (0xdeadbeef, 7)
  e0 = ops.rng_normal(k0, (B, 3072), device=cuda_device).cpu().numpy()
  e1 = ops.rng_normal(k1, (B, 3072), device=cuda_device).cpu().numpy()
  npy = {k: v.numpy() for k, v in inp.items()}
  for param in (0, 1):
    want = host.elbo_host(npy['x'], npy['a'], npy['b'], npy['c'], npy['t'], e0, e1, npy['net'],
                          param=param, want_grad=True)
    want = {k: np.array(v, copy=True) for k, v in want.items()}
    got = host.elbo_host(npy['x'], npy['a'], npy['b'], npy['c'], npy['t'], None, None, npy['net'],
                         param=param, want_grad=True, jax_keys=(k0, k1))
    for k in want:
      assert np.array_equal(got[k], want[k]), (param, k)
  with pytest.raises(ValueError):
    host.elbo_host(npy['x'], npy['a'], npy['b'], npy['c'], npy['t'], e0, None, npy['net'],
                   jax_keys=(k0, k1))


# ----------------------------------------------------------------------------------------
# mulan_fwd_pre_consts: the five fixed-end constants as ANOTHER platform rounds them
# ----------------------------------------------------------------------------------------
_CONST_FIELDS = ('exp_half_g0', 'exp_neg_half_g0', 'sigmoid_g0', 'sigmoid_g1', 'log_sigmoid_g1')
# which outputs a constant may move (everything else must stay bit-identical)
_CONST_MOVES = {'exp_half_g0': {'loss_recon'}, 'exp_neg_half_g0': {'loss_recon'},
                'sigmoid_g0': {'var_sums'}, 'sigmoid_g1': {'loss_klz_prior', 'var_sums'},
                'log_sigmoid_g1': {'loss_klz_prior'}}


def _own_consts(desc, B):
  import ctypes as C
  from mulan_b200 import _lib
  k = _lib.MulanEndConsts()
  d = desc.c(B)
  assert _lib.load().mulan_host_end_consts(C.byref(d), C.byref(k)) == 0
  return k


def test_caller_supplied_end_constants_are_the_librarys_own(cuda_device):
  """Handing the library's own constants back selects the same kernel and the same bits."""
  from mulan_b200 import ops
  B = 6
  inp = O.synth_inputs(B, 71)
  g = {k: v.to(cuda_device).contiguous() for k, v in inp.items()}
  desc = ops.Desc()
  args = (g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'], g['eps'])
  base = ops.fwd_pre(desc, *args)
  same = ops.fwd_pre(desc, *args, end_consts=_own_consts(desc, B))
  for key in base:
    assert torch.equal(base[key], same[key]), key


@pytest.mark.parametrize('field', _CONST_FIELDS)
@pytest.mark.parametrize('ulps', [-1, 1])
def test_caller_supplied_end_constants_track_the_oracle(cuda_device, field, ulps):
  """One ulp of each constant, as a framework with a different exp / log would supply it
  (ldm/model_mulan_epsilon.py:311-325, model_vdm.py:286): the CUDA result follows the oracle
  evaluated with THE SAME constant to 1e-5, and only the outputs that constant feeds move."""
  from mulan_b200 import _lib, ops
  B = 6
  inp = O.synth_inputs(B, 71)
  g = {k: v.to(cuda_device).contiguous() for k, v in inp.items()}
  desc = ops.Desc()
  args = (g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'], g['eps'])
  base = ops.fwd_pre(desc, *args)
  own = _own_consts(desc, B)
  k2 = _lib.MulanEndConsts.from_buffer_copy(own)
  v = np.float32(getattr(own, field))
  moved_v = np.nextafter(v, np.float32(np.inf if ulps > 0 else -np.inf))
  if field == 'sigmoid_g1' and moved_v > 1:
    pytest.skip('sigmoid(g_1) + 1 ulp exceeds 1')
  setattr(k2, field, float(moved_v))
  got = ops.fwd_pre(desc, *args, end_consts=k2)
  for key in base:
    if key in _CONST_MOVES[field] or base[key] is None:
      continue
    assert torch.equal(base[key], got[key]), (field, key)
  out, aux = O.elbo_terms(inp['x'], inp['a'], inp['b'], inp['c'], inp['t'], inp['eps_0'],
                          inp['eps'], lambda z, gg: inp['net'], O.MODE_EPS, O.OracleConfig(),
                          return_aux=True,
                          end_consts={f: float(np.float32(getattr(k2, f))) for f in _CONST_FIELDS})
  rel = lambda a_, b_: float(((a_.double().cpu() - b_.double()).abs() / b_.double().abs()).max())
  assert rel(got['loss_recon'], out.loss_recon) < 1e-5
  assert rel(got['loss_klz_prior'], aux['loss_klz_prior']) < 1e-5
  D = inp['a'].shape[1]
  assert abs(got['var_sums'][:, 0].sum().item() / (B * D) - out.var_0.item()) < 1e-6 * out.var_0.item()
  assert abs(got['var_sums'][:, 1].sum().item() / (B * D) - out.var_1.item()) < 1e-6
  # and the constant really is live where one ulp of it is above the float32 resolution of the
  # row sum it feeds (log sigmoid(g_1) ~ -7e-3 is absorbed by v_1 - log v_1 ~ 1; one ulp of
  # exp(-g_0/2) or sigmoid(g_0) can round away in the products)
  if field in ('exp_half_g0', 'sigmoid_g1'):
    assert any(not torch.equal(base[key], got[key]) for key in _CONST_MOVES[field]), field
