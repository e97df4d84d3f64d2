"""GPU tests of the ABI v2 additions: the dense_out_c pre-activation taken by the kernels
(MULAN_FLAG_C_RAW), broadcast noise rows, programmatic dependent launch, the fused / parallel
loss-scalar reduction (mulan_post_bpd, mulan_bpd_reduce with a workspace), the alternative
fwd_pre kernel shapes, and concurrent host threads (one per stream / device, as jax.pmap's
executor threads call the custom-call targets, ldm/experiment.py:89-91)."""
import ctypes as C
import math
import os
import threading

import numpy as np
import pytest
import torch

from oracle import mulan_oracle as O

pytestmark = pytest.mark.gpu

PARAMS = [O.MODE_EPS, O.MODE_VEL, O.MODE_VEL_FROM_EPS]


def dev_inputs(B, seed, dev, raw=False):
  inp = O.synth_inputs(B, seed)
  g = {k: v.to(dev).contiguous() for k, v in inp.items()}
  if raw:
    r = np.random.default_rng(seed + 99).standard_normal((B, 3072)) * 3.0
    r[0, :64] = np.linspace(-30, 30, 64)            # softplus tails on both sides
    inp['c_raw'] = torch.from_numpy(r.astype(np.float32))
    inp['c'] = O.coefficients_from_raw(inp['c_raw'])
    g['c_raw'], g['c'] = inp['c_raw'].to(dev), inp['c'].to(dev)
  return inp, g


def step(desc, g, c_key='c', gL=None, z_bar=None, g_bar=None, save_w=None):
  """fwd_pre -> fwd_bwd_post -> bwd_pre through the raw ops; returns every output."""
  from mulan_b200 import ops
  B = g['a'].shape[0]
  if gL is None:
    gL = torch.full((B,), 1.0 / (B * 3072 * math.log(2.0)), device=g['a'].device)
  save_w = ops.saves_w(desc) if save_w is None else save_w
  pre = ops.fwd_pre(desc, g['x'], g['a'], g['b'], g[c_key], g['t'], g['eps_0'], g['eps'],
                    save_w=save_w)
  diff, n_bar = ops.fwd_bwd_post(desc, g['x'], g['a'], g['b'], g[c_key], g['t'], g['eps'],
                                 g['net'], pre['w'], gL)
  ab, bb, cb = ops.bwd_pre(desc, g['x'], g['a'], g['b'], g[c_key], g['t'], g['eps'], g['net'],
                           z_bar, g_bar, gL)
  out = dict(pre)
  out.update(loss_diff=diff, n_bar=n_bar, a_bar=ab, b_bar=bb, c_bar=cb)
  return out


def rel(a, b):
  a, b = a.double(), b.double()
  return float(((a - b).abs() / b.abs().clamp_min(1e-30)).max())


def rel_l2(a, b):
  a, b = a.double(), b.double()
  return float((a - b).norm() / b.norm().clamp_min(1e-300))


@pytest.mark.parametrize('param', PARAMS)
def test_c_raw_matches_activated_c(cuda_device, param):
  """MULAN_FLAG_C_RAW: handing over the pre-activation r of dense_out_c gives the losses of
  c = 1e-3 + softplus(r) (ldm/model_mulan_epsilon.py:537) and returns c_bar * sigmoid(r)."""
  from mulan_b200 import ops
  B = 6
  inp, g = dev_inputs(B, 11, cuda_device, raw=True)
  z_bar = (1e-4 * torch.randn(B, 3072, generator=torch.Generator().manual_seed(1))).to(cuda_device)
  g_bar = (1e-3 * torch.randn(B, generator=torch.Generator().manual_seed(2))).to(cuda_device)
  base = step(ops.Desc(param=param), g, 'c', z_bar=z_bar, g_bar=g_bar)
  raw = step(ops.Desc(param=param, c_raw=True), g, 'c_raw', z_bar=z_bar, g_bar=g_bar)
  for k in ('loss_recon', 'loss_klz_prior', 'loss_diff'):
    assert rel(raw[k], base[k]) < 2e-6, k
  assert rel_l2(raw['z_t'], base['z_t']) < 1e-6
  assert rel_l2(raw['n_bar'], base['n_bar']) < 1e-5
  assert rel_l2(raw['a_bar'], base['a_bar']) < 1e-5
  assert rel_l2(raw['b_bar'], base['b_bar']) < 1e-5
  want = base['c_bar'] * torch.sigmoid(g['c_raw'])
  assert rel_l2(raw['c_bar'], want) < 1e-5
  # and the in-kernel softplus against the oracle's (float64) element by element
  c64 = O.coefficients_from_raw(inp['c_raw'].double())
  pix = ops.Desc(param=param, gt_mode=1, c_raw=True)
  g_pix = ops.fwd_pre(pix, g['x'], g['a'], g['b'], g['c_raw'], g['t'], g['eps_0'], g['eps'],
                      save_w=False)['g_net']
  cfg = O.OracleConfig()
  want_g = O.eval_polynomial(inp['a'].double(), inp['b'].double(), c64,
                             inp['t'].double().reshape(-1, 1), cfg)
  assert (g_pix.double().cpu() - want_g).abs().max() < 5e-4     # cancellation-limited pixels
  assert (g_pix.double().cpu() - want_g).abs().mean() < 5e-6


def test_c_raw_through_the_model(cuda_device):
  """model.VDM hands the built-in schedule head's pre-activation to the kernels; losses and
  parameter gradients equal the unfused path (softplus and its backward in torch)."""
  import sys
  sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
  import golden_inputs as GI
  from mulan_b200.model import VDM, VDMConfig, loss_fn
  dev = cuda_device
  B = 4
  torch.manual_seed(0)
  enc_w = (0.1 * torch.randn(256, 50)).to(dev)
  w = torch.nn.Parameter(torch.tensor(0.7, device=dev))
  model = VDM(VDMConfig(vdm_type='mulan_velocity'),
              lambda f, det: f.reshape(f.shape[0], -1)[:, :256] @ enc_w,
              lambda z, gt, cond, det: w * z + 0.01 * gt.reshape(-1, 1, 1, 1)).to(dev)
  model.gamma.load_flax(GI.mlp_weights(5))
  images = torch.randint(0, 256, (B, 32, 32, 3), device=dev,
                         generator=torch.Generator(device=dev).manual_seed(3))
  draws = model.make_draws(B, dev, torch.Generator(device=dev).manual_seed(4))
  res = {}
  for fused in (True, False):
    model.fused_softplus = fused
    model.zero_grad()
    w.grad = None
    bpd, _ = loss_fn(model, {'images': images}, draws=draws)
    bpd.backward()
    res[fused] = (bpd.item(), model.gamma.dense_out_c.bias.grad.clone(),
                  model.gamma.dense_out_b.bias.grad.clone(), w.grad.clone())
  assert abs(res[True][0] - res[False][0]) < 1e-5 * abs(res[False][0])
  for i in (1, 2, 3):
    assert rel_l2(res[True][i], res[False][i]) < 2e-5, i


@pytest.mark.parametrize('param', PARAMS)
def test_noise_rows_broadcast_is_bitwise_the_tiled_call(cuda_device, param):
  """noise_rows = N: eps_0 / eps are [N, D] and row b reads row b % N -- the same bits as
  materialising the tiled [B, D] arrays (dense-VLB evaluation, notebook_utils.py:178-185)."""
  from mulan_b200 import ops
  B, N = 12, 4
  _, g = dev_inputs(B, 21, cuda_device)
  gt = dict(g)
  gt['eps_0'] = g['eps_0'][:N].repeat(B // N, 1).contiguous()
  gt['eps'] = g['eps'][:N].repeat(B // N, 1).contiguous()
  gt['net'] = g['net']
  tiled = step(ops.Desc(param=param), gt)
  gb = dict(g)
  gb['eps_0'], gb['eps'] = g['eps_0'][:N].contiguous(), g['eps'][:N].contiguous()
  bc = step(ops.Desc(param=param, noise_rows=N), gb)
  for k, v in tiled.items():
    if v is not None:
      assert torch.equal(v, bc[k]), k


@pytest.mark.parametrize('param', PARAMS)
def test_pdl_launch_is_bitwise_the_plain_launch(cuda_device, param):
  from mulan_b200 import ops
  _, g = dev_inputs(9, 31, cuda_device)
  plain = step(ops.Desc(param=param), g)
  for _ in range(3):
    pdl = step(ops.Desc(param=param, pdl=True), g)
    for k, v in plain.items():
      if v is not None:
        assert torch.equal(v, pdl[k]), k


@pytest.mark.parametrize('rows', [1, 5, 128, 129, 1000])
@pytest.mark.parametrize('param', [O.MODE_EPS, O.MODE_VEL])
def test_post_bpd_equals_separate_post_and_reduce(cuda_device, rows, param):
  """mulan_post_bpd == mulan_fwd_bwd_post (or mulan_fwd_post) + mulan_bpd_reduce, bit for bit,
  in all three forms of the reduction (fused, parallel with workspace, single CTA); the
  workspace is left zero and is reusable; scalars match the oracle's loss_fn."""
  from mulan_b200 import ops
  dev = cuda_device
  inp, g = dev_inputs(rows, 41 + rows, dev)
  desc = ops.Desc(param=param)
  gL = torch.full((rows,), 1.0 / (rows * 3072 * math.log(2.0)), device=dev)
  kl_z = torch.rand(rows, generator=torch.Generator().manual_seed(5)).to(dev)
  pre = ops.fwd_pre(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'], g['eps'],
                    save_w=ops.saves_w(desc))
  diff, n_bar = ops.fwd_bwd_post(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'],
                                 g['net'], pre['w'], gL)
  ws = ops.reduce_workspace(rows, dev)
  sc_par, tot_par = ops.bpd_reduce(desc, pre['loss_recon'], pre['loss_klz_prior'], kl_z, diff,
                                   pre['var_sums'], want_klz_total=True, ws=ws)
  sc_one, tot_one = ops.bpd_reduce(desc, pre['loss_recon'], pre['loss_klz_prior'], kl_z, diff,
                                   pre['var_sums'], want_klz_total=True, ws=None)
  assert torch.equal(sc_par, sc_one) and torch.equal(tot_par, tot_one)
  n_counters = ((rows + 127) // 128 + 1 + 3) // 4 * 4      # the rest holds the group partials
  assert torch.count_nonzero(ws[:n_counters]).item() == 0
  for with_grad in (True, False, True):          # reuse the same workspace
    f = ops.post_bpd(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'], pre['w'],
                     gL if with_grad else None, pre['loss_recon'], pre['loss_klz_prior'], kl_z,
                     pre['var_sums'], want_klz_total=True, ws=ws)
    assert torch.equal(f['loss_diff'], diff)
    assert torch.equal(f['scalars'], sc_par)
    assert torch.equal(f['loss_klz_total'], tot_par)
    if with_grad:
      assert torch.equal(f['n_bar'], n_bar)
    assert torch.count_nonzero(ws[:n_counters]).item() == 0
  # against the oracle's loss_fn (ldm/experiment_vdm.py:62-74)
  out = O.elbo_terms(inp['x'], inp['a'], inp['b'], inp['c'], inp['t'], inp['eps_0'], inp['eps'],
                     lambda z, gg: inp['net'], param, O.OracleConfig(), kl_z=kl_z.cpu())
  bpd, sc = O.loss_fn_bpd(out)
  got = sc_par.cpu()
  for i, k in enumerate(('bpd', 'bpd_latent', 'bpd_recon', 'bpd_diff')):
    assert abs(got[i].item() - sc[k].item()) < 1e-5 * abs(sc[k].item()), k
  assert abs(got[4].item() - sc['var0'].item()) < 1e-6 * sc['var0'].item()
  assert abs(got[5].item() - sc['var'].item()) < 1e-6


def test_fwd_pre_kernel_shapes_agree(cuda_device, monkeypatch):
  """MULAN_FWD_PRE_V: CTA size / residency change how the work is scheduled, never the
  arithmetic -- per-pixel outputs are bit-identical, per-row sums agree to float32 summation
  order."""
  from mulan_b200 import ops
  _, g = dev_inputs(37, 51, cuda_device)
  desc = ops.Desc()
  args = (g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'], g['eps'])
  monkeypatch.delenv('MULAN_FWD_PRE_V', raising=False)
  base = ops.fwd_pre(desc, *args)
  for v in (0, 1, 5, 12):
    monkeypatch.setenv('MULAN_FWD_PRE_V', str(v))
    for save_w in (True, False):
      got = ops.fwd_pre(desc, *args, save_w=save_w)
      assert torch.equal(got['z_t'], base['z_t']), v
      if save_w:
        assert torch.equal(got['w'], base['w']), v
      for k in ('loss_recon', 'loss_klz_prior', 'g_net', 'var_sums'):
        assert rel(got[k], base[k]) < 2e-6, (v, k)
  monkeypatch.delenv('MULAN_FWD_PRE_V', raising=False)


def test_concurrent_host_threads_bitwise_serial(cuda_device):
  """The single-process contract of jax.pmap (ldm/experiment.py:89-91): XLA calls the custom-call
  targets from one executor thread per device.  8 host threads, each with its own stream (and
  its own device when several are visible), drive the legacy XLA targets concurrently; every
  result equals the serial one bit for bit."""
  from mulan_b200 import _lib, ops
  lib = _lib.load()
  n_dev = torch.cuda.device_count()
  n_threads, iters, B = 8, 6, 96
  jobs = []
  for i in range(n_threads):
    dev = torch.device(f'cuda:{i % n_dev}')
    _, g = dev_inputs(B, 100 + i, dev)
    jobs.append((dev, g))

  def run(dev, g, stream):
    """fwd_pre -> fwd_bwd_post -> bwd_pre -> bpd_reduce through the mulan_xla_* targets."""
    with torch.cuda.device(dev):
      f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
      z, gn, w, rec, klz, vs = f(B, 3072), f(B), f(B, 3072), f(B), f(B), f(B, 2)
      diff, nb, ab, bb, cb, sc, tot = f(B), f(B, 3072), f(B, 3072), f(B, 3072), f(B, 3072), f(6), f(B)
      gL = torch.full((B,), 1.0 / (B * 3072 * math.log(2.0)), device=dev)
      op = _lib.MulanXlaOpaque(desc=_lib.make_desc(rows=B), absent_mask=0, reserved=0)
      opb = _lib.MulanXlaOpaque(desc=_lib.make_desc(rows=B), absent_mask=(1 << 7) | (1 << 8),
                                reserved=0)
      opr = _lib.MulanXlaOpaque(desc=_lib.make_desc(rows=B), absent_mask=1 << 2, reserved=0)

      def call(name, bufs, o):
        arr = (C.c_void_p * len(bufs))(*[b.data_ptr() for b in bufs])
        getattr(lib, name)(C.c_void_p(stream.cuda_stream), arr, bytes(o), C.sizeof(o), None)
      call('mulan_xla_fwd_pre', [g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'], g['eps'],
                                 z, gn, w, rec, klz, vs], op)
      call('mulan_xla_fwd_bwd_post', [g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'],
                                      w, gL, diff, nb], op)
      call('mulan_xla_bwd_pre', [g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'],
                                 g['eps'], g['t'], gL, ab, bb, cb], opb)
      call('mulan_xla_bpd_reduce', [rec, klz, rec, diff, vs, sc, tot], opr)
      stream.synchronize()
      return [t.clone() for t in (z, gn, w, rec, klz, vs, diff, nb, ab, bb, cb, sc)]

  streams = [torch.cuda.Stream(device=dev) for dev, _ in jobs]
  serial = [run(dev, g, s) for (dev, g), s in zip(jobs, streams)]
  results, errors = [None] * n_threads, []
  start = threading.Barrier(n_threads)

  def worker(i):
    try:
      dev, g = jobs[i]
      torch.cuda.set_device(dev)
      start.wait()
      for _ in range(iters):
        results[i] = run(dev, g, streams[i])
    except Exception as exc:      # surfaced below
      errors.append((i, repr(exc)))
  threads = [threading.Thread(target=worker, args=(i,)) for i in range(n_threads)]
  for t in threads:
    t.start()
  for t in threads:
    t.join()
  assert not errors, errors
  for i in range(n_threads):
    for a_, b_ in zip(results[i], serial[i]):
      assert torch.equal(a_, b_), i
  assert lib.mulan_last_error() is not None


@pytest.mark.parametrize('rows', [2, 10, 300])
@pytest.mark.parametrize('gt,c_raw', [(0, False), (1, False), (0, True)])
def test_fwd_pre_keyed_equals_device_draws_plus_fwd_pre(cuda_device, rows, gt, c_raw):
  """mulan_fwd_pre_keyed draws eps_0 / eps INSIDE the kernel from their threefry keys
  (jax.random.normal(make_rng('sample'), f.shape), ldm/model_mulan_epsilon.py:315, :327): the
  draws are bit for bit mulan_rng_normal's, the per-pixel outputs bit for bit mulan_fwd_pre's on
  those arrays, the per-row sums equal to float32 summation order (different CTA shape)."""
  from mulan_b200 import ops
  inp, g = dev_inputs(rows, 77 + rows, cuda_device, raw=True)
  k0, k1 = (123, 456), (789, 1011)
  D = 3072
  eps0 = ops.rng_normal(k0, (rows, D), device=cuda_device)
  eps = ops.rng_normal(k1, (rows, D), device=cuda_device)
  desc = ops.Desc(gt_mode=gt, c_raw=c_raw)
  c = g['c_raw'] if c_raw else g['c']
  base = ops.fwd_pre(desc, g['x'], g['a'], g['b'], c, g['t'], eps0, eps)
  got = ops.fwd_pre_keyed(desc, k0, k1, g['x'], g['a'], g['b'], c, g['t'], want_eps0=True)
  assert torch.equal(got['eps'], eps) and torch.equal(got['eps_0'], eps0)
  assert torch.equal(got['z_t'], base['z_t']) and torch.equal(got['w'], base['w'])
  if gt == 1:
    assert torch.equal(got['g_net'], base['g_net'])
  else:
    assert rel(got['g_net'], base['g_net']) < 2e-6
  for k in ('loss_recon', 'loss_klz_prior', 'var_sums'):
    assert rel(got[k], base[k]) < 2e-6, k
  # without the optional copies
  lean = ops.fwd_pre_keyed(desc, k0, k1, g['x'], g['a'], g['b'], c, g['t'], save_w=False,
                           want_eps=False)
  assert lean['eps'] is None and lean['w'] is None
  assert torch.equal(lean['z_t'], base['z_t'])
  assert torch.equal(lean['loss_recon'], got['loss_recon'])


def test_fwd_pre_keyed_rejects_what_it_cannot_pair(cuda_device):
  from mulan_b200 import _lib, ops
  _, g = dev_inputs(3, 5, cuda_device)
  with pytest.raises(_lib.MulanError, match='even'):
    ops.fwd_pre_keyed(ops.Desc(), (1, 2), (3, 4), g['x'], g['a'], g['b'], g['c'], g['t'])


def _rows_l2(got, want):
  """per-row relative L2 distance, float64 on the host"""
  got, want = got.double().cpu(), want.double().cpu()
  return ((got - want).flatten(1).norm(dim=1) / want.flatten(1).norm(dim=1).clamp_min(1e-30))


def _oracle_terms_and_grads(inp, param, gL, zbar, gbar, dtype):
  """Forward terms, z_t and the cotangents of (a, b, c, net) for
  L = sum gL * loss_diff + <zbar, z_t> + <gbar, mean gamma_t>, evaluated in `dtype`."""
  cfg = O.OracleConfig()
  d = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in inp.items()}
  a, b, c, net = (d[k].clone().requires_grad_(True) for k in ('a', 'b', 'c', 'net'))
  o, x = O.elbo_terms(d['x'], a, b, c, d['t'], d['eps_0'], d['eps'], lambda z, g: net, param, cfg,
                      dtype=dtype, return_aux=True)
  B = a.shape[0]
  L = (gL.to(dtype) * o.loss_diff).sum() + (zbar.to(dtype) * x['z_t'].reshape(B, -1)).sum()
  L = L + (gbar.to(dtype) * O.score_model_gt(x['g_t'], cfg).reshape(B)).sum()
  grads = torch.autograd.grad(L, [a, b, c, net])
  terms = dict(loss_recon=o.loss_recon.detach(), loss_diff=o.loss_diff.detach(),
               loss_klz_prior=x['loss_klz_prior'].detach(),
               g_net=x['g_t'].detach().reshape(B, -1).mean(dim=1))
  return terms, x['z_t'].detach().reshape(B, -1), dict(zip(('a_bar', 'b_bar', 'c_bar', 'n_bar'), grads))


@pytest.mark.parametrize('D', [4, 100, 768, 3076, 12288])
@pytest.mark.parametrize('param', [O.MODE_EPS, O.MODE_VEL])
def test_other_row_lengths(cuda_device, D, param):
  """The descriptor's dim is not tied to 32x32x3: rows shorter than a CTA's stride, rows that
  are not a multiple of it, and 64x64x3 rows, in every launch shape (3 / 64 / 200 / 700 rows),
  against the oracle on the same inputs.  The bars are the usual ones (per-example terms 1e-5,
  gradients 1e-4 per-row L2, against the float64 oracle); where the float32 ORACLE itself is
  further than that from float64 -- short rows average less per-pixel rounding away than 3072
  sub-pixels do, and single rows cancel badly in the velocity gradient -- four times the float32
  oracle's own distance for that row (or its worst row's distance) is allowed instead, and the
  whole tensor must still meet the bar."""
  from mulan_b200 import ops
  for B, seed in ((3, 5), (64, 8), (200, 6), (700, 7)):
    if B * D > 1_000_000:       # keep the float64 autograd oracle in seconds
      continue
    inp = O.synth_inputs(B, seed, D=D)
    rng = np.random.default_rng(seed)
    gL = torch.from_numpy(rng.uniform(0.5, 1.5, B).astype(np.float32)) / (B * D * math.log(2))
    zbar = torch.from_numpy(rng.standard_normal((B, D)).astype(np.float32)) * 1e-4
    gbar = torch.from_numpy(rng.standard_normal((B,)).astype(np.float32)) * 1e-3
    t32, z32, g32 = _oracle_terms_and_grads(inp, param, gL, zbar, gbar, torch.float32)
    t64, z64, g64 = _oracle_terms_and_grads(inp, param, gL, zbar, gbar, torch.float64)
    g = {k: v.to(cuda_device).contiguous() for k, v in inp.items()}
    desc = ops.Desc(param=param, dim=D)
    got = step(desc, g, gL=gL.to(cuda_device), z_bar=zbar.to(cuda_device),
               g_bar=gbar.to(cuda_device))
    for name in ('loss_recon', 'loss_klz_prior', 'loss_diff', 'g_net'):
      e_got = (got[name].cpu().double() - t64[name]).abs()
      e_ref = (t32[name].double() - t64[name]).abs().max()
      assert bool((e_got <= 1e-5 * t64[name].abs() + 4 * e_ref + 1e-7).all()), \
          (name, B, D, float(e_got.max()), float(e_ref))
    checks = [('z_t', got['z_t'], z32, z64, 1e-5)]
    checks += [(n, got[n], g32[n], g64[n], 1e-4) for n in ('a_bar', 'b_bar', 'c_bar', 'n_bar')]
    for name, have, w32, w64, bar in checks:
      e_got, e_ref = _rows_l2(have, w64), _rows_l2(w32, w64)
      assert bool((e_got <= torch.clamp(4 * e_ref, min=max(bar, float(e_ref.max())))).all()), \
          (name, B, D, float(e_got.max()), float(e_ref.max()))
      assert rel_l2(have.cpu(), w64) < bar, (name, B, D)
