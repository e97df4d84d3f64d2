"""GPU tests of the fused / alternative code paths: value-and-grad post kernel, row rescale,
and the opt-in TMA-pipelined fwd_pre; each must reproduce the plain path bit for bit (or to
rounding where the arithmetic order legitimately differs)."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import mulan_oracle as O

pytestmark = pytest.mark.gpu

MODES = {'eps': O.MODE_EPS, 'vel': O.MODE_VEL, 'vel_from_eps': O.MODE_VEL_FROM_EPS}
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dev(d, device):
  return {k: v.to(device).contiguous() for k, v in d.items()}


@pytest.mark.parametrize('mode', list(MODES))
def test_fwd_bwd_post_equals_separate_passes(cuda_device, mode):
  from mulan_b200 import ops
  B = 64
  inp = O.synth_inputs(B, 91)
  g = _dev(inp, cuda_device)
  desc = ops.Desc(param=MODES[mode])
  rng = np.random.default_rng(1)
  gL = torch.from_numpy(rng.uniform(0.5, 1.5, B).astype(np.float32)).to(cuda_device) * 1e-6
  w = None
  if mode == 'eps':
    w = ops.fwd_pre(desc, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'], g['eps'])['w']
  args = (g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'], w)
  diff = ops.fwd_post(desc, *args)
  nbar = ops.bwd_post(desc, *args, gL)
  diff2, nbar2 = ops.fwd_bwd_post(desc, *args, gL)
  assert torch.equal(diff, diff2)
  assert torch.equal(nbar, nbar2)


def test_scale_rows(cuda_device):
  from mulan_b200 import ops
  B = 9
  v = torch.randn(B, 3072, device=cuda_device)
  v0 = v.clone()
  num = torch.rand(B, device=cuda_device) + 0.5
  den = num.clone()
  den[3] = num[3] * 2.0
  den[7] = num[7] * 0.25
  ops.scale_rows(v, num, den)
  want = v0.clone()
  want[3] *= 0.5
  want[7] *= 4.0
  assert torch.equal(v, want)          # untouched rows bit-identical, scaled rows exact (2^k)


@pytest.mark.parametrize('scale', [1.0, 3.0])
def test_fused_value_and_grad_in_model(cuda_device, scale):
  """VDM with fused_value_and_grad on/off gives identical losses and gradients, also when the
  upstream gradient is NOT the one the fused kernel assumed (loss scaled by 3)."""
  from mulan_b200.model import VDM, VDMConfig, loss_fn
  sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
  import golden_inputs as GI
  dev = cuda_device
  B = 16
  rng = np.random.default_rng(2)
  images = torch.from_numpy(rng.integers(0, 256, (B, 32, 32, 3), dtype=np.uint8)).to(dev)
  enc_w = torch.from_numpy(GI.encoder_weights(5)).to(dev)
  res = []
  for fused in (False, True):
    w1 = torch.tensor(0.8, device=dev, requires_grad=True)
    encoder = lambda f, det: f.reshape(f.shape[0], -1)[:, :256] @ enc_w
    score = lambda z, g_, cond, det: w1 * z + 0.01 * g_.reshape(-1, 1, 1, 1)
    model = VDM(VDMConfig(vdm_type='mulan_velocity', velocity_from_epsilon=True), encoder,
                score).to(dev)
    model.gamma.load_flax(GI.mlp_weights(6))
    model.fused_value_and_grad = fused
    gen = torch.Generator(device=dev).manual_seed(11)
    draws = model.make_draws(B, dev, gen)
    bpd, _ = loss_fn(model, {'images': images}, draws=draws)
    (scale * bpd).backward()
    res.append((bpd.item(), w1.grad.clone(), model.gamma.dense_out_b.bias.grad.clone()))
  assert res[0][0] == res[1][0]
  assert abs(res[0][1].item() - res[1][1].item()) <= 2e-6 * abs(res[0][1].item())
  num = (res[0][2] - res[1][2]).norm().item()
  assert num <= 2e-6 * res[0][2].norm().item()


def test_fwd_pre_variants_are_bit_identical(cuda_device):
  """The fwd_pre variants -- generic constants (MULAN_NO_BAKED=1), constants baked as immediates
  for the shipped configuration (default), and the opt-in cp.async.bulk + mbarrier pipeline
  (MULAN_FWD_PRE_TMA=1) -- run the same arithmetic per sub-pixel and the same reduction tree,
  so every output must be bit-identical."""
  code = r'''
import sys, torch, numpy as np
sys.path.insert(0, %r)
from mulan_b200 import ops
from oracle import mulan_oracle as O
dev = torch.device('cuda:0')
out = {}
for B in (3, 700, 2500):
  inp = O.synth_inputs(B, 7, group=128)
  g = {k: v.to(dev).contiguous() for k, v in inp.items()}
  for gt in (0, 1):
    r = ops.fwd_pre(ops.Desc(gt_mode=gt), g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'],
                    g['eps'], save_w=(gt == 0))
    for k, v in r.items():
      if v is not None:
        out['%%d_%%d_%%s' %% (B, gt, k)] = v.cpu().numpy()
np.savez(sys.argv[1], **out)
''' % ROOT
  import tempfile
  outs = []
  for tma, nobaked in (('0', '1'), ('0', '0'), ('1', '0'), ('1', '1')):
    with tempfile.NamedTemporaryFile(suffix='.npz', delete=False) as f:
      path = f.name
    env = dict(os.environ, MULAN_FWD_PRE_TMA=tma, MULAN_NO_BAKED=nobaked)
    subprocess.run([sys.executable, '-c', code, path], check=True, env=env)
    outs.append(np.load(path))
    os.unlink(path)
  # outs: [generic constants, baked (default), TMA + baked, TMA + generic constants]
  for i, other in enumerate(outs[1:], 1):
    assert set(outs[0].files) == set(other.files)
    for k in outs[0].files:
      per_row = k.split('_', 2)[2] in ('loss_recon', 'loss_klz_prior', 'var_sums') or \
          (k.split('_', 2)[2] == 'g_net' and k.split('_')[1] == '0')
      if per_row and i >= 2:
        # the TMA pipeline runs 256-thread CTAs, the direct-load default 128-thread ones: the
        # per-row sums differ by float32 summation order, the per-pixel outputs do not
        np.testing.assert_allclose(outs[0][k], other[k], rtol=2e-6, atol=1e-6, err_msg=k)
      else:
        assert np.array_equal(outs[0][k], other[k]), k
  for k in outs[2].files:                      # the two TMA builds agree bit for bit
    assert np.array_equal(outs[2][k], outs[3][k]), k


def test_vfe_eps_form_matches_literal_formula(cuda_device):
  """velocity_from_epsilon runs the epsilon form by default (mulan_kernel_param); the literal
  formula of ldm/model_mulan_velocity.py:246-260 stays reachable with MULAN_VFE_LITERAL=1.
  Both must agree with each other and with the oracle's literal float32 evaluation."""
  code = r'''
import sys, math, torch, numpy as np
sys.path.insert(0, %r)
from mulan_b200 import ops, _lib
from oracle import mulan_oracle as O
dev = torch.device('cuda:0')
B = 40
inp = O.synth_inputs(B, 17)
g = {k: v.to(dev).contiguous() for k, v in inp.items()}
desc = ops.Desc(param=2)
ws = ops.ElboWorkspace(desc, B, dev)
gL = torch.full((B,), 1.0 / (B * 3072 * math.log(2.0)), device=dev)
rng = np.random.default_rng(5)
zb = torch.from_numpy(1e-4 * rng.standard_normal((B, 3072)).astype(np.float32)).to(dev)
gb = torch.from_numpy(1e-3 * rng.standard_normal(B).astype(np.float32)).to(dev)
args = (g['x'], g['a'], g['b'], g['c'], g['t'])
ws.fwd_pre(*args, g['eps_0'], g['eps'])
ws.fwd_bwd_post(*args, g['eps'], g['net'], gL)
ws.bwd_pre(*args, g['eps'], g['net'], zb, gb, gL)
torch.cuda.synchronize()
np.savez(sys.argv[1], kparam=_lib.kernel_param(2), saved_w=ws.w is not None,
         loss_diff=ws.loss_diff.cpu().numpy(), n_bar=ws.n_bar.cpu().numpy(),
         a_bar=ws.a_bar.cpu().numpy(), b_bar=ws.b_bar.cpu().numpy(), c_bar=ws.c_bar.cpu().numpy(),
         zb=zb.cpu().numpy(), gb=gb.cpu().numpy(), gL=gL.cpu().numpy())
''' % ROOT
  import tempfile
  res = {}
  for literal in ('0', '1'):
    with tempfile.NamedTemporaryFile(suffix='.npz', delete=False) as f:
      path = f.name
    subprocess.run([sys.executable, '-c', code, path], check=True,
                   env=dict(os.environ, MULAN_VFE_LITERAL=literal))
    res[literal] = dict(np.load(path))
    os.unlink(path)
  assert int(res['0']['kparam']) == 0 and bool(res['0']['saved_w'])
  assert int(res['1']['kparam']) == 2 and not bool(res['1']['saved_w'])
  l2 = lambda a, b: np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64))
  e, l = res['0'], res['1']
  assert np.max(np.abs(e['loss_diff'] - l['loss_diff']) / np.abs(l['loss_diff'])) < 1e-5
  assert l2(e['n_bar'], l['n_bar']) < 1e-5
  for k in ('a_bar', 'b_bar', 'c_bar'):
    assert l2(e[k], l[k]) < 1e-4, k
  # the oracle's literal formula, float64, for L = sum gL loss_diff + <zb, z_t> + <gb, g_net>
  B = 40
  inp = O.synth_inputs(B, 17)
  cfg = O.OracleConfig()
  i = {k: (v.double() if v.is_floating_point() else v) for k, v in inp.items()}
  a, b, c, net = (i[k].clone().requires_grad_(True) for k in ('a', 'b', 'c', 'net'))
  out, aux = O.elbo_terms(i['x'], a, b, c, i['t'], i['eps_0'], i['eps'], lambda z, g: net,
                          O.MODE_VEL_FROM_EPS, cfg, dtype=torch.float64, return_aux=True)
  L = (torch.from_numpy(e['gL']).double() * out.loss_diff).sum()
  L = L + (torch.from_numpy(e['zb']).double() * aux['z_t'].reshape(B, -1)).sum()
  L = L + (torch.from_numpy(e['gb']).double() * O.score_model_gt(aux['g_t'], cfg).reshape(B)).sum()
  ga, gb_, gc, gn = torch.autograd.grad(L, [a, b, c, net])
  for r in (e, l):
    assert np.max(np.abs(r['loss_diff'] - out.loss_diff.detach().numpy())
                  / out.loss_diff.detach().numpy()) < 1e-5
    assert l2(r['n_bar'], gn.reshape(B, -1).numpy()) < 1e-4
    for k, want in (('a_bar', ga), ('b_bar', gb_), ('c_bar', gc)):
      assert l2(r[k], want.reshape(B, -1).numpy()) < 1e-4, k
