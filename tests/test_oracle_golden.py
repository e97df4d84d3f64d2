"""Pin the CPU oracle against the golden vectors produced by executing the REFERENCE'S OWN
SOURCE (tests/golden/make_golden.py; /root/reference is not needed at test time).

The f32 goldens and the f32 oracle run the same expression order on the same torch-CPU
kernels, so they agree to rounding of pow / forward-mode jvp details (<= 1e-6); the f64
goldens agree to 1e-12.
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import mulan_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
import golden_inputs as GI  # noqa: E402

MODES = {'eps': O.MODE_EPS, 'vel': O.MODE_VEL, 'vfe': O.MODE_VEL_FROM_EPS}


def load(name):
  return np.load(os.path.join(HERE, 'golden', name + '.npz'))


def run_oracle(kind, seed, B, dtype, full, T=0, unet_type='vdm'):
  inp = GI.glue_inputs(seed, B)
  cfg = O.OracleConfig(sm_n_timesteps=T, unet_type=unet_type)
  tt = lambda v: torch.from_numpy(np.asarray(v)).to(dtype)
  leaf = lambda v: tt(v).requires_grad_(True)
  w1, w2, w3 = leaf(inp['w1']), leaf(inp['w2']), leaf(inp['w3'])
  noise = tt(inp['noise'])
  cap = {}

  def score_fn(z, g, cond):
    cap['z_t'], cap['g_net'], cap['cond'] = z, g, cond
    gg = g.reshape(-1, 1, 1, 1) if g.ndim == 1 else g
    return (w1 * z + w2 * gg
            + w3 * cond.sum(dim=1).reshape(-1, 1, 1, 1) + noise)

  leaves = {'w1': w1, 'w2': w2, 'w3': w3}
  if full:
    W = {k: tt(v) for k, v in GI.mlp_weights(seed + 1000).items()}
    for k in list(W):
      if k.endswith('/bias'):
        W[k] = W[k].requires_grad_(True)
        leaves[k] = W[k]
    We = leaf(GI.encoder_weights(seed + 2000))
    leaves['We'] = We
    encoder_fn = lambda f: f.reshape(B, -1)[:, :256] @ We
    def coeff_fn(emb):
      cap['abc'] = O.compute_coefficients(W, emb)
      return cap['abc']
  else:
    logits = leaf(inp['logits'])
    a, b, c = leaf(inp['a']), leaf(inp['b']), leaf(inp['c'])
    leaves.update(a=a, b=b, c=c, logits=logits)
    encoder_fn = lambda f: logits
    coeff_fn = lambda emb: (a, b, c)
  draws = dict(t0=float(inp['t0']) if dtype == torch.float64 else inp['t0'], G=tt(inp['G']),
               eps_0=tt(inp['eps_0']), eps=tt(inp['eps']))
  draws['t0'] = tt(inp['t0'])
  out = O.vdm_call(torch.from_numpy(inp['images']), draws, coeff_fn, encoder_fn, score_fn,
                   MODES[kind], cfg, dtype=dtype)
  bpd, _ = O.loss_fn_bpd(out)
  names = list(leaves)
  grads = torch.autograd.grad(bpd, [leaves[n] for n in names], allow_unused=True)
  res = dict(loss_recon=out.loss_recon, loss_klz=out.loss_klz, loss_diff=out.loss_diff,
             var_0=out.var_0, var_1=out.var_1, bpd=bpd, z_t=cap['z_t'], g_net=cap['g_net'],
             embedding=cap['cond'])
  if full:
    res.update(a=cap['abc'][0], b=cap['abc'][1], c=cap['abc'][2])
  for n, g in zip(names, grads):
    res['grad_' + n.replace('/', '_')] = torch.zeros_like(leaves[n]) if g is None else g
  return {k: v.detach().numpy() for k, v in res.items()}


def _close(got, want, rtol, atol=0.0):
  got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
  err = np.abs(got - want)
  tol = atol + rtol * np.abs(want)
  assert np.all(err <= tol), f'max err {err.max():.3e} (|want| max {np.abs(want).max():.3e})'


def _rel_l2(got, want):
  got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
  return np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-300)


CASES = [('glue_eps', 'eps', False, 0), ('glue_vel', 'vel', False, 0),
         ('glue_vfe', 'vfe', False, 0), ('full_eps', 'eps', True, 0),
         ('full_vfe', 'vfe', True, 0), ('glue_eps_T1000', 'eps', False, 1000)]


@pytest.mark.parametrize('name,kind,full,T', CASES)
@pytest.mark.parametrize('tag', ['f32', 'f64'])
def test_oracle_matches_reference_source(name, kind, full, T, tag):
  g = load(name)
  seed, B = int(g['seed']), int(g['B'])
  dtype = torch.float32 if tag == 'f32' else torch.float64
  r = run_oracle(kind, seed, B, dtype, full, T)
  # f32: same op order on the same CPU kernels, only t**n (pow vs multiply chain) and the
  # jvp expansion can differ by an ulp.  f64: 1e-12.
  rt = 2e-6 if tag == 'f32' else 1e-11
  if T > 0 and tag == 'f32':
    rt = 1e-5   # expm1(g_t - g_s) with t - s = 1/T amplifies the 1-ulp t**n difference
  for k in ('loss_recon', 'loss_klz', 'loss_diff', 'bpd', 'var_0', 'var_1'):
    _close(r[k], g[f'{tag}_{k}'], rt)
  _close(r['g_net'], g[f'{tag}_g_net'], rt, atol=1e-6 if tag == 'f32' else 1e-12)
  if tag == 'f32':
    # bit-identical except rows whose t**n differs by an ulp (torch pow vs multiply chain),
    # amplified in pixels where P/S cancels
    _close(r['z_t'], g['f32_z_t'], 0, atol=5e-5)
    assert _rel_l2(r['z_t'], g['f32_z_t']) < 1e-6
    np.testing.assert_array_equal(r['embedding'] > 0.5, g['f32_embedding'] > 0.5)
    _close(r['embedding'], g['f32_embedding'], 0, atol=1e-6)
    if full:
      for k in 'abc':
        assert _rel_l2(r[k], g['f32_' + k]) < 1e-6
  gt = 1e-4 if tag == 'f32' else 1e-9
  if T > 0 and tag == 'f32':
    gt = 5e-3   # float32 cancellation in g_t - g_s (t - s = 1/T); the f64 fixture pins it at 1e-9
  for k in [k for k in g.files if k.startswith(f'{tag}_grad_')]:
    want = g[k]
    got = r[k[len(tag) + 1:]]
    if np.linalg.norm(want) == 0:
      assert np.linalg.norm(got) == 0
    else:
      assert _rel_l2(got, want) < gt, (k, _rel_l2(got, want))


def test_f32_golden_close_to_f64_golden():
  """The tolerances of the north_star are meaningful only if float32 itself is that close to
  exact arithmetic: document how close the reference's f32 evaluation is."""
  for name, *_ in CASES[:5]:
    g = load(name)
    assert abs(float(g['f32_bpd']) - float(g['f64_bpd'])) < 1e-4   # BPD tolerance
    for k in ('loss_klz', 'loss_diff'):
      _close(g['f32_' + k], g['f64_' + k], 2e-5)
    _close(g['f32_loss_recon'], g['f64_loss_recon'], 1e-4)   # z_0 rounding amplified by e^{-g0/2}


@pytest.mark.parametrize('name,full', [('glue_vfe', False), ('full_vfe', True)])
@pytest.mark.parametrize('tag', ['f32', 'f64'])
def test_vfe_is_eps(name, full, tag, monkeypatch):
  """velocity_from_epsilon == the epsilon loss, values AND gradients: the oracle run in
  MODE_EPS reproduces the goldens the reference's own velocity_from_epsilon source produced
  (ldm/model_mulan_velocity.py:246-260).  This is what lets the kernels evaluate the epsilon
  form for MULAN_PARAM_VEL_FROM_EPS (include/mulan_b200.h: mulan_kernel_param)."""
  g = load(name)
  seed, B = int(g['seed']), int(g['B'])
  monkeypatch.setitem(MODES, 'vfe', O.MODE_EPS)
  dtype = torch.float32 if tag == 'f32' else torch.float64
  r = run_oracle('vfe', seed, B, dtype, full)
  _close(r['loss_diff'], g[f'{tag}_loss_diff'], 1e-6 if tag == 'f32' else 1e-12)
  _close(r['bpd'], g[f'{tag}_bpd'], 1e-6 if tag == 'f32' else 1e-12)
  for k in [k for k in g.files if k.startswith(f'{tag}_grad_')]:
    want, got = g[k], r[k[len(tag) + 1:]]
    if np.linalg.norm(want) == 0:
      assert np.linalg.norm(got) == 0
    else:
      # 1e-4 is the north_star gradient tolerance; measured 3e-6 (a,b,c), 6e-5 (logits, whose
      # float32 golden is itself that far from the float64 one)
      assert _rel_l2(got, want) < (1e-4 if tag == 'f32' else 1e-12), (k, _rel_l2(got, want))


@pytest.mark.skipif(not os.path.isdir('/root/reference/ldm'),
                    reason='the reference source only exists in the build container')
@pytest.mark.parametrize('script,files', [
    ('make_golden.py', ['glue_eps', 'glue_vel', 'glue_vfe', 'full_eps', 'full_vfe',
                        'glue_eps_T1000', 'glue_vel_ldm']),
    ('make_golden_configs.py', ['cfg1_eps_B8', 'cfg2_eps_B128', 'cfg3_vel_B128', 'cfg4_vfe_B256',
                                'edge_eps', 'edge_vel', 'edge_vfe', 'dense_vel', 'dense_vfe']),
    ('make_golden_sampler.py', ['sampler']),
    ('make_golden_latent.py', ['latent'])])
def test_committed_goldens_are_what_the_reference_source_produces(script, files, tmp_path):
  """Re-run the generator (it imports and executes /root/reference/ldm/*.py in place) into a
  scratch directory and compare with the committed fixtures: the pin is reproducible and the
  committed vectors really are the reference source's outputs."""
  import subprocess
  env = dict(os.environ, MULAN_GOLDEN_OUT=str(tmp_path))
  subprocess.run([sys.executable, os.path.join(HERE, 'golden', script)], check=True, env=env,
                 capture_output=True, timeout=900)
  for name in files:
    new, old = np.load(tmp_path / (name + '.npz')), load(name)
    assert set(new.files) == set(old.files)
    for k in old.files:
      a, b = np.asarray(new[k]), np.asarray(old[k])
      assert a.shape == b.shape and a.dtype == b.dtype, k
      if a.dtype.kind != 'f':
        assert np.array_equal(a, b), k
        continue
      # identical up to the BLAS thread count of the Dense layers in the full_* fixtures
      tol = 2e-6 if a.dtype == np.float32 else 1e-12
      scale = max(float(np.nanmax(np.abs(b))) if b.size else 0.0, 1e-30)
      assert np.array_equal(np.isnan(a), np.isnan(b)), k
      assert np.nanmax(np.abs(a - b), initial=0.0) <= tol * scale, (k, np.nanmax(np.abs(a - b)))


@pytest.mark.parametrize('tag', ['f32', 'f64'])
def test_oracle_matches_reference_source_ldm(tag):
  """unet_type='ldm' (ldm/model_mulan_epsilon.py:277-278): the denoiser receives the per-pixel
  gamma_t, so its cotangent reaches (a, b, c) per pixel instead of through a row mean."""
  g = load('glue_vel_ldm')
  seed, B = int(g['seed']), int(g['B'])
  dtype = torch.float32 if tag == 'f32' else torch.float64
  r = run_oracle('vel', seed, B, dtype, False, unet_type='ldm')
  rt = 2e-6 if tag == 'f32' else 1e-11
  for k in ('loss_recon', 'loss_klz', 'loss_diff', 'bpd', 'var_0', 'var_1'):
    _close(r[k], g[f'{tag}_{k}'], rt)
  assert r['g_net'].shape == (B, 32, 32, 3)
  _close(r['g_net'], g[f'{tag}_g_net'], rt, atol=2e-5 if tag == 'f32' else 1e-11)
  for k in [k for k in g.files if k.startswith(f'{tag}_grad_')]:
    want, got = g[k], r[k[len(tag) + 1:]]
    if np.linalg.norm(want) == 0:
      assert np.linalg.norm(got) == 0
    else:
      assert _rel_l2(got, want) < (1e-4 if tag == 'f32' else 1e-9), (k, _rel_l2(got, want))


# ----------------------------------------------------------------------------------------
# Reference-source fixtures at BASELINE.json's batch sizes, on edge inputs and for the dense-VLB
# tile (tests/golden/make_golden_configs.py)
# ----------------------------------------------------------------------------------------

def run_oracle_inputs(kind, inp, dtype, antithetic=True):
  """The oracle on an explicit input dict (a, b, c, logits injected)."""
  cfg = O.OracleConfig(antithetic_time_sampling=antithetic)
  tt = lambda v: torch.from_numpy(np.asarray(v)).to(dtype)
  leaf = lambda v: tt(v).requires_grad_(True)
  w1, w2, w3, noise = tt(inp['w1']), tt(inp['w2']), tt(inp['w3']), tt(inp['noise'])
  cap = {}

  def score_fn(z, g, cond):
    cap['z_t'], cap['g_net'] = z, g
    return w1 * z + w2 * g.reshape(-1, 1, 1, 1) + w3 * cond.sum(dim=1).reshape(-1, 1, 1, 1) + noise
  logits = leaf(inp['logits'])
  a, b, c = leaf(inp['a']), leaf(inp['b']), leaf(inp['c'])
  draws = dict(G=tt(inp['G']), eps_0=tt(inp['eps_0']), eps=tt(inp['eps']))
  if 't' in inp:
    draws['t'] = tt(inp['t'])
  else:
    draws['t0'] = tt(inp['t0'])
  out = O.vdm_call(torch.from_numpy(np.asarray(inp['images'])), draws, lambda emb: (a, b, c),
                   lambda f: logits, score_fn, MODES[kind], cfg, dtype=dtype)
  bpd, _ = O.loss_fn_bpd(out)
  ga, gb, gc, gl = torch.autograd.grad(bpd, [a, b, c, logits])
  n = lambda v: v.detach().numpy()
  return dict(loss_recon=n(out.loss_recon), loss_klz=n(out.loss_klz), loss_diff=n(out.loss_diff),
              var_0=n(out.var_0), var_1=n(out.var_1), bpd=n(bpd), z_t=n(cap['z_t']),
              g_net=n(cap['g_net']), grad_a=n(ga), grad_b=n(gb), grad_c=n(gc), grad_logits=n(gl))


def check_compact(got, g, tag, key, rtol, row_scale=None):
  """got [B, ...] against the compact statistics of a [B, D] golden tensor: per row, the error
  seen through the K seeded projections (an unbiased estimate of |got - want|) stays below
  rtol * max(|want|, row_scale)."""
  c = GI.compact(np.asarray(got, np.float64))
  want_n, want_p = g[f'{tag}_{key}_norm'], g[f'{tag}_{key}_proj']
  err = np.sqrt(np.mean((c['proj'] - want_p) ** 2, axis=1))
  scale = want_n if row_scale is None else np.maximum(want_n, row_scale)
  assert np.all(err <= rtol * scale + 1e-300), (key, float(np.max(err / np.maximum(scale, 1e-300))))
  return err / np.maximum(scale, 1e-300)


def grad_row_scale(g, tag):
  """Common scale of a row's coefficient gradients (a row where one of them vanishes
  mathematically -- a = b = 0 makes gamma independent of c -- carries only rounding noise)."""
  return np.max(np.stack([g[f'{tag}_grad_{k}_norm'] for k in 'abc']), axis=0)


@pytest.mark.parametrize('name', list(GI.SIZE_CASES))
def test_oracle_matches_reference_source_at_baseline_sizes(name):
  kind, seed, B = GI.SIZE_CASES[name]
  g = load(name)
  tags = ('f32', 'f64') if B <= 8 else ('f32',)     # keep the CPU suite short
  for tag in tags:
    r = run_oracle_inputs(kind, GI.glue_inputs(seed, B), torch.float32 if tag == 'f32'
                          else torch.float64)
    rt = 2e-6 if tag == 'f32' else 1e-11
    for k in ('loss_recon', 'loss_klz', 'loss_diff', 'bpd', 'var_0', 'var_1'):
      _close(r[k], g[f'{tag}_{k}'], rt)
    _close(r['g_net'], g[f'{tag}_g_net'], rt, atol=1e-6 if tag == 'f32' else 1e-12)
    # rows whose t**n differs by an ulp (torch pow vs XLA's multiply chain) move z_t by ~1e-6
    check_compact(r['z_t'], g, tag, 'z_t', 5e-6)
    scale = grad_row_scale(g, tag)
    for k in 'abc':
      check_compact(r['grad_' + k], g, tag, 'grad_' + k, 1e-4 if tag == 'f32' else 1e-9, scale)
    assert _rel_l2(r['grad_logits'], g[f'{tag}_grad_logits']) < (1e-4 if tag == 'f32' else 1e-9)


@pytest.mark.parametrize('name', list(GI.EDGE_CASES))
@pytest.mark.parametrize('tag', ['f32', 'f64'])
def test_oracle_matches_reference_source_on_edge_inputs(name, tag):
  kind, seed = GI.EDGE_CASES[name]
  g = load(name)
  r = run_oracle_inputs(kind, GI.edge_inputs(seed), torch.float32 if tag == 'f32'
                        else torch.float64, antithetic=False)
  rt = 2e-6 if tag == 'f32' else 1e-11
  for k in ('loss_recon', 'loss_klz', 'loss_diff', 'bpd', 'var_0', 'var_1'):
    assert np.all(np.isfinite(r[k]))
    _close(r[k], g[f'{tag}_{k}'], rt)
  _close(r['g_net'], g[f'{tag}_g_net'], rt, atol=1e-6 if tag == 'f32' else 1e-12)
  scale = grad_row_scale(g, tag)
  for k in 'abc':
    check_compact(r['grad_' + k], g, tag, 'grad_' + k, 1e-4 if tag == 'f32' else 1e-9, scale)
  if tag == 'f32':
    assert _rel_l2(r['z_t'], g['f32_z_t']) < 1e-6


@pytest.mark.parametrize('name', list(GI.DENSE_CASES))
def test_oracle_dense_vlb_tile_matches_reference_source(name):
  """O.eval_bpd_dense (ldm/notebook_utils.py:176-191) on the reference-source dense fixture."""
  kind, seed = GI.DENSE_CASES[name]
  g = load(name)
  base, per_image = GI.dense_inputs(seed)
  images = torch.from_numpy(np.concatenate([im['image'] for im in per_image]))
  state = {'i': 0, 'rows': []}

  def run_loss_fn(tiled):
    im = per_image[state['i']]
    state['i'] += 1
    r = run_oracle_inputs(kind, dict(base, images=tiled.numpy(), a=im['a'], b=im['b'], c=im['c'],
                                     logits=im['logits']), torch.float32)
    state['rows'].append(np.stack([r['loss_recon'], r['loss_klz'], r['loss_diff']], axis=1))
    return r['bpd']
  mean_bpd, bpds = O.eval_bpd_dense(images, GI.DENSE_T, run_loss_fn)
  _close(np.asarray(bpds), g['f32_bpd'], 2e-6)
  _close(np.stack(state['rows']), g['f32_rows'], 2e-6)
  assert abs(mean_bpd - float(np.mean(g['f64_bpd']))) < 1e-4
