#!/usr/bin/env python
"""Reference-source golden vectors at BASELINE.json's batch sizes, on edge-case inputs and for
the dense-VLB tile -- generated, like make_golden.py's, by EXECUTING THE REFERENCE'S OWN SOURCE
(/root/reference/ldm/model_mulan_*.py, model_vdm.py) on the jaxshim stand-in.  Build container
only; the .npz outputs are committed.

  python tests/golden/make_golden_configs.py

  cfg1_eps_B8, cfg2_eps_B128, cfg3_vel_B128, cfg4_vfe_B256
        VDM.__call__ + loss_fn at the per-GPU batch of BASELINE.json configs[0..3]: per-example
        loss terms, bpd, g_net in full; z_t and the gradients of bpd w.r.t. a, b, c COMPACT
        (per-row sum, norm, 8 seeded projections -- golden_inputs.compact), float32 and float64.
  edge_eps, edge_vel, edge_vfe
        B = 8, one edge per row (golden_inputs.edge_inputs), antithetic_time_sampling=False
        so t in {0, 1, 1e-6, ...} is injected; stored in full (float32) + compact (float64).
  dense_vel, dense_vfe
        notebook_utils.eval_bpd_dense_sampling's per-image call (:183-185): ONE image tiled over
        128 antithetic timesteps, is_train=False, same draws for every image -> per-image bpd
        and the per-row loss terms.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import golden_inputs as GI  # noqa: E402
import make_golden as MG  # noqa: E402  (installs jaxshim, imports the reference modules)
import jaxshim  # noqa: E402


def run(kind, inp, dtype, deterministic=False, **overrides):
  """One VDM.__call__ of the reference source with (a, b, c) and the logits injected."""
  torch.set_default_dtype(dtype)
  vdm = MG.build_vdm(kind, **overrides)
  B = inp['a'].shape[0]
  tt = lambda v: torch.from_numpy(np.asarray(v)).to(dtype)
  leaf = lambda v: tt(v).requires_grad_(True)
  cap = {}
  w1, w2, w3 = tt(inp['w1']), tt(inp['w2']), tt(inp['w3'])
  noise = tt(inp['noise'])

  def score_model(z, g_t, conditioning, deterministic, time=False):
    cap['z_t'], cap['g_net'] = z, g_t
    g = g_t.reshape(-1, 1, 1, 1) if g_t.ndim == 1 else g_t
    return w1 * z + w2 * g + w3 * conditioning.sum(dim=1).reshape(-1, 1, 1, 1) + noise
  vdm.score_model = score_model
  logits = leaf(inp['logits'])
  a, b, c = leaf(inp['a']), leaf(inp['b']), leaf(inp['c'])
  vdm.encoder_model = lambda orig_f, deterministic: logits
  vdm.gamma._compute_coefficients = lambda embedding: (a, b, c)
  t_draw = inp['t'] if 't' in inp else inp['t0']      # per-row t: antithetic sampling off
  jaxshim.set_draws([('uniform', t_draw), ('gamma', inp['G']), ('normal', inp['eps_0']),
                     ('normal', inp['eps'])])
  images = torch.from_numpy(np.asarray(inp['images']))
  out = vdm(images, labels=torch.zeros(B), conditioning=torch.zeros(B), step=0,
            deterministic=deterministic)
  bpd = MG.loss_fn_bpd(out)
  ga, gb, gc, gl = torch.autograd.grad(bpd, [a, b, c, logits], allow_unused=True)
  n = lambda v: v.detach().cpu().numpy()
  return dict(loss_recon=n(out.loss_recon), loss_klz=n(out.loss_klz), loss_diff=n(out.loss_diff),
              var_0=n(out.var_0), var_1=n(out.var_1), bpd=n(bpd), z_t=n(cap['z_t']),
              g_net=n(cap['g_net']), grad_a=n(ga), grad_b=n(gb), grad_c=n(gc),
              grad_logits=n(gl))


SMALL = ('loss_recon', 'loss_klz', 'loss_diff', 'var_0', 'var_1', 'bpd', 'g_net', 'grad_logits')
BIG = ('z_t', 'grad_a', 'grad_b', 'grad_c')


def pack(r32, r64, full32: bool):
  out = {}
  for tag, r in (('f32', r32), ('f64', r64)):
    for k in SMALL:
      out[f'{tag}_{k}'] = r[k]
    for k in BIG:
      for s, v in GI.compact(r[k]).items():
        out[f'{tag}_{k}_{s}'] = v
  if full32:
    for k in BIG:
      out[f'f32_{k}'] = r32[k]
  return out


def save(name, out):
  path = os.path.join(os.environ.get('MULAN_GOLDEN_OUT', HERE), name + '.npz')
  np.savez_compressed(path, **out)
  print(f'{name}: bpd f32 {float(np.mean(out["f32_bpd"])):.7f} f64 '
        f'{float(np.mean(out["f64_bpd"])):.7f} -> {os.path.getsize(path) / 1024:.0f} KiB')


def main():
  for name, (kind, seed, B) in GI.SIZE_CASES.items():
    inp = GI.glue_inputs(seed, B)
    out = pack(run(kind, inp, torch.float32), run(kind, inp, torch.float64), full32=False)
    out['seed'], out['B'] = np.int64(seed), np.int64(B)
    save(name, out)
  for name, (kind, seed) in GI.EDGE_CASES.items():
    inp = GI.edge_inputs(seed)
    ov = dict(antithetic_time_sampling=False)
    out = pack(run(kind, inp, torch.float32, **ov), run(kind, inp, torch.float64, **ov),
               full32=True)
    out['seed'], out['B'] = np.int64(seed), np.int64(8)
    save(name, out)
  for name, (kind, seed) in GI.DENSE_CASES.items():
    base, per_image = GI.dense_inputs(seed)
    out = {}
    for tag, dtype in (('f32', torch.float32), ('f64', torch.float64)):
      rows, bpds = [], []
      for im in per_image:
        inp = dict(base, images=np.tile(im['image'], (GI.DENSE_T, 1, 1, 1)), a=im['a'],
                   b=im['b'], c=im['c'], logits=im['logits'])
        r = run(kind, inp, dtype, deterministic=True)
        rows.append(np.stack([r['loss_recon'], r['loss_klz'], r['loss_diff']], axis=1))
        bpds.append(r['bpd'])
      out[f'{tag}_rows'] = np.stack(rows)        # [images, T, 3]
      out[f'{tag}_bpd'] = np.asarray(bpds)       # [images]
    out['seed'] = np.int64(seed)
    save(name, out)
  torch.set_default_dtype(torch.float32)


if __name__ == '__main__':
  main()
