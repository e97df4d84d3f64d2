#!/usr/bin/env python
"""Golden vectors for the auxiliary-latent variants neither shipped config selects, generated
by EXECUTING THE REFERENCE'S OWN `_get_embedding_and_kl_z`
(/root/reference/ldm/model_mulan_epsilon.py:257-271 -> :195-219, :233-252, :264-270) on the
torch-backed jax/flax stand-in.  Build container only; writes tests/golden/latent.npz.

  python tests/golden/make_golden_latent.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import golden_inputs as GI  # noqa: E402
import jaxshim  # noqa: E402

jaxshim.install('/root/reference')
from ldm import model_mulan_epsilon, model_vdm  # noqa: E402

SEED, B, L = 401, 6, 50


def latent_inputs():
  r = np.random.default_rng(SEED)
  f32 = np.float32
  return dict(logits=(2.0 * r.standard_normal((B, L))).astype(f32),
              gumbel=r.gumbel(size=(B, L)).astype(f32),
              mu=r.standard_normal((B, L)).astype(f32),
              var=np.logaddexp(r.standard_normal((B, L)), 0.0).astype(f32),
              eps_z=r.standard_normal((B, L)).astype(f32),
              emb_bar=r.standard_normal((B, L)).astype(f32),
              kl_bar=r.standard_normal((B,)).astype(f32))


def run(case, dtype):
  torch.set_default_dtype(dtype)
  cfg = dict(GI.CONFIG)
  step = 0
  if case == 'topk_gumbel':
    cfg.update(latent_type='topk', topk_noise_type='gumbel')
  elif case.startswith('gumbel'):
    cfg.update(latent_type='gumbel')
    step = 30000.0 if case == 'gumbel_step30000' else 0.0       # tau = exp(-0.3) = 0.74
  else:
    cfg.update(latent_type='gaussian')
  vdm = model_mulan_epsilon.VDM(model_vdm.VDMConfig(**cfg))
  inp = latent_inputs()
  tt = lambda v: torch.from_numpy(np.asarray(v)).to(dtype)
  leaf = lambda v: tt(v).requires_grad_(True)
  if case == 'gaussian':
    mu, var = leaf(inp['mu']), leaf(inp['var'])
    vdm.encoder_model = lambda f, det: (mu, var)
    jaxshim.set_draws([('normal', inp['eps_z'])])
    leaves = {'mu': mu, 'var': var}
  else:
    logits = leaf(inp['logits'])
    vdm.encoder_model = lambda f, det: logits
    jaxshim.set_draws([('gumbel', inp['gumbel'])])
    leaves = {'logits': logits}
  emb, kl = vdm._get_embedding_and_kl_z(torch.zeros(B, 32, 32, 3), step=step, deterministic=False)
  loss = (emb * tt(inp['emb_bar'])).sum() + (kl * tt(inp['kl_bar'])).sum()
  grads = torch.autograd.grad(loss, list(leaves.values()))
  out = {'emb': emb, 'kl': kl}
  for n, g in zip(leaves, grads):
    out['grad_' + n] = g
  return {k: v.detach().numpy() for k, v in out.items()}


def main():
  res = {}
  for case in ('topk_gumbel', 'gumbel_step0', 'gumbel_step30000', 'gaussian'):
    for dtype, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
      for k, v in run(case, dtype).items():
        res[f'{case}_{tag}_{k}'] = v
  torch.set_default_dtype(torch.float32)
  path = os.path.join(os.environ.get('MULAN_GOLDEN_OUT', HERE), 'latent.npz')
  np.savez_compressed(path, **res)
  print(f'latent.npz: {os.path.getsize(path) / 1024:.0f} KiB, {len(res)} arrays')


if __name__ == '__main__':
  main()
