#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE
(/root/reference/ldm/model_mulan_epsilon.py, model_mulan_velocity.py, model_vdm.py) on the
torch-backed jax/flax stand-in of jaxshim.py.  Run in the build container only
(/root/reference does not exist on the GPU box); the .npz outputs are committed.

  python tests/golden/make_golden.py

Two fixture families, float32 and float64 each:
  glue_{eps,vel,vfe}     VDM.__call__ with the encoder logits and the schedule coefficients
                         (a, b, c) injected -> loss terms, z_t, g_t, bpd and the gradients of
                         bpd w.r.t. a, b, c, logits and the stand-in denoiser's weights.
  full_{eps,vfe}         the same call with the reference's own _compute_coefficients running
                         its five Dense layers on seeded weights.
  glue_eps_T1000         the epsilon model's discrete-time branch (sm_n_timesteps=1000).
  glue_vel_ldm           the velocity model with unet_type='ldm' (per-pixel gamma_t to the denoiser).
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
from ldm import model_mulan_epsilon, model_mulan_velocity, model_vdm  # noqa: E402


def loss_fn_bpd(outputs, image_shape=(32, 32, 3)):
  """Experiment_VDM.loss_fn, ldm/experiment_vdm.py:62-66 (that module cannot be imported:
  it pulls in clu/optax/tensorflow), restated."""
  rescale_to_bpd = 1. / (np.prod(image_shape) * np.log(2.))
  bpd_latent = torch.mean(outputs.loss_klz) * rescale_to_bpd
  bpd_recon = torch.mean(outputs.loss_recon) * rescale_to_bpd
  bpd_diff = torch.mean(outputs.loss_diff) * rescale_to_bpd
  return bpd_recon + bpd_latent + bpd_diff


def build_vdm(kind, **overrides):
  cfg = dict(GI.CONFIG)
  cfg.update(overrides)
  if kind == 'vfe':
    cfg['velocity_from_epsilon'] = True
  config = model_vdm.VDMConfig(**cfg)
  mod = model_mulan_epsilon if kind == 'eps' else model_mulan_velocity
  return mod.VDM(config)


def run(kind, seed, B, dtype, full, **overrides):
  torch.set_default_dtype(dtype)
  vdm = build_vdm(kind, **overrides)
  inp = GI.glue_inputs(seed, B)
  tt = lambda v: torch.from_numpy(np.asarray(v)).to(dtype)
  leaf = lambda v: tt(v).requires_grad_(True)
  captured = {}

  w1, w2, w3 = leaf(inp['w1']), leaf(inp['w2']), leaf(inp['w3'])
  noise = tt(inp['noise'])

  def score_model(z, g_t, conditioning, deterministic, time=False):
    captured['z_t'], captured['g_net'], captured['cond'] = z, g_t, conditioning
    g = g_t.reshape(-1, 1, 1, 1) if g_t.ndim == 1 else g_t     # unet_type 'vdm' | 'ldm'
    return (w1 * z + w2 * g
            + w3 * conditioning.sum(dim=1).reshape(-1, 1, 1, 1) + noise)
  vdm.score_model = score_model

  grads_of = {'w1': w1, 'w2': w2, 'w3': w3}
  if full:
    W = GI.mlp_weights(seed + 1000)
    We = leaf(GI.encoder_weights(seed + 2000))
    vdm.encoder_model = lambda orig_f, deterministic: orig_f.reshape(B, -1)[:, :256] @ We
    for layer, name in ((vdm.gamma.l1, 'dense_1'), (vdm.gamma.l2, 'dense_2'),
                        (vdm.gamma.l3_a, 'dense_out_a'), (vdm.gamma.l3_b, 'dense_out_b'),
                        (vdm.gamma.l3_c, 'dense_out_c')):
      layer.kernel = tt(W[name + '/kernel'])
      layer.bias = leaf(W[name + '/bias'])
      grads_of[name + '/bias'] = layer.bias
    grads_of['We'] = We
    orig_cc = vdm.gamma._compute_coefficients
    def cc(embedding):
      out = orig_cc(embedding)
      captured['abc'] = out
      return out
    vdm.gamma._compute_coefficients = cc
  else:
    logits = leaf(inp['logits'])
    a, b, c = leaf(inp['a']), leaf(inp['b']), leaf(inp['c'])
    vdm.encoder_model = lambda orig_f, deterministic: logits
    vdm.gamma._compute_coefficients = lambda embedding: (a, b, c)
    grads_of.update(a=a, b=b, c=c, logits=logits)

  # make_rng('sample') call order: t0, gamma noise, eps_0, eps (epsilon.py:287,222,315,327)
  jaxshim.set_draws([('uniform', inp['t0']), ('gamma', inp['G']), ('normal', inp['eps_0']),
                     ('normal', inp['eps'])])
  images = torch.from_numpy(inp['images'])
  out = vdm(images, labels=torch.zeros(B), conditioning=torch.zeros(B), step=0,
            deterministic=False)
  bpd = loss_fn_bpd(out)
  names = list(grads_of)
  grads = torch.autograd.grad(bpd, [grads_of[n] for n in names], allow_unused=True)
  res = dict(loss_recon=out.loss_recon, loss_klz=out.loss_klz, loss_diff=out.loss_diff,
             var_0=out.var_0, var_1=out.var_1, bpd=bpd, z_t=captured['z_t'],
             g_net=captured['g_net'], embedding=captured['cond'])
  if full:
    res.update(a=captured['abc'][0], b=captured['abc'][1], c=captured['abc'][2])
  for n, g in zip(names, grads):
    res['grad_' + n.replace('/', '_')] = torch.zeros_like(grads_of[n]) if g is None else g
  return {k: v.detach().cpu().numpy() for k, v in res.items()}


def main():
  B = 4
  jobs = [('glue_eps', 'eps', 101, False, {}), ('glue_vel', 'vel', 102, False, {}),
          ('glue_vfe', 'vfe', 103, False, {}), ('full_eps', 'eps', 201, True, {}),
          ('full_vfe', 'vfe', 203, True, {}),
          ('glue_eps_T1000', 'eps', 104, False, {'sm_n_timesteps': 1000}),
          # unet_type='ldm': the denoiser receives the per-pixel gamma_t (epsilon.py:277-278)
          ('glue_vel_ldm', 'vel', 105, False, {'unet_type': 'ldm'})]
  for name, kind, seed, full, ov in jobs:
    out = {}
    for dtype, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
      r = run(kind, seed, B, dtype, full, **ov)
      for k, v in r.items():
        if tag == 'f64' and k in ('z_t', 'embedding', 'a', 'b', 'c'):
          continue   # keep the fixtures small
        out[f'{tag}_{k}'] = v
    out['seed'], out['B'] = np.int64(seed), np.int64(B)
    path = os.path.join(os.environ.get('MULAN_GOLDEN_OUT', HERE), name + '.npz')
    np.savez_compressed(path, **out)
    print(f'{name}: bpd f32 {out["f32_bpd"]:.7f} f64 {out["f64_bpd"]:.7f}  -> '
          f'{os.path.getsize(path) / 1024:.0f} KiB')
  torch.set_default_dtype(torch.float32)


if __name__ == '__main__':
  main()
