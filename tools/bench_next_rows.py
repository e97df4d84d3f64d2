#!/usr/bin/env python
"""Per-kernel measurement of the SURVEY.md 8f "next" rows (sampler, ODE pieces, RK45 state,
JAX-compatible draws, AdamW+EMA, gradient norm, auxiliary latent): algorithmic bytes (each
declared input read once, each output written once) / CUDA-event time against the measured HBM
peak of MEASURED_PEAKS.json.  Sizes are larger than L2 (126 MB) so every launch streams from
HBM.  One JSON line per kernel on stdout.

  python tools/bench_next_rows.py [--rows 16384] [--reps 20] > gpurun_out/next_rows.jsonl
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mulan_b200 import _lib, ops  # noqa: E402

D = 3072


def peak():
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    return float(json.load(open(p))['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs'
  return 6650.0, 'B200_PROFILING.md fallback'


def timed(fn, reps):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps * 1e-3


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--rows', type=int, default=16384)
  ap.add_argument('--reps', type=int, default=20)
  args = ap.parse_args()
  dev = torch.device('cuda:0')
  _lib.load()
  B, N = args.rows, args.rows * D
  pk, pk_src = peak()
  g = torch.Generator(device=dev).manual_seed(3)
  rn = lambda *s: torch.randn(s, generator=g, device=dev, dtype=torch.float32)
  a1, b1 = rn(1, D), rn(1, D)
  c1 = 1e-3 + torch.nn.functional.softplus(rn(1, D))
  a, b = rn(B, D), rn(B, D)
  c = 1e-3 + torch.nn.functional.softplus(rn(B, D))
  t = torch.rand(B, generator=g, device=dev) * 0.9 + 0.1
  s = t - 0.001
  z, net, eps = rn(B, D), rn(B, D), rn(B, D)
  out = torch.empty_like(z)
  desc = ops.Desc()
  rows = []

  def rec(name, nbytes, fn, note=''):
    sec = timed(fn, args.reps)
    gbs = nbytes / sec / 1e9
    rows.append({'kernel': name, 'us': sec * 1e6, 'algo_bytes': nbytes, 'gbs': gbs,
                 'frac_of_measured': gbs / pk, 'note': note})

  # ---- row 3: ancestral sampler (ldm/model_mulan_epsilon.py:377-457) ----
  tu, su = torch.full((B,), 0.4, device=dev), torch.full((B,), 0.399, device=dev)
  rec('mulan_sample_step (one coefficient row broadcast, one t: VDM.sample)', 16 * N,
      lambda: ops.sample_step(desc, a1, b1, c1, tu, su, z, net, eps, out=out),
      'z_t4 + net4 + eps4 -> z_s4; a,b,c L2-resident, factors cached per CTA (unconditional sampler)')
  rec('mulan_sample_step (one coefficient row broadcast, per-row t)', 16 * N,
      lambda: ops.sample_step(desc, a1, b1, c1, t, s, z, net, eps, out=out),
      'same traffic; factors recomputed per row: issue-bound')
  rec('mulan_sample_step (per-example coefficients)', 28 * N,
      lambda: ops.sample_step(desc, a, b, c, t, s, z, net, eps, out=out),
      '+ a,b,c 12 (conditional sampler)')
  rec('mulan_generate_x', 5 * N, lambda: ops.generate_x(desc, z), 'z_0 4 -> x 1')
  desc_pix = ops.Desc(gt_mode=1)
  rec('mulan_sample_gamma (per-pixel)', 16 * N, lambda: ops.sample_gamma(desc_pix, a, b, c, t),
      'a,b,c 12 -> g 4 (unet_type=ldm; includes the output allocation)')
  # ---- row 4: probability-flow ODE (ldm/model_mulan_epsilon.py:459-478) ----
  v = torch.sign(rn(B, D))
  rec('mulan_ode_drift (drift only)', 24 * N,
      lambda: ops.ode_drift(desc, a, b, c, t, z, net, None, False, out=out),
      'a,b,c 12 + x_t 4 + eps_hat 4 -> drift 4')
  nb, dd = torch.empty_like(z), torch.empty(B, device=dev)
  d = desc.c(B)
  h = _lib.load()
  P = lambda q: C.c_void_p(q.data_ptr())
  st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
  rec('mulan_ode_drift (+ Hutchinson pieces)', 32 * N,
      lambda: _lib.check(h.mulan_ode_drift(C.byref(d), B, P(a), P(b), P(c), P(t), P(z), P(net), P(v),
                                           0, P(out), P(nb), P(dd), st())),
      '+ v 4 -> net_bar 4 (+ div_direct[B])')
  dot = torch.empty(B, device=dev)
  rec('mulan_row_dot', 8 * N, lambda: ops.row_dot(z, v, add=dd, out=dot), 'u4 + v4 -> [B]')
  # RK45 state: y float64, 7 float32 stage rows
  n = N
  y = torch.randn(n, generator=g, device=dev, dtype=torch.float64)
  y_new = torch.empty_like(y)
  K = rn(7, n)
  y32 = torch.empty(n, device=dev)
  coef6 = (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84)
  rec('mulan_rk45_stage (6 stages -> y32, y_new)', (8 + 6 * 4 + 4 + 8) * n,
      lambda: ops.rk45_stage(6, coef6, 0.01, y, K, y_stage=y32, y_out=y_new),
      'y8 + 6 K rows 24 -> y_stage4 + y_out8')
  coef7 = (-71 / 57600, 0.0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40)
  scratch = torch.empty(_lib.MULAN_RK45_SCRATCH, dtype=torch.float64, device=dev)
  o1 = torch.empty(1, dtype=torch.float64, device=dev)
  rec('mulan_rk45_norm (7 stages)', (8 + 8 + 7 * 4) * n,
      lambda: ops.rk45_norm(7, coef7, 0.01, 1e-5, 1e-5, y, y_new, K, False, scratch, o1),
      'y8 + y_new8 + 7 K rows 28 -> scalar')
  del y, y_new, K, y32
  # ---- row 2: JAX-compatible draws ----
  rec('mulan_rng_normal', 4 * N, lambda: ops.rng_normal((1, 2), (N,), out=out.view(-1)),
      'write-only 4 B per draw; threefry2x32 + erfinv: instruction-bound')
  rec('mulan_rng_uniform', 4 * N, lambda: ops.rng_uniform((1, 2), (N,), out=out.view(-1)))
  rec('torch randn (Philox, context)', 4 * N, lambda: out.normal_())
  # ---- row 1: optimizer (ldm/train_state.py:70-102) ----
  npar = 71_150_000 // 4 * 4
  f = lambda k: torch.zeros(k, dtype=torch.float32, device=dev)
  params, mu, nu, ema = rn(npar), f(npar), f(npar), rn(npar)
  grads = rn(npar)
  sumsq = torch.zeros(1, device=dev)
  sc = torch.empty(_lib.MULAN_SUMSQ_SCRATCH, dtype=torch.float64, device=dev)
  step = [0]

  def adamw(clip):
    step[0] += 1
    dd_ = _lib.MulanAdamwDesc(npar, npar - 100_000, step[0], 0, 2e-4, 0.9, 0.99, 1e-8, 0.01,
                              0.9999, 1.0, clip, sumsq.data_ptr() if clip > 0 else None)
    _lib.check(h.mulan_adamw_ema(C.byref(dd_), P(params), P(grads), P(mu), P(nu), P(ema), st()))
  rec('mulan_adamw_ema (71.15 M parameters)', 36 * npar, lambda: adamw(0.0),
      'p,g,mu,nu,ema read 20 -> p,mu,nu,ema written 16')
  rec('mulan_grad_sumsq', 4 * npar,
      lambda: _lib.check(h.mulan_grad_sumsq(npar, P(grads), P(sc), P(sumsq), st())),
      'g 4 -> scalar')
  rec('mulan_adamw_ema with global-norm clip', 36 * npar, lambda: adamw(1.0))
  del params, mu, nu, ema, grads
  # ---- a11: auxiliary latent (tiny: [B,50]) ----
  L, Kk = 50, 15
  Ba = 131072
  logits = rn(Ba, L)
  draw = torch.rand(10, Ba, L, generator=g, device=dev) + 1e-3
  rec('mulan_aux_topk_fwd (131072 rows)', (L * 4 + 10 * L * 4 + L * 4 + 4) * Ba,
      lambda: ops.aux_topk_fwd(logits, draw, Kk), 'logits + 10 gamma draws -> embedding, kl_z')
  eb, kb = rn(Ba, L), rn(Ba)
  rec('mulan_aux_topk_bwd (131072 rows)', (L * 4 + 10 * L * 4 + L * 4 + 4 + L * 4) * Ba,
      lambda: ops.aux_topk_bwd(logits, draw, Kk, eb, kb))
  for r in rows:
    r.update(peak_gbs=pk, peak_source=pk_src, rows=B)
    print(json.dumps(r))


if __name__ == '__main__':
  main()
