#!/usr/bin/env python
"""A/B timings behind this round's kernel decisions (one B200; prints JSON lines):
  * fwd_pre kernel shapes (MULAN_FWD_PRE_V) for the epsilon form (w saved, 29 B/sub-pixel) and the
    plain velocity model (25 B)
  * loss-scalar reduction: single CTA vs parallel vs fused into the post kernel
  * the 128-row step: plain launches vs programmatic dependent launch vs fused reduction
  * c_raw (in-kernel 1e-3 + softplus) vs a separate torch softplus pass
  * broadcast noise rows (dense VLB)
    python tools/variants_r2.py [rows]
"""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_inputs, load_peak, D  # noqa: E402
from mulan_b200 import ops  # noqa: E402

dev = torch.device('cuda:0')
PEAK, _ = load_peak()


def timeit(fn, iters=30):
  st = torch.cuda.current_stream()
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(st)
  for _ in range(iters):
    fn()
  e1.record(st)
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / iters


def graph_time(fn, replays=200):
  st = torch.cuda.Stream()
  with torch.cuda.stream(st):
    for _ in range(3):
      fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
      fn()
    for _ in range(5):
      g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(replays):
      g.replay()
    e1.record(st)
    torch.cuda.synchronize()
  return e0.elapsed_time(e1) * 1000 / replays      # us


def emit(**kw):
  print(json.dumps(kw), flush=True)


def main():
  rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
  inp = make_inputs(rows, dev, 1)
  nsub = rows * D
  gL = torch.full((rows,), 1.0 / (rows * D * math.log(2.0)), device=dev)

  # ---- fwd_pre shapes
  for name, param, save_w, nbytes in (('eps', 0, True, 29), ('vel', 1, False, 25)):
    ws = ops.ElboWorkspace(ops.Desc(param=param), rows, dev, save_w=save_w)
    f = lambda: ws.fwd_pre(inp['x'], inp['a'], inp['b'], inp['c'], inp['t'], inp['eps0'],
                           inp['eps'])
    for v in [None, 0, 1, 5, 12]:
      if v is None:
        os.environ.pop('MULAN_FWD_PRE_V', None)
      else:
        os.environ['MULAN_FWD_PRE_V'] = str(v)
      ms = timeit(f)
      gbs = nbytes * nsub / (ms * 1e-3) / 1e9
      emit(what='fwd_pre_shape', model=name, v=v, ms=ms, gbs=gbs, frac=gbs / PEAK)
    os.environ.pop('MULAN_FWD_PRE_V', None)
    del ws
  if os.environ.get('VARIANTS_ONLY_SHAPES'):
    return

  # ---- reduction forms (epsilon, value-and-grad)
  ws = ops.ElboWorkspace(ops.Desc(), rows, dev)
  a = (inp['x'], inp['a'], inp['b'], inp['c'], inp['t'], inp['eps'], inp['net'])
  ws.fwd_pre(inp['x'], inp['a'], inp['b'], inp['c'], inp['t'], inp['eps0'], inp['eps'])
  emit(what='post_vg', ms=timeit(lambda: ws.fwd_bwd_post(*a, gL)))
  emit(what='post_vg_fused_reduce', ms=timeit(lambda: ws.post_bpd(*a, gL)))
  emit(what='bpd_reduce_single', ms=timeit(lambda: ws.bpd_reduce(None, parallel=False)))
  emit(what='bpd_reduce_parallel', ms=timeit(lambda: ws.bpd_reduce(None, parallel=True)))
  emit(what='bwd_pre', ms=timeit(lambda: ws.bwd_pre(*a, inp['z_bar'], inp['g_bar'], gL)))

  # ---- c_raw vs a separate softplus pass
  c_raw = torch.randn((rows, D), device=dev)
  wr = ops.ElboWorkspace(ops.Desc(c_raw=True), rows, dev)
  emit(what='fwd_pre_c_raw', ms=timeit(lambda: wr.fwd_pre(inp['x'], inp['a'], inp['b'], c_raw,
                                                           inp['t'], inp['eps0'], inp['eps'])))
  emit(what='torch_softplus_fwd', ms=timeit(lambda: 1e-3 + torch.nn.functional.softplus(c_raw)))
  ar = (inp['x'], inp['a'], inp['b'], c_raw, inp['t'], inp['eps'], inp['net'])
  emit(what='bwd_pre_c_raw', ms=timeit(lambda: wr.bwd_pre(*ar, inp['z_bar'], inp['g_bar'], gL)))
  cb = torch.randn((rows, D), device=dev)
  emit(what='torch_softplus_bwd', ms=timeit(lambda: cb * torch.sigmoid(c_raw)))
  del wr, c_raw, cb

  # ---- in-kernel draws (f2): mulan_fwd_pre_keyed vs two stand-alone draws + mulan_fwd_pre
  import ctypes as C
  from mulan_b200 import _lib
  lib = _lib.load()
  wk = ops.ElboWorkspace(ops.Desc(), rows, dev)
  eps_out = torch.empty((rows, D), device=dev)
  k0, k1 = (C.c_uint32 * 2)(1, 2), (C.c_uint32 * 2)(3, 4)
  pp = lambda t_: C.c_void_p(t_.data_ptr()) if t_ is not None else None
  def keyed(w_save, eps_o):
    _lib.check(lib.mulan_fwd_pre_keyed(
        C.byref(wk._d), k0, k1, pp(inp['x']), pp(inp['a']), pp(inp['b']), pp(inp['c']), pp(inp['t']),
        pp(wk.z_t), pp(wk.g_net), pp(wk.w if w_save else None), None, pp(eps_o),
        pp(wk.loss_recon), pp(wk.loss_klz_prior), pp(wk.var_sums),
        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
  e0b, e1b = torch.empty((rows, D), device=dev), torch.empty((rows, D), device=dev)
  def separate():
    ops.rng_normal((1, 2), (rows, D), device=dev, out=e0b)
    ops.rng_normal((3, 4), (rows, D), device=dev, out=e1b)
    wk.fwd_pre(inp['x'], inp['a'], inp['b'], inp['c'], inp['t'], e0b, e1b)
  emit(what='fwd_pre_keyed_w_and_eps_out', ms=timeit(lambda: keyed(True, eps_out)))
  emit(what='fwd_pre_keyed_no_copies', ms=timeit(lambda: keyed(False, None)))
  emit(what='two_rng_normal_plus_fwd_pre', ms=timeit(separate))
  emit(what='fwd_pre_reading_existing_draws',
       ms=timeit(lambda: wk.fwd_pre(inp['x'], inp['a'], inp['b'], inp['c'], inp['t'], e0b, e1b)))
  g_ = torch.Generator(device=dev).manual_seed(0)
  emit(what='two_torch_philox_randn_plus_fwd_pre', ms=timeit(lambda: (
      e0b.normal_(generator=g_), e1b.normal_(generator=g_),
      wk.fwd_pre(inp['x'], inp['a'], inp['b'], inp['c'], inp['t'], e0b, e1b))))
  del wk, eps_out, e0b, e1b

  # ---- full steps at `rows` (graph): separate reduce vs fused, plain vs PDL
  def make_step(desc, fused, rws):
    def step():
      rws.fwd_pre(inp['x'][:rws.rows], inp['a'][:rws.rows], inp['b'][:rws.rows],
                  inp['c'][:rws.rows], inp['t'][:rws.rows], inp['eps0'][:rws.rows],
                  inp['eps'][:rws.rows])
      aa = tuple(v[:rws.rows] for v in a)
      if fused:
        rws.post_bpd(*aa, gL[:rws.rows])
      else:
        rws.fwd_bwd_post(*aa, gL[:rws.rows])
        rws.bpd_reduce(None)
      rws.bwd_pre(*aa, inp['z_bar'][:rws.rows], inp['g_bar'][:rws.rows], gL[:rws.rows])
    return step
  for r in (128, 256, 2048, rows):
    for pdl in (False, True):
      for fused in (False, True):
        desc = ops.Desc(pdl=pdl)
        rws = ops.ElboWorkspace(desc, r, dev)
        us = graph_time(make_step(desc, fused, rws), replays=200 if r <= 2048 else 30)
        emit(what='step_graph', rows=r, pdl=pdl, fused_reduce=fused, us=us,
             samples_per_s=r / (us * 1e-6))

  # ---- dense VLB forward: tiled noise vs broadcast noise rows, per launch size
  for lrows in (2048, 4736, 16384):
    if lrows > rows:
      continue
    for bc in (False, True):
      desc = ops.Desc(noise_rows=128 if bc else 0)
      rws = ops.ElboWorkspace(desc, lrows, dev)
      e0 = inp['eps0'][:128].contiguous() if bc else inp['eps0'][:lrows]
      e = inp['eps'][:128].contiguous() if bc else inp['eps'][:lrows]
      def fwd():
        rws.fwd_pre(inp['x'][:lrows], inp['a'][:lrows], inp['b'][:lrows], inp['c'][:lrows],
                    inp['t'][:lrows], e0, e)
        rws.post_bpd(inp['x'][:lrows], inp['a'][:lrows], inp['b'][:lrows], inp['c'][:lrows],
                     inp['t'][:lrows], e, inp['net'][:lrows], None)
      us = graph_time(fwd, replays=100)
      nb = (21 + 8 * (128 / lrows) + 8 + 4 * (128 / lrows)) if bc else 41
      emit(what='dense_vlb_launch', rows=lrows, broadcast_noise=bc, us=us,
           rows_per_s=lrows / (us * 1e-6), algo_bytes_per_subpixel=nb,
           frac=nb * lrows * D / (us * 1e-6) / 1e9 / PEAK)


if __name__ == '__main__':
  main()
