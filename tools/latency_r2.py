#!/usr/bin/env python
"""Device time of one small-batch ELBO step (fwd_pre -> post_bpd -> bwd_pre), measured from a
CUDA graph holding 20 consecutive steps (a one-step graph replayed back to back measures the
host's graph-launch rate, not the GPU): rows x {plain, PDL} x fwd_pre shape, and each kernel alone.
    python tools/latency_r2.py > gpurun_out/latency_r2.jsonl"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_inputs, D  # noqa: E402
from mulan_b200 import ops  # noqa: E402

dev = torch.device('cuda:0')
inp = make_inputs(4096, dev, 1)


def graph_us(fn, per_graph=20, replays=20):
  st = torch.cuda.Stream()
  with torch.cuda.stream(st):
    for _ in range(3):
      fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
      for _ in range(per_graph):
        fn()
    for _ in range(3):
      g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(replays):
      g.replay()
    e1.record(st)
    torch.cuda.synchronize()
  return e0.elapsed_time(e1) * 1000 / (replays * per_graph)


# The floor one-CTA-per-row launches can reach: the same three kernels over ONE row (one SM busy,
# everything else idle) and a chain of near-empty launches (mulan_scale_rows over 1 x 4 floats).
if os.environ.get('LATENCY_FLOOR', '1') == '1':
  one = {k: v[:1].contiguous() for k, v in inp.items()}
  tiny = torch.zeros(1, 4, device=dev)
  tiny_s = torch.ones(1, device=dev)
  null = lambda: ops.scale_rows(tiny, tiny_s, tiny_s)
  print(json.dumps(dict(floor='null_launch_chain', per_launch_us=graph_us(null, per_graph=60))),
        flush=True)
  for rows in (1, 8, 32, 64, 96, 128, 148):
    sm = {k: v[:rows].contiguous() for k, v in inp.items()}
    gL = torch.full((rows,), 1.0 / (rows * D * math.log(2.0)), device=dev)
    for pdl in (False, True):
      ws = ops.ElboWorkspace(ops.Desc(pdl=pdl), rows, dev)
      a = (sm['x'], sm['a'], sm['b'], sm['c'], sm['t'], sm['eps'], sm['net'])

      def step():
        ws.fwd_pre(sm['x'], sm['a'], sm['b'], sm['c'], sm['t'], sm['eps0'], sm['eps'])
        ws.post_bpd(*a, gL)
        ws.bwd_pre(*a, sm['z_bar'], sm['g_bar'], gL)
      print(json.dumps(dict(floor='elbo_step', rows=rows, pdl=pdl, step_us=graph_us(step))),
            flush=True)
  if os.environ.get('LATENCY_FLOOR_ONLY') == '1':
    sys.exit(0)

for rows in (128, 256, 512, 1024, 2048, 4096):
  sm = {k: v[:rows].contiguous() for k, v in inp.items()}
  gL = torch.full((rows,), 1.0 / (rows * D * math.log(2.0)), device=dev)
  for shape in (None, '0', '12'):
    if shape is None:
      os.environ.pop('MULAN_FWD_PRE_V', None)
    else:
      os.environ['MULAN_FWD_PRE_V'] = shape
    for pdl in (False, True):
      ws = ops.ElboWorkspace(ops.Desc(pdl=pdl), rows, dev)
      a = (sm['x'], sm['a'], sm['b'], sm['c'], sm['t'], sm['eps'], sm['net'])
      k_pre = lambda: ws.fwd_pre(sm['x'], sm['a'], sm['b'], sm['c'], sm['t'], sm['eps0'], sm['eps'])
      k_post = lambda: ws.post_bpd(*a, gL)
      k_bwd = lambda: ws.bwd_pre(*a, sm['z_bar'], sm['g_bar'], gL)

      def step():
        k_pre(); k_post(); k_bwd()
      rec = dict(rows=rows, fwd_pre_shape=shape, pdl=pdl, step_us=graph_us(step))
      if not pdl:
        rec.update(fwd_pre_us=graph_us(k_pre), post_bpd_us=graph_us(k_post),
                   bwd_pre_us=graph_us(k_bwd))
        k_sep = lambda: (ws.fwd_bwd_post(*a, gL), ws.bpd_reduce(None))
        rec['post_plus_reduce_us'] = graph_us(k_sep)
      print(json.dumps(rec), flush=True)
os.environ.pop('MULAN_FWD_PRE_V', None)
