#!/usr/bin/env python
"""Warm-up + one ELBO step (fwd_pre -> post value-and-grad + loss scalars -> bwd_pre) at `rows`
rows, for ncu captures of the small-batch launch shapes:
    ncu --set full --clock-control none --cache-control none -k regex:'fwd_pre|post|bwd_pre' \
        -s 9 -c 3 -o gpurun_out/ncu_step128 python tools/one_step.py 128"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_inputs, D  # noqa: E402
from mulan_b200 import ops  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device('cuda:0')
sm = make_inputs(rows, dev, 1)
gL = torch.full((rows,), 1.0 / (rows * D * math.log(2.0)), device=dev)
ws = ops.ElboWorkspace(ops.Desc(), rows, dev)
a = (sm['x'], sm['a'], sm['b'], sm['c'], sm['t'], sm['eps'], sm['net'])
for _ in range(4):
  ws.fwd_pre(sm['x'], sm['a'], sm['b'], sm['c'], sm['t'], sm['eps0'], sm['eps'])
  ws.post_bpd(*a, gL)
  ws.bwd_pre(*a, sm['z_bar'], sm['g_bar'], gL)
torch.cuda.synchronize()
