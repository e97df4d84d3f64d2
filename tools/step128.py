#!/usr/bin/env python
"""A few eager 128-row ELBO steps (for an ncu launch list of the small-batch kernels)."""
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
for _ in range(6):
  ws.fwd_pre(sm['x'], sm['a'], sm['b'], sm['c'], sm['t'], sm['eps0'], sm['eps'])
  ws.post_bpd(sm['x'], sm['a'], sm['b'], sm['c'], sm['t'], sm['eps'], sm['net'], gL)
  ws.bwd_pre(sm['x'], sm['a'], sm['b'], sm['c'], sm['t'], sm['eps'], sm['net'], sm['z_bar'],
             sm['g_bar'], gL)
torch.cuda.synchronize()
