#!/usr/bin/env python
"""One warm-up + two launches of mulan_fwd_pre at `rows` rows (for ncu captures of a kernel shape):
    MULAN_FWD_PRE_V=5 python tools/one_fwd_pre.py [rows] [eps|vel]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_inputs  # noqa: E402
from mulan_b200 import ops  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
model = sys.argv[2] if len(sys.argv) > 2 else 'eps'
dev = torch.device('cuda:0')
inp = make_inputs(rows, dev, 1)
ws = ops.ElboWorkspace(ops.Desc(param=0 if model == 'eps' else 1), rows, dev,
                       save_w=(model == 'eps'))
for _ in range(3):
  ws.fwd_pre(inp['x'], inp['a'], inp['b'], inp['c'], inp['t'], inp['eps0'], inp['eps'])
torch.cuda.synchronize()
