#!/usr/bin/env python
"""A/B of the in-switch (multicast) form of mulan_adamw_ema_peer on the CIFAR-10 bucket: CTAs per
SM x {plain, next-chunk prefetch}; each variant is first checked bit for bit against the default
variant from the same state.  Run under torchrun (N >= 2); one JSON line per variant.
    torchrun --nproc-per-node 8 tools/peer_mc_sweep.py > gpurun_out/peer_mc_sweep.jsonl"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mulan_b200.optim import FlatTrainState  # noqa: E402

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
local = int(os.environ.get('LOCAL_RANK', rank))
torch.cuda.set_device(local)
dev = torch.device(f'cuda:{local}')
dist.init_process_group('nccl', device_id=dev)
n = 71153852 // 4 * 4
st = FlatTrainState([('w', torch.nn.Parameter(torch.zeros(n, device=dev)))], comm='peer',
                    bucket_mb=1e9, multicast=True)
if not st.peer.multicast:
  if rank == 0:
    print(json.dumps({'multicast': False}))
  sys.exit(0)
g = torch.Generator(device=dev); g.manual_seed(100 + rank)
grads0 = torch.randn(n, device=dev, generator=g) * 1e-2
g.manual_seed(7)
params0 = torch.randn(n, device=dev, generator=g)


def reset():
  st.params.copy_(params0); st.ema.copy_(params0)
  st.mu.zero_(); st.nu.zero_()
  st.grads[:n].copy_(grads0)
  st.step = 0
  torch.cuda.synchronize(); dist.barrier()


def one_call():
  st._reset_ranges()
  st.peer_update_range(0, st.n)


def timed(reps=5):
  one_call()
  dist.barrier(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps):
    one_call()
  e1.record()
  torch.cuda.synchronize(); dist.barrier()
  t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
  return t.item()


def setenv(per_sm, pf):
  os.environ['MULAN_PEER_CTAS_PER_SM'] = str(per_sm)
  os.environ['MULAN_PEER_MC_PREFETCH'] = str(pf)


setenv(4, 0)
reset(); one_call(); torch.cuda.synchronize(); dist.barrier()
want = (st.params.clone(), st.ema.clone(), st.mu.clone(), st.nu.clone())
for per_sm, pf in ((4, 0), (4, 1), (8, 1), (8, 0), (2, 1), (6, 1), (2, 0), (6, 0), (3, 1)):
  setenv(per_sm, pf)
  reset(); one_call(); torch.cuda.synchronize(); dist.barrier()
  same = all(torch.equal(a, b) for a, b in zip(want, (st.params, st.ema, st.mu, st.nu)))
  flag = torch.tensor([int(same)], device=dev)
  dist.all_reduce(flag, op=dist.ReduceOp.MIN)
  ms = timed()
  if rank == 0:
    print(json.dumps({'world': world, 'ctas_per_sm': per_sm, 'prefetch': pf, 'ms': ms,
                      'bit_identical_to_default': bool(flag.item()),
                      'timed_out': st.peer.timed_out()}), flush=True)
st.peer.close()
dist.destroy_process_group()
