#!/usr/bin/env python
"""NVLink reference numbers next to mulan_adamw_ema_peer (run under torchrun, N >= 2):
SM-driven peer READ and peer WRITE bandwidth for the CIFAR-10 gradient bucket (torch's copy kernel
on a mapped peer pointer), NCCL all-reduce, and the fused kernel in one call.
    torchrun --nproc-per-node 2 tools/peer_probe.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mulan_b200.optim import FlatTrainState  # noqa: E402
from mulan_b200.peer import device_view  # noqa: E402

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
local = int(os.environ.get('LOCAL_RANK', rank))
torch.cuda.set_device(local)
dev = torch.device(f'cuda:{local}')
dist.init_process_group('nccl', device_id=dev)
n = 71153852 // 4 * 4
p = torch.nn.Parameter(torch.zeros(n, device=dev))
st = FlatTrainState([('w', p)], comm='peer', bucket_mb=1e9, multicast=False)
peer = (rank + 1) % world
remote = device_view(st.peer.maps['grads'][peer], n, '<f4', dev)
localbuf = torch.empty(n, device=dev)


def timed(fn, reps=5):
  fn()
  dist.barrier(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps):
    fn()
  e1.record()
  torch.cuda.synchronize(); dist.barrier()
  t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
  return t.item()


out = {'world': world, 'bytes': 4 * n}
t = timed(lambda: localbuf.copy_(remote))
out['peer_read_ms'], out['peer_read_gbs'] = t, 4 * n / t / 1e6
t = timed(lambda: remote.copy_(localbuf))
out['peer_write_ms'], out['peer_write_gbs'] = t, 4 * n / t / 1e6
t = timed(lambda: dist.all_reduce(st.grads[:n]))
out['nccl_allreduce_ms'] = t
out['nccl_busbw_gbs'] = 2 * (world - 1) / world * 4 * n / t / 1e6
t = timed(lambda: (st._reset_ranges(), st.peer_update_range(0, st.n)))
out['fused_ms'] = t
# per rank and direction: peer gradient shards in + the peers' parameter shards in (out: mirror)
out['fused_nvlink_gbs_per_direction'] = 2 * (world - 1) / world * 4 * n / t / 1e6
out['timed_out'] = st.peer.timed_out()
st.peer.close()
st2 = FlatTrainState([('w', torch.nn.Parameter(torch.zeros(n, device=dev)))], comm='peer',
                     bucket_mb=1e9, multicast=True)
out['multicast'] = st2.peer.multicast
t = timed(lambda: (st2._reset_ranges(), st2.peer_update_range(0, st2.n)))
out['fused_multicast_ms'] = t
out['timed_out_mc'] = st2.peer.timed_out()
st2.peer.close()
if rank == 0:
  print(json.dumps(out), flush=True)
dist.destroy_process_group()
