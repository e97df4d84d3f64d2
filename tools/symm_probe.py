#!/usr/bin/env python
"""Does torch's symmetric memory (incl. the NVSwitch multicast mapping) work on this box?
    torchrun --nproc-per-node 2 tools/symm_probe.py"""
import os
import torch
import torch.distributed as dist

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
local = int(os.environ.get('LOCAL_RANK', rank))
torch.cuda.set_device(local)
dev = torch.device(f'cuda:{local}')
dist.init_process_group('nccl', device_id=dev)
try:
  import torch.distributed._symmetric_memory as symm_mem
  t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
  t.fill_(float(rank + 1))
  hdl = symm_mem.rendezvous(t, group=dist.group.WORLD.group_name)
  info = dict(rank=rank, world=hdl.world_size, buffer_ptrs=[hex(p) for p in hdl.buffer_ptrs],
              multicast_ptr=hex(getattr(hdl, 'multicast_ptr', 0) or 0),
              signal_pads=len(hdl.signal_pad_ptrs))
  hdl.barrier()
  peer = hdl.get_buffer((rank + 1) % world, (8,), torch.float32)
  info['peer_value'] = peer[0].item()
  try:
    out = torch.ops.symm_mem.multimem_all_reduce_(t, 'sum', dist.group.WORLD.group_name)
    torch.cuda.synchronize()
    info['multimem_all_reduce'] = t[0].item()
  except Exception as exc:
    info['multimem_all_reduce'] = 'failed: ' + repr(exc)[:200]
  print(info, flush=True)
except Exception as exc:
  print(dict(rank=rank, error=repr(exc)[:400]), flush=True)
dist.destroy_process_group()
