"""CUDA-IPC peer memory for one process per GPU on one node (NVLink 5 / NVSwitch): buffers
allocated by libmulan_b200 (mulan_peer_alloc) that every other rank maps (mulan_peer_open), so
that kernels can load from and store to any peer.  Users: the gradient exchange fused with the
optimizer (optim.PeerBuffers -> mulan_adamw_ema_peer) and the loss scalars' pmean without a
collective (ScalarBoard -> mulan_post_bpd_peer / mulan_scalar_board_read).

torch.distributed is plumbing here: it carries the 64-byte IPC handles once, at construction.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import torch
import torch.distributed as dist

from . import _lib


class _DevArray:
  """A raw device allocation exposed through __cuda_array_interface__ (zero-copy torch view)."""

  def __init__(self, ptr: int, n: int, typestr: str):
    self.__cuda_array_interface__ = {'shape': (n,), 'typestr': typestr, 'data': (ptr, False),
                                     'version': 2}


def device_view(ptr: int, n: int, typestr: str, device) -> torch.Tensor:
  return torch.as_tensor(_DevArray(ptr, n, typestr), device=device)


class PeerAllocations:
  """A named set of zero-initialised buffers of THIS rank in peer memory plus this process's
  mappings of every other rank's: own[name] (device address), maps[name][rank]."""

  def __init__(self, sizes: Dict[str, int], device):
    lib = _lib.load()
    self.world, self.rank = dist.get_world_size(), dist.get_rank()
    if self.world > _lib.MULAN_PEER_MAX:
      raise ValueError(f'peer memory serves up to {_lib.MULAN_PEER_MAX} ranks, got {self.world}')
    torch.cuda.set_device(device)
    self.device = device
    self.own: Dict[str, int] = {}
    self.maps: Dict[str, List[int]] = {k: [] for k in sizes}
    self._opened: List[int] = []
    # Every step below is collective-safe: a rank that fails still takes part in the exchanges,
    # and EVERY rank raises when any rank failed (nobody is left waiting in a barrier).
    handles, err = {}, None
    try:
      for k, nbytes in sizes.items():
        ptr, h = C.c_void_p(), (C.c_char * _lib.MULAN_PEER_HANDLE_BYTES)()
        _lib.check(lib.mulan_peer_alloc(int(nbytes), C.byref(ptr), h))
        self.own[k], handles[k] = ptr.value, bytes(h)
    except Exception as exc:            # reported collectively below
      err = f'rank {self.rank}: {exc}'
    everyone = [None] * self.world
    dist.all_gather_object(everyone, (handles, err))
    self._raise_if_any([e for _, e in everyone])
    err = None
    try:
      for r in range(self.world):
        for k in sizes:
          if r == self.rank:
            self.maps[k].append(self.own[k])
            continue
          ptr = C.c_void_p()
          _lib.check(lib.mulan_peer_open(everyone[r][0][k], C.byref(ptr)))
          self.maps[k].append(ptr.value)
          self._opened.append(ptr.value)
    except Exception as exc:
      err = f'rank {self.rank}: {exc}'
    errs = [None] * self.world
    dist.all_gather_object(errs, err)    # doubles as the "everyone has mapped everything" barrier
    self._raise_if_any(errs)

  def _raise_if_any(self, errs):
    bad = [e for e in errs if e]
    if bad:
      self.close(collective=False)
      raise RuntimeError('peer memory is not available on this node: ' + '; '.join(bad))

  def close(self, collective: bool = True):
    lib = _lib.load()
    torch.cuda.synchronize()
    if collective and dist.is_initialized():
      dist.barrier()
    for ptr in self._opened:
      lib.mulan_peer_close(C.c_void_p(ptr))
    self._opened = []
    if collective and dist.is_initialized():
      dist.barrier()        # nobody frees while a peer still maps
    for ptr in self.own.values():
      lib.mulan_peer_free(C.c_void_p(ptr))
    self.own = {}


class ScalarBoard:
  """pmean of the six loss_fn scalars (ldm/experiment.py:347-348) without a collective: the post
  kernel's finalising thread stores this rank's scalars into every rank's board
  (mulan_post_bpd_peer); `read()` averages the latest step on demand, identically on every rank."""

  def __init__(self, device):
    lib = _lib.load()
    self.mem = PeerAllocations({'board': int(lib.mulan_scalar_board_bytes())}, device)
    self.c = _lib.MulanScalarBoard()
    self.c.world, self.c.rank = self.mem.world, self.mem.rank
    for r in range(self.mem.world):
      self.c.boards[r] = self.mem.maps['board'][r]
    self._mean = torch.empty(6, dtype=torch.float32, device=device)
    self._epoch = torch.zeros(1, dtype=torch.int32, device=device)

  def byref(self):
    return C.byref(self.c)

  def read(self):
    """-> (mean[6] device tensor, step device tensor) of the latest step this rank published
    (NaNs / step 0 if a peer had already lapped the 64-slot ring or the wait timed out)."""
    _lib.check(_lib.load().mulan_scalar_board_read(
        C.byref(self.c), C.c_void_p(self._mean.data_ptr()), C.c_void_p(self._epoch.data_ptr()),
        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return self._mean, self._epoch

  def close(self):
    self.mem.close()
