"""Flat train state + fused AdamW/EMA update (SURVEY.md 8f "next" row 1).

Mirrors TrainState / apply_gradients (ldm/train_state.py:56-119) with the optimizer of
ldm/experiment.py:132-182 (optax.adamw b1 .9, b2 .99, eps 1e-8, wd .01 masked to non-bias
parameters) and the schedule of ldm/experiment.py:106-129 (linear warm-up, no decay).

Every parameter, gradient, Adam moment and EMA copy lives in ONE contiguous float32 buffer
each (decayed parameters first), so that
  * pmean(grads) (ldm/experiment.py:341) is a single NCCL all-reduce over `grads` (+ the six
    loss scalars in its tail), and
  * the update is a single launch of mulan_adamw_ema (36 B per parameter) with the 1/world of
    the mean folded in.

Three ways of running the exchange (`comm=`):
  'allreduce'  one ncclAllReduce of the whole bucket after backward, then the full-size update
               on every rank (round 1).
  'overlap'    the bucket is cut into ranges; a post-accumulate-grad hook fires a range's
               ncclAllReduce (async, NCCL's stream) as soon as its last gradient has been
               written, so the exchange overlaps the rest of backward; then the full-size update.
  'peer'       B200-native: parameters and gradients live in peer memory; as each range
               completes, ONE kernel per rank (mulan_adamw_ema_peer, on a side stream) sums the
               range's gradient shards over NVLink -- inside the NVSwitch by multimem.ld_reduce
               when the buffers have a multicast mapping, by peer loads in rank order otherwise
               -- updates 1/world of it (AdamW + EMA on the local shard of the optimizer state)
               and stores the new parameters into every rank's buffer (multimem.st / peer
               stores): reduce-scatter, update and all-gather in one launch, overlapped with
               backward, no NCCL call on the gradient path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from .peer import PeerAllocations, device_view


def lr_schedule(step: int, learning_rate: float = 2e-4, num_steps_lr_warmup: int = 100,
                lr_decay: bool = False, num_steps_train: int = 0) -> float:
  """get_lr_schedule (ldm/experiment.py:106-129): optax.linear_schedule(0, lr, warmup), joined
  at `warmup` with a linear decay to 0 over num_steps_train - warmup steps when lr_decay.
  `step` is the 0-based count optax sees before the update."""
  step = max(step, 0)
  if lr_decay and step >= num_steps_lr_warmup:
    span = num_steps_train - num_steps_lr_warmup
    if span <= 0:
      return 0.0 if step > num_steps_lr_warmup else learning_rate
    frac = 1.0 - min(step - num_steps_lr_warmup, span) / span
    return learning_rate * frac
  if num_steps_lr_warmup <= 0:
    return learning_rate
  frac = min(step, num_steps_lr_warmup) / num_steps_lr_warmup
  return learning_rate * frac


def decay_mask(name: str) -> bool:
  """ldm/experiment.py:135-141: decay everything whose leaf is not a bias (the two layer-norm
  exceptions named there do not occur in these models; GroupNorm scales ARE decayed)."""
  return not name.endswith('bias')


def shard_range(lo: int, hi: int, world: int, rank: int) -> Tuple[int, int]:
  """The part of the flat range [lo, hi) that `rank` reduces, updates and broadcasts in
  mulan_adamw_ema_peer: whole float4 columns, ceil-divided, the last ranks take what is left
  (mirrors the kernel launcher, csrc/mulan_peer.cu)."""
  cols = (hi - lo) // 4
  per = (cols + world - 1) // world
  first = min(per * rank, cols)
  last = min(first + per, cols)
  return lo + 4 * first, lo + 4 * last


def plan_buckets(layout: List[Tuple[str, int, int]], n: int, bucket_elems: int):
  """Cut the flat buffer [0, n) into ranges of ~bucket_elems elements on parameter boundaries
  (multiples of 4 by construction).  Returns (ranges [(lo, hi)], members: range index -> names)."""
  ranges, members = [], []
  lo, names = 0, []
  for name, off, k in layout:
    names.append(name)
    end = off + (k + 3) // 4 * 4
    if end - lo >= bucket_elems:
      ranges.append((lo, end)); members.append(names)
      lo, names = end, []
  if lo < n or not ranges:
    ranges.append((lo, n)); members.append(names)
  return ranges, members


class PeerBuffers:
  """Parameter buffer, gradient bucket and flag block of THIS rank in peer memory, and this
  process's mappings of every other rank's.

  multicast=True: ONE symmetric allocation of torch.distributed._symmetric_memory (plumbing: it
  owns the CUDA-IPC exchange and binds the allocation to an NVSwitch multicast object) holding
  [flags | params | grads]; the multicast address lets mulan_adamw_ema_peer reduce in the switch
  (multimem.ld_reduce) and broadcast with one store (multimem.st).  multicast=False (or no
  multicast support): cudaMalloc + CUDA-IPC buffers of libmulan_b200 itself (mulan_b200/peer.py),
  plain peer loads / stores, bit-exact rank-order sums."""

  def __init__(self, n_params: int, n_grads: int, device, multicast: bool = True):
    world = dist.get_world_size()
    if world & (world - 1):
      raise ValueError(f'peer mode needs world in (1, 2, 4, 8), got {world}')
    self.world, self.rank = world, dist.get_rank()
    self.mc_grads = self.mc_params = None
    self.mem = self._symm = None
    nflag = _lib.MULAN_PEER_FLAG_WORDS
    if multicast:
      self._symm = self._try_symmetric(nflag + n_params + n_grads, device)
    if self._symm is not None:
      t, hdl = self._symm
      off_p, off_g = 4 * nflag, 4 * (nflag + n_params)
      self.maps = {'flags': [int(p) for p in hdl.buffer_ptrs],
                   'params': [int(p) + off_p for p in hdl.buffer_ptrs],
                   'grads': [int(p) + off_g for p in hdl.buffer_ptrs]}
      self.mc_params, self.mc_grads = int(hdl.multicast_ptr) + off_p, int(hdl.multicast_ptr) + off_g
      self.flags = t[:nflag].view(torch.int32)
      self.params, self.grads = t[nflag:nflag + n_params], t[nflag + n_params:]
    else:
      self.mem = PeerAllocations({'params': 4 * n_params, 'grads': 4 * n_grads,
                                  'flags': 4 * nflag}, device)
      self.maps = self.mem.maps
      self.params = device_view(self.mem.own['params'], n_params, '<f4', device)
      self.grads = device_view(self.mem.own['grads'], n_grads, '<f4', device)
      self.flags = device_view(self.mem.own['flags'], nflag, '<i4', device)
    self.epoch = 0

  def _try_symmetric(self, n_floats: int, device):
    """(tensor, handle) of a zeroed symmetric allocation with a multicast mapping, or None --
    decided collectively, so every rank takes the same path."""
    ok, res = 1, None
    try:
      import torch.distributed._symmetric_memory as symm_mem
      t = symm_mem.empty((n_floats + 3) // 4 * 4, dtype=torch.float32, device=device)
      t.zero_()
      hdl = symm_mem.rendezvous(t, group=dist.group.WORLD.group_name)
      if not int(getattr(hdl, 'multicast_ptr', 0) or 0):
        ok = 0
      res = (t, hdl)
    except Exception:
      ok = 0
    flag = torch.tensor([ok], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    torch.cuda.synchronize()
    if int(flag.item()) == 0:
      return None
    res[1].barrier()
    return res

  @property
  def multicast(self) -> bool:
    return self.mc_grads is not None

  def desc(self) -> '_lib.MulanPeerDesc':
    self.epoch += 1
    d = _lib.MulanPeerDesc()
    d.world, d.rank, d.epoch = self.world, self.rank, self.epoch
    for r in range(self.world):
      d.grads[r], d.params[r], d.flags[r] = (self.maps['grads'][r], self.maps['params'][r],
                                             self.maps['flags'][r])
    d.mc_grads, d.mc_params = self.mc_grads, self.mc_params
    return d

  def timed_out(self) -> bool:
    return bool(self.flags[_lib.MULAN_PEER_FLAG_ERR].item())

  def close(self):
    torch.cuda.synchronize()
    if self.mem is not None:
      self.mem.close()
    elif dist.is_initialized():
      dist.barrier()          # the symmetric allocation is released with its tensor
    self._symm = None


class FlatTrainState:
  """step, params, ema_params, opt_state of the reference's TrainState as flat buffers."""

  def __init__(self, named_params: Iterable[Tuple[str, torch.nn.Parameter]], extra: int = 8,
               b1: float = 0.9, b2: float = 0.99, eps: float = 1e-8, weight_decay: float = 0.01,
               learning_rate: float = 2e-4, num_steps_lr_warmup: int = 100,
               ema_rate: float = 0.9999, gradient_clip_norm: Optional[float] = None,
               lr_decay: bool = False, num_steps_train: int = 0, comm: str = 'allreduce',
               bucket_mb: float = 32.0, multicast='auto'):
    if comm not in ('allreduce', 'overlap', 'peer'):
      raise ValueError("comm must be 'allreduce', 'overlap' or 'peer'")
    multi = dist.is_initialized() and dist.get_world_size() > 1
    self.comm = comm if multi else 'allreduce'
    if self.comm == 'peer' and gradient_clip_norm:
      raise ValueError("comm='peer' cannot clip by the global norm (the reduced gradient never "
                       "exists in one place before the update); use 'overlap'")
    named = [(n, p) for n, p in named_params if p.requires_grad]
    if not named:
      raise ValueError('no trainable parameters')
    dev = named[0][1].device
    if dev.type != 'cuda':
      raise TypeError('FlatTrainState needs CUDA parameters (libmulan_b200 has no CPU path)')
    dec = [(n, p) for n, p in named if decay_mask(n)]
    nodec = [(n, p) for n, p in named if not decay_mask(n)]
    pad4 = lambda k: (k + 3) // 4 * 4
    self.layout: List[Tuple[str, int, int]] = []      # name, offset, numel
    off = 0
    for n, p in dec:
      self.layout.append((n, off, p.numel()))
      off += pad4(p.numel())
    self.n_decay = off
    for n, p in nodec:
      self.layout.append((n, off, p.numel()))
      off += pad4(p.numel())
    self.n = off
    self.extra = pad4(extra)
    f = lambda k: torch.zeros(k, dtype=torch.float32, device=dev)
    self.mu, self.nu = f(self.n), f(self.n)
    self.peer = None
    if self.comm == 'peer':
      # in-switch reduction pays from 8 ranks on (285 MB bucket: 0.72 vs 0.91 ms; 4 ranks: a tie;
      # 2 ranks: 0.85 vs 0.53 ms -- profiles/r2_variants.md section 5); below, peer loads in rank
      # order, which are also bit-reproducible against a single-process emulation
      if multicast == 'auto':
        multicast = dist.get_world_size() >= 8
      self.peer = PeerBuffers(self.n, self.n + self.extra, dev, multicast=bool(multicast))
      self.params, self.grads = self.peer.params, self.peer.grads
    else:
      self.params = f(self.n)
      self.grads = f(self.n + self.extra)        # tail: loss scalars ride in the all-reduce
    self.tail = self.grads[self.n:]
    by_name = dict(named)
    with torch.no_grad():
      for n, o, k in self.layout:
        p = by_name[n]
        self.params[o:o + k].copy_(p.detach().reshape(-1))
        p.data = self.params[o:o + k].view_as(p)             # parameters live in the flat buffer
        p.grad = self.grads[o:o + k].view_as(p)              # backward accumulates in place
    self.ema = self.params.clone()                            # ldm/train_state.py:108
    self.step = 0
    self.hp = dict(b1=b1, b2=b2, eps=eps, weight_decay=weight_decay)
    self.learning_rate, self.warmup, self.ema_rate = learning_rate, num_steps_lr_warmup, ema_rate
    self.lr_decay, self.num_steps_train = lr_decay, num_steps_train
    # optax.clip_by_global_norm in front of the chain when the config has gradient_clip_norm
    # (ldm/experiment.py:176-178)
    self.clip_norm = float(gradient_clip_norm) if gradient_clip_norm else 0.0
    if self.clip_norm > 0.0:
      self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
      self._scratch = torch.empty(_lib.MULAN_SUMSQ_SCRATCH, dtype=torch.float64, device=dev)
    # ---- ranges of the flat buffer exchanged as soon as their gradients are complete
    self.ranges, self._members = [(0, self.n)], [[n for n, _, _ in self.layout]]
    self._hooks = []
    if self.comm in ('overlap', 'peer'):
      self.ranges, self._members = plan_buckets(self.layout, self.n,
                                                int(bucket_mb * (1 << 20) / 4))
      self._range_of = {name: i for i, names in enumerate(self._members) for name in names}
      self._side = torch.cuda.Stream(device=dev)
      self._pending, self._fired, self._works = [], [], []
      for name, p in named:
        self._hooks.append(p.register_post_accumulate_grad_hook(
            lambda _p, _name=name: self._grad_ready(_name)))
      self._reset_ranges()

  # ---- overlap of the exchange with backward -------------------------------------------------
  def _reset_ranges(self):
    self._pending = [len(m) for m in self._members]
    self._fired = [False] * len(self.ranges)
    self._works = []
    self._step_desc = None

  def _grad_ready(self, name: str):
    i = self._range_of[name]
    self._pending[i] -= 1
    if self._pending[i] == 0 and not self._fired[i]:
      self._fire(i)

  def _adamw_desc(self, grad_scale: float):
    if self._step_desc is None:                    # one (step, lr) for every range of a step
      lr = lr_schedule(self.step, self.learning_rate, self.warmup, self.lr_decay,
                       self.num_steps_train)
      self.step += 1
      self._lr = lr
      self._step_desc = _lib.MulanAdamwDesc(
          self.n, self.n_decay, self.step, 0, lr, self.hp['b1'], self.hp['b2'], self.hp['eps'],
          self.hp['weight_decay'], self.ema_rate, grad_scale, 0.0, None)
    return self._step_desc

  def _fire(self, i: int):
    """Range i's gradients are final on this rank: start its exchange."""
    self._fired[i] = True
    lo, hi = self.ranges[i]
    if self.comm == 'overlap':
      self._works.append(dist.all_reduce(self.grads[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
      return
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream())
    self._side.wait_event(ev)
    self.peer_update_range(lo, hi, self._side)

  def peer_update_range(self, lo: int, hi: int, stream=None):
    """ONE mulan_adamw_ema_peer call over the flat range [lo, hi) (every rank must make the same
    call): reduce-scatter by peer loads, AdamW+EMA on this rank's shard, all-gather by peer
    stores.  The first call of a step fixes the step count and learning rate."""
    ptr = lambda t: C.c_void_p(t.data_ptr())
    stream = stream or torch.cuda.current_stream()
    d = self._adamw_desc(1.0 / self.peer.world)
    pd = self.peer.desc()
    _lib.check(_lib.load().mulan_adamw_ema_peer(
        C.byref(d), C.byref(pd), lo, hi, ptr(self.mu), ptr(self.nu), ptr(self.ema),
        C.c_void_p(stream.cuda_stream)))

  def finish_exchange(self):
    """After backward: fire the ranges no hook completed (unused parameters), in index order --
    the same on every rank -- and make the current stream wait for everything in flight."""
    if self.comm == 'allreduce':
      return
    for i in range(len(self.ranges)):
      if not self._fired[i]:
        self._fire(i)
    if self.comm == 'overlap':
      for w in self._works:
        w.wait()
    else:
      ev = torch.cuda.Event()
      ev.record(self._side)
      torch.cuda.current_stream().wait_event(ev)

  def grad_global_norm(self) -> torch.Tensor:
    """optax.global_norm of the (all-reduced, not yet averaged) gradient bucket: device scalar."""
    ptr = lambda t: C.c_void_p(t.data_ptr())
    if self.clip_norm <= 0.0:
      raise RuntimeError('grad_global_norm needs gradient_clip_norm to be configured')
    _lib.check(_lib.load().mulan_grad_sumsq(
        self.n, ptr(self.grads), ptr(self._scratch), ptr(self._sumsq),
        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return self._sumsq.sqrt()

  def zero_grad(self):
    self.grads.zero_()
    if self.comm != 'allreduce':
      self._reset_ranges()

  def all_reduce(self):
    """pmean(grads) + pmean(scalars).  'allreduce': ONE all-reduce (sum) of bucket + tail; the
    1/world is applied by the update kernel (gradients) / here (the few tail scalars).
    'overlap' / 'peer': the gradient ranges are already in flight (or done); only the tail's
    handful of scalars is reduced here."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
      return
    if self.comm == 'allreduce':
      dist.all_reduce(self.grads, op=dist.ReduceOp.SUM)
    else:
      self.finish_exchange()
      dist.all_reduce(self.tail, op=dist.ReduceOp.SUM)
    self.tail.div_(dist.get_world_size())

  def apply_gradients(self, grad_scale: Optional[float] = None):
    """TrainState.apply_gradients (ldm/train_state.py:70-102): one fused launch.  In 'peer' mode
    the update already ran range by range inside mulan_adamw_ema_peer."""
    if self.comm == 'peer':
      self.finish_exchange()
      if self._step_desc is None:
        raise RuntimeError('apply_gradients: no gradient range was exchanged this step')
      return self._lr
    world = dist.get_world_size() if dist.is_initialized() else 1
    if grad_scale is None:
      grad_scale = 1.0 / world
    lr = lr_schedule(self.step, self.learning_rate, self.warmup, self.lr_decay,
                     self.num_steps_train)
    self.step += 1
    ptr = lambda t: C.c_void_p(t.data_ptr())
    sumsq = None
    if self.clip_norm > 0.0:
      self.grad_global_norm()
      sumsq = self._sumsq.data_ptr()
    d = _lib.MulanAdamwDesc(self.n, self.n_decay, self.step, 0, lr, self.hp['b1'], self.hp['b2'],
                            self.hp['eps'], self.hp['weight_decay'], self.ema_rate, grad_scale,
                            self.clip_norm, sumsq)
    _lib.check(_lib.load().mulan_adamw_ema(
        C.byref(d), ptr(self.params), ptr(self.grads), ptr(self.mu), ptr(self.nu), ptr(self.ema),
        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return lr

  def gather_sharded_state(self):
    """'peer' mode keeps mu / nu / ema up to date only on the shards this rank owns.  Before a
    checkpoint or an evaluation with ema_params (ldm/experiment.py:292-303) every rank collects
    the other shards from their owners (one broadcast per range, shard and tensor: a rare path)."""
    if self.comm != 'peer':
      return
    world = dist.get_world_size()
    for lo, hi in self.ranges:
      for r in range(world):
        a, b = shard_range(lo, hi, world, r)
        if b > a:
          for t in (self.mu, self.nu, self.ema):
            dist.broadcast(t[a:b], src=r)

  def ema_state_dict(self) -> Dict[str, torch.Tensor]:
    self.gather_sharded_state()
    return {n: self.ema[o:o + k] for n, o, k in self.layout}


def train_step(model, state: FlatTrainState, batch: dict, generator=None, draws=None):
  """Experiment.train_step (ldm/experiment.py:335-356) on the flat state: value_and_grad of
  loss_fn, ONE all-reduce (gradients + the six scalars), ONE fused AdamW+EMA launch."""
  from .model import loss_fn
  state.zero_grad()
  bpd, metrics = loss_fn(model, batch, step=state.step, is_train=True, draws=draws,
                         generator=generator)
  bpd.backward()
  keys = sorted(metrics['scalars'])
  with torch.no_grad():
    state.tail[:len(keys)] = torch.stack([metrics['scalars'][k].detach().float().reshape(())
                                          for k in keys])
  state.all_reduce()
  state.apply_gradients()
  out = state.tail[:len(keys)].clone()       # the bucket is zeroed by the next step
  return {k: out[i] for i, k in enumerate(keys)}
