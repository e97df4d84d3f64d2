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
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib


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


class FlatTrainState:
  """step, params, ema_params, opt_state of the reference's TrainState as flat buffers."""

  def __init__(self, named_params: Iterable[Tuple[str, torch.nn.Parameter]], extra: int = 8,
               b1: float = 0.9, b2: float = 0.99, eps: float = 1e-8, weight_decay: float = 0.01,
               learning_rate: float = 2e-4, num_steps_lr_warmup: int = 100,
               ema_rate: float = 0.9999, gradient_clip_norm: Optional[float] = None,
               lr_decay: bool = False, num_steps_train: int = 0):
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
    self.params, self.mu, self.nu = f(self.n), f(self.n), f(self.n)
    self.grads = f(self.n + self.extra)          # tail: loss scalars ride in the all-reduce
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

  def all_reduce(self):
    """pmean(grads) + pmean(scalars): ONE all-reduce (sum); the 1/world is applied by the
    update kernel (gradients) / here (the few tail scalars)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
      dist.all_reduce(self.grads, op=dist.ReduceOp.SUM)
      self.tail.div_(dist.get_world_size())

  def apply_gradients(self, grad_scale: Optional[float] = None):
    """TrainState.apply_gradients (ldm/train_state.py:70-102): one fused launch."""
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

  def ema_state_dict(self) -> Dict[str, torch.Tensor]:
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
