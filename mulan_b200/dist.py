"""One-process-per-GPU data parallelism around the hot path (torch.distributed; NCCL over
NVLink on the GPU box, gloo in CPU tests).

The reference is single-process `jax.pmap` over the local devices with axis name 'batch'
(ldm/experiment.py:89-95).  Rows are independent, so the path shards by example with no
data-path collective; the exchanges that FOLLOW it are:
  * pmean of the gradient pytree        ldm/experiment.py:341   -> ONE flat-bucket all-reduce
  * pmean of the six loss scalars       ldm/experiment.py:347-348, 365-366
  * dense-VLB evaluation: images are sharded, one final (sum, count) reduction
                                        ldm/notebook_utils.py:176-191
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def init_distributed(backend: Optional[str] = None) -> Tuple[int, int, int]:
  """Initialise from the torchrun environment. Returns (rank, world_size, local_rank)."""
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  if world > 1 and not dist.is_initialized():
    if backend is None:
      backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    kw = {}
    if backend == 'nccl':
      torch.cuda.set_device(local)
      kw['device_id'] = torch.device(f'cuda:{local}')
    dist.init_process_group(backend, **kw)
  return rank, world, local


def world_size() -> int:
  return dist.get_world_size() if dist.is_initialized() else 1


def shard_rows(n_rows: int, rank: int, world: int) -> slice:
  """Contiguous shard of the batch axis for this rank (rows are independent)."""
  per = (n_rows + world - 1) // world
  lo = min(rank * per, n_rows)
  return slice(lo, min(lo + per, n_rows))


class FlatGradBucket:
  """All gradients of a parameter list viewed as ONE contiguous float32 buffer, so the
  pmean of the whole gradient pytree (ldm/experiment.py:341) is a single all-reduce
  (285 MB for the CIFAR-10 config, 682 MB for ImageNet32) instead of one per tensor."""

  def __init__(self, params: Iterable[torch.nn.Parameter], extra: int = 0):
    self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
    n = sum(p.numel() for p in self.params)
    dev = self.params[0].device if self.params else torch.device('cpu')
    self.extra = extra                         # tail slots (e.g. the six loss scalars)
    self.flat = torch.zeros(n + extra, dtype=torch.float32, device=dev)
    off = 0
    for p in self.params:
      p.grad = self.flat[off:off + p.numel()].view_as(p)   # backward accumulates in place
      off += p.numel()
    self.tail = self.flat[off:off + extra]

  def zero(self):
    self.flat.zero_()

  def all_reduce_mean(self):
    if world_size() > 1:
      dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
      self.flat.div_(world_size())


def pmean_scalars(scalars: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
  """ldm/experiment.py:347-348: pmean of each metric scalar -- packed into one tensor."""
  if world_size() == 1:
    return scalars
  keys = sorted(scalars)
  buf = torch.stack([scalars[k].detach().to(torch.float32).reshape(()) for k in keys])
  dist.all_reduce(buf, op=dist.ReduceOp.SUM)
  buf = buf / world_size()
  return {k: buf[i] for i, k in enumerate(keys)}


def train_step(model, optimizer, bucket: FlatGradBucket, batch: dict, step: int,
               generator: Optional[torch.Generator] = None, draws: Optional[dict] = None):
  """Experiment.train_step (ldm/experiment.py:335-356): value_and_grad(loss_fn), pmean of the
  gradients, optimizer update, pmean of the scalars.  The six scalars ride in the tail of
  the gradient bucket, so one collective follows the backward pass."""
  from .model import loss_fn
  bucket.zero()
  bpd, metrics = loss_fn(model, batch, step=step, is_train=True, draws=draws,
                         generator=generator)
  bpd.backward()
  keys = sorted(metrics['scalars'])
  if bucket.extra >= len(keys):
    with torch.no_grad():
      for i, k in enumerate(keys):
        bucket.tail[i] = metrics['scalars'][k].detach()
    bucket.all_reduce_mean()
    scalars = {k: bucket.tail[i].clone() for i, k in enumerate(keys)}
  else:
    bucket.all_reduce_mean()
    scalars = pmean_scalars(metrics['scalars'])
  optimizer.step()
  return scalars


def dense_images_per_launch(n_timesteps: int, max_rows: int = 16384) -> int:
  """Images per launch for the dense-VLB driver: as many as fit `max_rows` rows.  Short launches
  pay their ramp-up / drain and the partial last wave (2048 rows = 1.7 waves of resident CTAs:
  36 M rows/s on one stream; 16384 rows: 52 M rows/s, profiles/r2_variants.md), and the U-Net
  that sits between the two kernels in a real evaluation is batch-size agnostic."""
  return max(1, max_rows // n_timesteps)


@torch.no_grad()
def eval_bpd_dense_sampling(model, images: torch.Tensor, n_timesteps: int = 128,
                            images_per_launch: Optional[int] = None, seed: int = 0,
                            base_draws: Optional[dict] = None, broadcast_noise: bool = True):
  """eval_bpd_dense_sampling (ldm/notebook_utils.py:176-191), example-sharded.

  For every test image: tile it n_timesteps times (antithetic t gives a stratified
  n_timesteps-point estimate of the diffusion integral), evaluate loss_fn with is_train=False
  and THE SAME key for every image (:178,:185), collect bpd; return the mean.
  `images` is this process's view of the whole test set [N,32,32,3] uint8; each rank takes
  images[rank::world], `images_per_launch` images (x n_timesteps rows) per kernel launch
  (default: dense_images_per_launch).
  base_draws: the four draws of ONE loss_fn call over n_timesteps rows (model.make_draws);
  default: drawn here from `seed` -- the reference uses PRNGKey(0) for every image.
  broadcast_noise: hand eps_0 / eps to the kernels as [n_timesteps, D] (read by row %
  n_timesteps) instead of tiling them over the images of a launch -- same bits, 8 B/sub-pixel
  less HBM traffic in fwd_pre, 4 B less in the post kernel.
  Returns (mean_bpd over all ranks' images, this rank's per-image bpds).
  """
  from .model import sample_t
  rank = dist.get_rank() if dist.is_initialized() else 0
  world = world_size()
  mine = images[rank::world]
  dev = mine.device
  cfg = model.config
  gen = torch.Generator(device=dev).manual_seed(seed)
  base = base_draws if base_draws is not None else model.make_draws(n_timesteps, dev, gen)
  t_img = (base['t'].to(torch.float32).reshape(n_timesteps) if 't' in base
           else sample_t(base['t0'], n_timesteps, cfg))   # one key for every image
  rescale = 1. / (np.prod(images.shape[1:]) * np.log(2.))
  if images_per_launch is None:
    images_per_launch = dense_images_per_launch(n_timesteps)
  bpds = []
  for s in range(0, mine.shape[0], images_per_launch):
    chunk = mine[s:s + images_per_launch]
    m = chunk.shape[0]
    tiled = chunk.repeat_interleave(n_timesteps, dim=0)   # image-major: rows of one image adjacent
    G = base['G']
    # latent noise: [10, n, L] for the gamma draw (tile axis 1), [n, L] for the gumbel /
    # gaussian / additive-noise variants (tile axis 0)
    G = G.repeat(1, m, 1) if G.dim() == 3 else G.repeat(m, 1)
    if broadcast_noise:
      e0, e = base['eps_0'], base['eps']
    else:
      e0, e = base['eps_0'].repeat(m, 1, 1, 1), base['eps'].repeat(m, 1, 1, 1)
    draws = dict(t=t_img.repeat(m), G=G, eps_0=e0, eps=e)
    out = model(tiled, labels=None, conditioning=None, step=0, deterministic=True, draws=draws)
    per_row = out.loss_recon + out.loss_klz + out.loss_diff
    # per image: mean over its rows of each term, summed (ldm/experiment_vdm.py:62-66)
    bpds.append(per_row.reshape(m, n_timesteps).mean(dim=1) * rescale)
  bpds = torch.cat(bpds) if bpds else torch.zeros(0, device=dev)
  acc = torch.stack([bpds.sum().to(torch.float64),
                     torch.tensor(float(bpds.numel()), dtype=torch.float64, device=dev)])
  if world > 1:
    dist.all_reduce(acc, op=dist.ReduceOp.SUM)
  return (acc[0] / acc[1]).item(), bpds
