#!/usr/bin/env python
"""Times the JAX-compatible device draws against torch's Philox randn for the bench's shard
(16384 x 3072 sub-pixels).  CUDA events, 20 launches each; prints one JSON line."""
import json
import sys

import torch

sys.path.insert(0, '.')
from mulan_b200 import ops  # noqa: E402

dev = torch.device('cuda:0')
n = 16384 * 3072
out = torch.empty(n, device=dev)


def timed(fn, k=20):
  fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(k):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / k * 1e3


res = {
    'n': n,
    'rng_normal_us': timed(lambda: ops.rng_normal((1, 2), (n,), out=out)),
    'rng_uniform_us': timed(lambda: ops.rng_uniform((1, 2), (n,), out=out)),
    'rng_bits_us': timed(lambda: ops.rng_bits((1, 2), (n,), out=out.view(torch.int32))),
    'torch_randn_us': timed(lambda: out.normal_()),
    'hbm_write_floor_us': n * 4 / 6542.7e9 * 1e6,
}
print(json.dumps(res))
