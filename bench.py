#!/usr/bin/env python
"""bench.py -- MuLAN schedule + ELBO hot path on B200 (see DESIGN.md "Measurement").

A "step" is one pass of the hot path (fwd_pre -> post value-and-grad -> bpd_reduce ->
bwd_pre, i.e. ELBO loss + gradients w.r.t. the schedule coefficients and the denoiser
output) over one batch of synthetic uint8 32x32x3 examples per GPU.  The denoiser (U-Net)
is NOT part of the path: its output and its backward cotangents are supplied
as resident tensors.

  python bench.py [--gpus N --steps K --warmup W]           # this repo's CUDA path
  python bench.py --impl reference [...]                     # reference algorithm on host cores

Prints ONE JSON line (rank 0).  Never run under a profiler for a bench value.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

D = 3072
GROUP = 128          # per-GPU batch of configs[1]; antithetic t is drawn per group
PARAMS = {'eps': 0, 'vel': 1, 'vel_from_eps': 2}
# algorithmic bytes per sub-pixel, kernel -> bytes (DESIGN.md "Kernels"; SURVEY.md 8d rule:
# every declared input read once, every output written once)
ALGO_BYTES = {
    # post_vg = mulan_fwd_bwd_post: loss_diff and n_bar in one pass (value-and-grad)
    'eps': {'fwd_pre': 29, 'fwd_post': 12, 'post_vg': 16, 'bwd_post': 16, 'bwd_pre': 37},
    'vel': {'fwd_pre': 25, 'fwd_post': 21, 'post_vg': 25, 'bwd_post': 25, 'bwd_pre': 37},
    'vel_from_eps': {'fwd_pre': 25, 'fwd_post': 21, 'post_vg': 25, 'bwd_post': 25, 'bwd_pre': 37},
}


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=50)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', choices=['native', 'reference'], default='native')
  ap.add_argument('--rows', type=int, default=128 * GROUP,
                  help='examples per GPU per step (default 128 stacked batches of 128)')
  ap.add_argument('--param', choices=list(PARAMS), default='eps')
  ap.add_argument('--ref-rows', type=int, default=GROUP,
                  help='rows of the bounded CPU sample (cpu_baseline / --impl reference)')
  ap.add_argument('--workload', choices=['train', 'dense_vlb', 'train_step'], default='train',
                  help='train: ELBO loss+grad (configs[1..3]); dense_vlb: forward-only VLB '
                       'evaluation, 16 images x 128 timesteps per launch (configs[4])')
  ap.add_argument('--net-config', choices=['cifar10', 'imagenet32'], default='cifar10',
                  help='train_step: sm_n_embd 128 / 256 stand-in networks (mulan_b200/standin.py)')
  ap.add_argument('--batch', type=int, default=128, help='train_step: per-GPU batch')
  ap.add_argument('--global-batch', type=int, default=0,
                  help='train_step: fixed global batch (strong scaling) instead of --batch')
  ap.add_argument('--tf32', action='store_true',
                  help='train_step: allow TF32 in the stand-in networks (reference: float32)')
  ap.add_argument('--launch-rows', type=int, default=2048,
                  help='dense_vlb: rows per launch (16 images x 128 antithetic timesteps)')
  ap.add_argument('--streams', type=int, default=8,
                  help='streams the independent launches of one step are spread over '
                       '(only matters when a step has several launches: dense_vlb)')
  ap.add_argument('--separate-post', action='store_true',
                  help='train: run mulan_fwd_post and mulan_bwd_post as two passes instead of '
                       'the fused value-and-grad pass')
  ap.add_argument('--no-save-w', action='store_true',
                  help='epsilon form: recompute the loss weight in the post kernels instead of '
                       'saving it in fwd_pre')
  ap.add_argument('--no-e2e', action='store_true')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  return ap.parse_args()


def workload_name(args):
  model = {'eps': 'mulan_epsilon', 'vel': 'mulan_velocity',
           'vel_from_eps': 'mulan_velocity(velocity_from_epsilon)'}[args.param]
  if args.workload == 'dense_vlb':
    return (f'eval_bpd dense VLB forward ({model}: recon + prior + diffusion terms), '
            f'{args.launch_rows // GROUP} images x {GROUP} timesteps per launch, '
            f'{args.rows} rows/step/GPU, synthetic uint8 32x32x3, denoiser output supplied')
  return (f'cifar10-conditioned {model} ELBO loss+grad hot path, per-GPU batch {GROUP} x '
          f'{args.rows // GROUP} stacked batches = {args.rows} rows/step/GPU, synthetic uint8 '
          f'32x32x3, denoiser output supplied')


# ----------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------
class ClockSampler:
  """SM clock + throttle reasons sampled DURING the timed region: NVML in a thread (2 ms
  period, so even a 20 ms region gets samples); nvidia-smi -lms as the fallback."""
  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
       'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
       'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

  def __init__(self, index: int, uuid=None):
    self.index, self.uuid = index, uuid
    self.samples, self.lines, self.proc = [], [], None
    self.smax, self.stop_flag, self.thread, self.how = None, False, None, None

  def start(self):
    try:
      import pynvml as nv
      nv.nvmlInit()
      h = None
      if self.uuid is not None:
        try:
          h = nv.nvmlDeviceGetHandleByUUID(('GPU-' + str(self.uuid)).encode())
        except Exception:
          h = None
      if h is None:
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
      self.smax = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
      self.how = 'nvml'

      def loop():
        while not self.stop_flag:
          try:
            self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                 int(nv.nvmlDeviceGetCurrentClocksEventReasons(h)),
                                 nv.nvmlDeviceGetPowerUsage(h) / 1000.0))
          except Exception:
            pass
          time.sleep(0.002)
      self.thread = threading.Thread(target=loop, daemon=True)
      self.thread.start()
      return
    except Exception:
      self.how = None
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
           '--format=csv,noheader,nounits', '-lms', '20'],
          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.how = 'nvidia-smi'
      threading.Thread(target=self._pump, daemon=True).start()
    except OSError:
      self.proc = None

  def _pump(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def mark(self):
    """Number of samples so far (to slice out the timed region)."""
    return len(self.samples) if self.how == 'nvml' else len(self.lines)

  def stop(self, lo=0, hi=None):
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    sm, reasons, power = [], set(), []
    if self.how == 'nvml':
      self.stop_flag = True
      self.thread.join(timeout=1)
      bits = {'hw_slowdown': 0x8, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20,
              'sw_power_cap': 0x4}
      for clk, r, pw in self.samples[lo:hi]:
        sm.append(clk); power.append(pw)
        for n, bit in bits.items():
          if r & bit:
            reasons.add(n)
    elif self.proc is not None:
      self.proc.terminate()
      try:
        self.proc.wait(timeout=2)
      except subprocess.TimeoutExpired:
        self.proc.kill()
      for ln in self.lines[lo:hi]:
        f = [x.strip() for x in ln.split(',')]
        if len(f) < 9:
          continue
        try:
          sm.append(float(f[1])); self.smax = float(f[2]); power.append(float(f[3]))
        except ValueError:
          continue
        for n, v in zip(names, f[5:9]):
          if v.lower().startswith('active'):
            reasons.add(n)
    else:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no NVML / nvidia-smi']}
    sm.sort()
    return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': self.smax,
            'power_w_max': max(power) if power else None, 'samples': len(sm),
            'reasons': sorted(reasons), 'how': self.how}


# ----------------------------------------------------------------------------------------
# synthetic inputs generated on the device
# ----------------------------------------------------------------------------------------
def make_inputs(rows, device, seed):
  import torch
  g = torch.Generator(device=device).manual_seed(seed)
  rn = lambda: torch.randn((rows, D), generator=g, device=device, dtype=torch.float32)
  x = torch.randint(0, 256, (rows, D), generator=g, device=device, dtype=torch.uint8)
  a, b = rn(), rn()
  c = 1e-3 + torch.nn.functional.softplus(rn())
  eps0, eps = rn(), rn()
  net = eps + 0.3 * rn()
  ngroups = (rows + GROUP - 1) // GROUP
  t0 = torch.rand((ngroups, 1), generator=g, device=device)
  t = torch.remainder(t0 + torch.arange(GROUP, device=device) / GROUP, 1.0).reshape(-1)[:rows]
  z_bar = 1e-4 * rn()
  g_bar = 1e-3 * torch.randn((rows,), generator=g, device=device)
  return dict(x=x, a=a, b=b, c=c, t=t.contiguous(), eps0=eps0, eps=eps, net=net, z_bar=z_bar,
              g_bar=g_bar)


# ----------------------------------------------------------------------------------------
# native arm
# ----------------------------------------------------------------------------------------
def run_native(args):
  import torch
  import torch.distributed as dist
  from mulan_b200 import ops, host, _lib

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
  assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world}'
  torch.cuda.set_device(local)
  dev = torch.device(f'cuda:{local}')
  _lib.load()   # loud failure if the CUDA library is missing

  rows, K, W = args.rows, args.steps, args.warmup
  param = PARAMS[args.param]
  desc = ops.Desc(param=param)
  inp = make_inputs(rows, dev, seed=1234 + rank)
  train = args.workload == 'train'
  lrows = rows if train else min(args.launch_rows, rows)   # rows per launch
  # velocity_from_epsilon evaluates the (algebraically identical) epsilon form: same kernels,
  # same bytes as 'eps' (mulan_kernel_param; MULAN_VFE_LITERAL=1 restores the literal formula)
  eps_form = _lib.kernel_param(param) == PARAMS['eps']
  save_w = eps_form and not args.no_save_w   # fwd_pre +4 B, post -8 B per sub-pixel
  assert rows % lrows == 0
  chunks = []
  for s0 in range(0, rows, lrows):
    ci = {k: v[s0:s0 + lrows] for k, v in inp.items()}
    chunks.append((ops.ElboWorkspace(desc, lrows, dev, save_w=save_w),
                   ci, torch.full((lrows,), 1.0 / (lrows * D * math.log(2.0)), device=dev)))
  ws = chunks[0][0]

  # Several launches per step (dense VLB: 2048 rows each = 3.46 waves of 592 resident CTAs) are
  # independent, so they go round-robin over a few streams: the ragged last wave of one launch
  # overlaps the first wave of the next.  Fork/join by events, capturable in the CUDA graph.
  n_side = min(len(chunks), args.streams) - 1
  side = [torch.cuda.Stream() for _ in range(max(n_side, 0))]

  def fan_out(per_chunk):
    cur = torch.cuda.current_stream()
    if not side:
      for ch in chunks:
        per_chunk(*ch)
      return
    fork = torch.cuda.Event()
    fork.record(cur)
    lanes = [cur] + side
    for s in side:
      s.wait_event(fork)
    for j, ch in enumerate(chunks):
      with torch.cuda.stream(lanes[j % len(lanes)]):
        per_chunk(*ch)
    for s in side:
      join = torch.cuda.Event()
      join.record(s)
      cur.wait_event(join)

  def each(fn):
    return lambda: fan_out(fn)
  per_chunk = {   # name -> launch on one chunk (all write into the preallocated workspaces)
      'fwd_pre': lambda w_, i, g: w_.fwd_pre(i['x'], i['a'], i['b'], i['c'], i['t'],
                                             i['eps0'], i['eps']),
  }
  if train and not args.separate_post:
    # value-and-grad: the loss cotangent of a mean is known up front (jax.value_and_grad)
    per_chunk['post_vg'] = lambda w_, i, g: w_.fwd_bwd_post(i['x'], i['a'], i['b'], i['c'],
                                                            i['t'], i['eps'], i['net'], g)
  else:
    per_chunk['fwd_post'] = lambda w_, i, g: w_.fwd_post(i['x'], i['a'], i['b'], i['c'], i['t'],
                                                         i['eps'], i['net'])
  per_chunk['bpd_reduce'] = lambda w_, i, g: w_.bpd_reduce(None)
  if train:
    if args.separate_post:
      per_chunk['bwd_post'] = lambda w_, i, g: w_.bwd_post(i['x'], i['a'], i['b'], i['c'],
                                                           i['t'], i['eps'], i['net'], g)
    per_chunk['bwd_pre'] = lambda w_, i, g: w_.bwd_pre(i['x'], i['a'], i['b'], i['c'], i['t'],
                                                       i['eps'], i['net'], i['z_bar'],
                                                       i['g_bar'], g)
  names = list(per_chunk)
  kernels = {n: each(fn) for n, fn in per_chunk.items()}   # one kernel over every chunk
  launches_per_step = len(names) * len(chunks)

  def whole_chunk(w_, i, g):
    for n in names:
      per_chunk[n](w_, i, g)

  def step():
    fan_out(whole_chunk)                  # chunk-major: a chunk's kernels stay in stream order

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  # ---- warm-up (eager), then capture ONE step in a CUDA graph: the timed loop replays it,
  #      so host launch latency / Python jitter cannot starve the GPU -------------------
  stream = torch.cuda.Stream()
  with torch.cuda.stream(stream):
    for _ in range(max(W, 3)):
      step()
      if world > 1:
        dist.all_reduce(ws.scalars, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=stream):
      step()
    graph.replay()
  barrier()

  try:
    uuid = torch.cuda.get_device_properties(local).uuid
  except Exception:
    uuid = None
  sampler = ClockSampler(local, uuid)
  if rank == 0:
    sampler.start()
    time.sleep(0.05)
  t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  with torch.cuda.stream(stream):
    barrier()
    mark0 = sampler.mark()
    t_start.record(stream)
    works, reduced = [], torch.empty((K, 6), dtype=torch.float32, device=dev)
    for k in range(K):
      graph.replay()
      if world > 1:
        # the one exchange that follows the path: pmean of the six loss scalars
        # (ldm/experiment.py:347-348).  Nothing downstream waits for it, so it runs on NCCL's
        # own stream from a snapshot of the scalars and overlaps the next step.
        reduced[k].copy_(ws.scalars)
        works.append(dist.all_reduce(reduced[k], op=dist.ReduceOp.AVG, async_op=True))
    for wk in works:
      wk.wait()                      # the timed region ends only when every pmean has landed
    t_end.record(stream)
    barrier()
    mark1 = sampler.mark()
  elapsed_ms = t_start.elapsed_time(t_end)
  tmax = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
  if world > 1:
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
  elapsed_ms = tmax.item()
  timed_launches = launches_per_step * K
  bpd = (reduced[-1][0] if world > 1 else ws.scalars[0]).item()

  # ---- per-kernel durations, live, CUDA events on the launching stream: each kernel K times
  #      back to back (its inputs alone exceed L2, so every launch streams from HBM) --------
  kern_ms = {}
  with torch.cuda.stream(stream):
    for n in names:
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      kernels[n]()
      torch.cuda.synchronize()
      e0.record(stream)
      for _ in range(K):
        kernels[n]()
      e1.record(stream)
      torch.cuda.synchronize()
      kern_ms[n] = e0.elapsed_time(e1) / K
  clocks = sampler.stop(mark0, None) if rank == 0 else None

  # ---- e2e: host buffers through the C ABI (mulan_elbo_host), copies inside the timing ----
  e2e = None
  if not args.no_e2e and train:
    pin = lambda v: v.cpu().pin_memory()
    h = {k: pin(inp[k]) for k in ('x', 'a', 'b', 'c', 't', 'eps0', 'eps', 'net')}
    out = host.HostOutputs(rows, D, want_grad=True, pinned=True)
    call = lambda: host.elbo_host(h['x'], h['a'], h['b'], h['c'], h['t'], h['eps0'], h['eps'],
                                  h['net'], param=param, want_grad=True, out=out)
    ke = max(2, min(K, 5))
    call(); call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(ke):
      r = call()           # synchronises before returning
    t1 = time.perf_counter()
    te = torch.tensor([t1 - t0], device=dev, dtype=torch.float64)
    if world > 1:
      dist.all_reduce(te, op=dist.ReduceOp.MAX)
    h2d = rows * D * (1 + 6 * 4) + rows * 4
    d2h = rows * D * 4 * 4 + (3 * rows + 6) * 4
    e2e = {'value': world * rows * ke / te.item(), 'unit': 'samples/s', 'steps': ke,
           'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
           'api': 'mulan_elbo_host (C ABI, pinned host buffers)',
           'bpd': float(r['scalars'][0])}
    # the same call with eps_0 / eps drawn on the device from their threefry keys (what
    # VDM.__call__ itself does with its rng): 8 of the 25 H2D bytes per sub-pixel stay home
    callk = lambda: host.elbo_host(h['x'], h['a'], h['b'], h['c'], h['t'], None, None, h['net'],
                                   param=param, want_grad=True, out=out,
                                   jax_keys=((1234, rank), (5678, rank)))
    callk(); callk()
    barrier()
    t0 = time.perf_counter()
    for _ in range(ke):
      rk = callk()
    t1 = time.perf_counter()
    tk = torch.tensor([t1 - t0], device=dev, dtype=torch.float64)
    if world > 1:
      dist.all_reduce(tk, op=dist.ReduceOp.MAX)
    e2e['device_draws'] = {
        'value': world * rows * ke / tk.item(), 'unit': 'samples/s',
        'h2d_bytes_per_step': rows * D * (1 + 4 * 4) + rows * 4 + 16,
        'd2h_bytes_per_step': d2h, 'api': 'mulan_elbo_host_keyed (eps_0, eps from JAX keys)',
        'bpd': float(rk['scalars'][0])}
    _lib.load().mulan_host_workspace_release()

  # ---- latency of configs[1]'s literal size (one batch of 128 rows, CUDA graph) ----
  lat = None
  if rank == 0 and train:
    sm = {k: (v[:GROUP].contiguous()) for k, v in inp.items()}
    gLs = torch.full((GROUP,), 1.0 / (GROUP * D * math.log(2.0)), device=dev)
    wss = ops.ElboWorkspace(desc, GROUP, dev)
    def small_step():
      wss.fwd_pre(sm['x'], sm['a'], sm['b'], sm['c'], sm['t'], sm['eps0'], sm['eps'])
      wss.fwd_bwd_post(sm['x'], sm['a'], sm['b'], sm['c'], sm['t'], sm['eps'], sm['net'], gLs)
      wss.bpd_reduce(None)
      wss.bwd_pre(sm['x'], sm['a'], sm['b'], sm['c'], sm['t'], sm['eps'], sm['net'], sm['z_bar'],
                  sm['g_bar'], gLs)
    with torch.cuda.stream(stream):
      for _ in range(3):
        small_step()
      torch.cuda.synchronize()
      g2 = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g2, stream=stream):
        small_step()
      for _ in range(5):
        g2.replay()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(stream)
      for _ in range(200):
        g2.replay()
      e1.record(stream)
      torch.cuda.synchronize()
      us = e0.elapsed_time(e1) * 1000 / 200
      lat = {'rows': GROUP, 'us_per_step': us, 'samples_per_s': GROUP / (us * 1e-6),
             'how': 'CUDA graph of the 4 launches, 200 replays, L2-resident'}

  # ---- cpu baseline (oracle port on host cores; rank 0, N=1 only) ----
  cpu = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    cpu = cpu_reference(args, steps=0, warmup=1, min_seconds=10.0)

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  nsub = rows * D
  dom = max((n for n in names if n != 'bpd_reduce'), key=lambda n: kern_ms[n])
  ab = dict(ALGO_BYTES['eps' if eps_form else args.param])
  if eps_form and not save_w:
    ab.update(fwd_pre=25, fwd_post=20, post_vg=24, bwd_post=24)   # w recomputed from a, b, c
  ab = {k: v for k, v in ab.items() if k in names}
  peaks, peak_src = load_peak()
  kinfo = {}
  for n in names:
    if n == 'bpd_reduce':
      kinfo[n] = {'ms': kern_ms[n]}
      continue
    gbs = ab[n] * nsub / (kern_ms[n] * 1e-3) / 1e9
    kinfo[n] = {'ms': kern_ms[n], 'algo_bytes_per_subpixel': ab[n], 'gbs': gbs,
                'frac_of_measured': gbs / peaks, 'frac_of_8TBs': gbs / 8000.0}
  total_algo = sum(ab.values()) * nsub
  traffic = load_traffic(dom, rows) if train else None
  ms_step = elapsed_ms / K
  line = {
      'metric': 'mulan_elbo_train_samples_per_s' if train else 'mulan_dense_vlb_rows_per_s',
      'value': world * rows * K / (elapsed_ms * 1e-3),
      'unit': 'samples/s', 'n_gpus': world, 'steps': K, 'warmup': max(W, 3),
      'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': workload_name(args), 'rows_per_gpu': rows, 'dim': D,
                 'param': args.param, 'loss_form': 'eps' if eps_form else 'velocity',
                 'saved_w': save_w, 'l2': 'inputs larger than L2 (%.2f GB of HBM traffic per '
                 'step)' % (total_algo / 1e9), 'parallelism': f'dp{world} (rows sharded)',
                 'timed_loop': 'CUDA-graph replay of one step (%d launches)' % launches_per_step,
                 'rows_per_launch': lrows, 'streams': 1 + len(side)},
      'roofline': {'bound': 'hbm', 'kernel': dom, 'achieved': kinfo[dom]['gbs'],
                   'peak': peaks, 'peak_source': peak_src, 'unit': 'GB/s',
                   'frac': kinfo[dom]['gbs'] / peaks, 'traffic': traffic,
                   'algo_bytes_per_launch': ab[dom] * nsub,
                   'how': 'CUDA events around %d back-to-back launches on the launch stream' % K},
      'step_hbm': {'algo_bytes_per_step': total_algo,
                   'gbs': total_algo / (ms_step * 1e-3) / 1e9,
                   'frac_of_measured': total_algo / (ms_step * 1e-3) / 1e9 / peaks,
                   'sum_kernel_ms': sum(kern_ms.values())},
      'kernels': kinfo, 'gpu_launches': timed_launches, 'clocks': clocks, 'e2e': e2e,
      'latency_b128': lat, 'cpu_baseline': cpu, 'bpd': bpd,
  }
  emit(line)
  if world > 1:
    dist.destroy_process_group()


def load_peak():
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    try:
      return float(json.load(open(p))['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (of measured)'
    except Exception:
      pass
  return 6650.0, 'B200_PROFILING.md fallback 6.65 TB/s (of fallback)'


def load_traffic(kernel, rows):
  """dram bytes per launch from the committed ncu --set full capture (profiles/), or None."""
  p = os.path.join(ROOT, 'profiles', 'traffic.json')
  if not os.path.exists(p):
    return None
  try:
    t = json.load(open(p)).get(kernel)
    if t and t.get('rows') == rows:
      return t['dram_bytes_per_launch']
  except Exception:
    return None
  return None


# ----------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference on the host cores
# ----------------------------------------------------------------------------------------
def cpu_reference(args, steps, warmup, min_seconds=0.0):
  import torch
  from oracle import mulan_oracle as O   # CPU baseline leg: allowed to execute oracle/
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  B = args.ref_rows
  mode = PARAMS[args.param]
  inp = O.synth_inputs(B, seed=0)
  cfg = O.OracleConfig()

  def one():
    a, b, c, net = (inp[k].clone().requires_grad_(True) for k in ('a', 'b', 'c', 'net'))
    out = O.elbo_terms(inp['x'], a, b, c, inp['t'], inp['eps_0'], inp['eps'],
                       lambda z, g: net, mode, cfg)
    bpd, _ = O.loss_fn_bpd(out)
    torch.autograd.grad(bpd, [a, b, c, net])
    return bpd.item()

  for _ in range(warmup):
    one()
  t0 = time.perf_counter()
  if min_seconds:        # bounded sample: whole steps until ~min_seconds of CPU work are in
    steps = 0
    while steps < 2 or (time.perf_counter() - t0 < min_seconds and steps < 256):
      one()
      steps += 1
  else:
    for _ in range(steps):
      one()
  dt = time.perf_counter() - t0
  return {'value': B * steps / dt, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
          'ms_per_step': dt / steps * 1e3,
          'sample': f'{steps} steps of {B} rows (one per-GPU batch of the workload), oracle '
                    f'float32 loss+grad with torch CPU, {cores} threads'}


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  cpu = cpu_reference(args, steps=args.steps, warmup=args.warmup)
  line = {
      'impl': 'reference', 'metric': 'mulan_elbo_train_samples_per_s', 'value': cpu['value'],
      'unit': 'samples/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': cpu['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': workload_name(args), 'param': args.param,
                 'note': 'reference algorithm (CPU oracle port; JAX is not installable in this '
                         'image) on the host cores, bounded sample per step'},
      'cpu_baseline': cpu,
      'e2e': {'value': cpu['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0,
              'd2h_bytes_per_step': 0},
      'gpu_launches': 0,
  }
  emit(line)


_REAL_STDOUT = None


def emit(line: dict):
  """The ONE JSON line, on the real stdout (libraries such as NCCL print banners to fd 1;
  main() points fd 1 at stderr for the duration of the run)."""
  data = (json.dumps(line) + '\n').encode()
  if _REAL_STDOUT is not None:
    os.write(_REAL_STDOUT, data)
  else:
    sys.stdout.write(data.decode())
    sys.stdout.flush()


def run_train_step(args):
  """Whole train step around the kernels (BASELINE.json configs[1..3]): stand-in encoder and
  U-Net on cuDNN/cuBLAS float32 (mulan_b200/standin.py -- the reference keeps them on the
  framework path), the ELBO kernels, ONE all-reduce of the flat gradient bucket (+ scalars),
  ONE fused AdamW+EMA launch.  Reports samples/s and the ELBO kernels' share of the step."""
  import torch
  import torch.distributed as dist
  from mulan_b200 import _lib, ops
  from mulan_b200.model import VDM, VDMConfig
  from mulan_b200.optim import FlatTrainState, train_step
  from mulan_b200.standin import ScoreUNet, UnetEncoder

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
  assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world}'
  torch.cuda.set_device(local)
  dev = torch.device(f'cuda:{local}')
  _lib.load()
  torch.backends.cuda.matmul.allow_tf32 = args.tf32      # reference: matmul precision float32
  torch.backends.cudnn.allow_tf32 = args.tf32
  torch.backends.cudnn.benchmark = True
  torch.manual_seed(1234)                                 # same init on every rank
  n_embd = 128 if args.net_config == 'cifar10' else 256
  B = args.global_batch // world if args.global_batch else args.batch
  cfg = VDMConfig(vdm_type='mulan_epsilon' if args.param == 'eps' else 'mulan_velocity',
                  velocity_from_epsilon=(args.param == 'vel_from_eps'))
  model = VDM(cfg, UnetEncoder(n_embd, 4), ScoreUNet(n_embd, 32)).to(dev)
  model.train()
  state = FlatTrainState(model.named_parameters())
  gen = torch.Generator(device=dev).manual_seed(100 + rank)
  host_images = torch.randint(0, 256, (B, 32, 32, 3), dtype=torch.uint8).pin_memory()
  K, W = args.steps, max(args.warmup, 3)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def one_step():
    images = host_images.to(dev, non_blocking=True)       # H2D of the step's inputs
    return train_step(model, state, {'images': images}, generator=gen)

  for _ in range(W):
    sc = one_step()
  barrier()
  try:
    uuid = torch.cuda.get_device_properties(local).uuid
  except Exception:
    uuid = None
  sampler = ClockSampler(local, uuid)
  if rank == 0:
    sampler.start()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  mark0 = sampler.mark()
  t0 = time.perf_counter()
  e0.record()
  for _ in range(K):
    sc = one_step()
    bpd = float(sc['bpd'])                                # D2H read of the step's result
  e1.record()
  barrier()
  wall = time.perf_counter() - t0
  clocks = sampler.stop(mark0, None) if rank == 0 else None
  tm = torch.tensor([e0.elapsed_time(e1), wall * 1e3], device=dev, dtype=torch.float64)
  if world > 1:
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
  ms, wall_ms = tm[0].item() / K, tm[1].item() / K

  # ELBO kernels alone at this batch (CUDA graph of the 4 launches)
  elbo_us = None
  if rank == 0:
    desc = model.desc
    inp = make_inputs(B, dev, seed=7)
    ws = ops.ElboWorkspace(desc, B, dev)
    gL = torch.full((B,), 1.0 / (B * D * math.log(2.0)), device=dev)
    def elbo():
      ws.fwd_pre(inp['x'], inp['a'], inp['b'], inp['c'], inp['t'], inp['eps0'], inp['eps'])
      ws.fwd_bwd_post(inp['x'], inp['a'], inp['b'], inp['c'], inp['t'], inp['eps'], inp['net'], gL)
      ws.bpd_reduce(None)
      ws.bwd_pre(inp['x'], inp['a'], inp['b'], inp['c'], inp['t'], inp['eps'], inp['net'],
                 inp['z_bar'], inp['g_bar'], gL)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
      for _ in range(3):
        elbo()
      torch.cuda.synchronize()
      g2 = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g2, stream=st):
        elbo()
      a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a0.record(st)
      for _ in range(100):
        g2.replay()
      a1.record(st)
      torch.cuda.synchronize()
      elbo_us = a0.elapsed_time(a1) * 10.0
  # fused AdamW+EMA alone (36 B per parameter) and, for context, torch's fused AdamW + foreach EMA
  optim = None
  if rank == 0:
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    state.apply_gradients()
    torch.cuda.synchronize()
    o0.record()
    for _ in range(20):
      state.apply_gradients()
    o1.record()
    torch.cuda.synchronize()
    us = o0.elapsed_time(o1) * 1000 / 20
    peaks, peak_src = load_peak()
    gbs = 36.0 * state.n / (us * 1e-6) / 1e9
    optim = {'kernel': 'mulan_adamw_ema', 'us': us, 'algo_bytes': 36 * state.n, 'gbs': gbs,
             'frac_of_measured': gbs / peaks}
    try:
      tp = [torch.nn.Parameter(torch.randn(state.n // 8, device=dev)) for _ in range(8)]
      for q in tp:
        q.grad = torch.randn_like(q)
      te = [q.detach().clone() for q in tp]
      topt = torch.optim.AdamW(tp, lr=2e-4, betas=(0.9, 0.99), weight_decay=0.01, fused=True)
      def torch_step():
        topt.step()
        torch._foreach_lerp_(te, [q.detach() for q in tp], 1e-4)
      torch_step()
      torch.cuda.synchronize()
      o0.record()
      for _ in range(20):
        torch_step()
      o1.record()
      torch.cuda.synchronize()
      optim['torch_fused_adamw_plus_foreach_ema_us'] = o0.elapsed_time(o1) * 1000 / 20
      del tp, te, topt
    except Exception as exc:      # context figure only
      optim['torch_fused_adamw_plus_foreach_ema_us'] = f'unavailable: {exc}'
  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return
  nparam = state.n
  line = {
      'metric': 'mulan_train_step_samples_per_s', 'value': world * B / (ms * 1e-3),
      'unit': 'samples/s', 'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms,
      'higher_is_better': True, 'scaling': 'strong' if args.global_batch else 'weak',
      'vs_baseline': None, 'dtype': 'tf32' if args.tf32 else 'f32', 'data': 'synthetic',
      'config': {'workload': f'{args.net_config} {cfg.vdm_type}'
                             f'{"(velocity_from_epsilon)" if cfg.velocity_from_epsilon else ""} '
                             f'train step, per-GPU batch {B}, global {B * world}; STAND-IN '
                             f'encoder/U-Net on cuDNN/cuBLAS (mulan_b200/standin.py), ELBO in '
                             f'libmulan_b200, flat-bucket all-reduce, fused AdamW+EMA',
                 'sm_n_embd': n_embd, 'sm_n_layer': 32, 'parameters': nparam,
                 'grad_allreduce_bytes': 4 * (nparam + state.extra), 'parallelism': f'dp{world}'},
      'elbo_kernels': {'us_per_step': elbo_us, 'share_of_step': elbo_us / (ms * 1e3),
                       'how': 'CUDA graph of fwd_pre, post value-and-grad, bpd_reduce, bwd_pre '
                              'at this batch'},
      'optimizer': optim,
      'e2e': {'value': world * B / (wall_ms * 1e-3), 'unit': 'samples/s',
              'h2d_bytes_per_step': B * D, 'd2h_bytes_per_step': 4,
              'api': 'optim.train_step(model, state, batch) with host uint8 images'},
      'gpu_launches': 7 * K, 'clocks': clocks, 'bpd': bpd,   # aux fwd/bwd, fwd_pre, post_vg, scale_rows, bwd_pre, adamw_ema
  }
  emit(line)
  if world > 1:
    dist.destroy_process_group()


def main():
  global _REAL_STDOUT
  args = parse_args()
  sys.stdout.flush()
  _REAL_STDOUT = os.dup(1)
  os.dup2(2, 1)
  if args.impl == 'reference':
    run_reference(args)
  elif args.workload == 'train_step':
    run_train_step(args)
  else:
    run_native(args)


if __name__ == '__main__':
  main()
