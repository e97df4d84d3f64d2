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
  ap.add_argument('--no-configs', action='store_true',
                  help='skip the `configs` block (velocity, v-from-eps, dense VLB, train step)')
  ap.add_argument('--no-train-step', action='store_true',
                  help='skip the train-step legs of the `configs` block')
  ap.add_argument('--no-imagenet', action='store_true',
                  help='skip the ImageNet-32 train-step leg at 8 GPUs')
  ap.add_argument('--train-steps', type=int, default=3,
                  help='timed steps of each train-step leg of the `configs` block')
  ap.add_argument('--comm', choices=['all', 'allreduce', 'overlap', 'peer'], default='all',
                  help='train_step: how the gradient exchange runs (mulan_b200/optim.py)')
  ap.add_argument('--nccl-scalars', action='store_true',
                  help='N > 1: pmean of the six loss scalars as one NCCL all-reduce per step '
                       '(round 1) instead of the peer-memory board written by the post kernel')
  ap.add_argument('--no-pdl', action='store_true',
                  help='plain launches instead of programmatic dependent launch')
  ap.add_argument('--no-sustained', action='store_true',
                  help='skip the ~2 s sustained replay of the step graph')
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

  def window(self, lo, hi):
    """Median SM clock / max power / throttle reasons of samples [lo, hi) (NVML mode only)."""
    if self.how != 'nvml':
      return None
    bits = {'hw_slowdown': 0x8, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20,
            'sw_power_cap': 0x4}
    sm = sorted(c for c, _, _ in self.samples[lo:hi])
    reasons = sorted({n for _, r, _ in self.samples[lo:hi] for n, b in bits.items() if r & b})
    pw = [p for _, _, p in self.samples[lo:hi]]
    return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': self.smax,
            'power_w_max': max(pw) if pw else None, 'samples': len(sm), 'reasons': reasons}

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
class Ctx:
  """Process-wide state of the native arm: rank / device / torch.distributed."""

  def __init__(self, args):
    import torch
    import torch.distributed as dist
    from mulan_b200 import _lib
    self.torch, self.dist = torch, dist
    self.world = int(os.environ.get('WORLD_SIZE', '1'))
    self.rank = int(os.environ.get('RANK', '0'))
    self.local = int(os.environ.get('LOCAL_RANK', '0'))
    bind_to_gpu_numa(self.local)      # before any pinned allocation (first touch decides)
    if self.world > 1:
      dist.init_process_group('nccl', device_id=torch.device(f'cuda:{self.local}'))
    assert self.world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={self.world}'
    torch.cuda.set_device(self.local)
    self.dev = torch.device(f'cuda:{self.local}')
    _lib.load()   # loud failure if the CUDA library is missing
    try:
      self.uuid = torch.cuda.get_device_properties(self.local).uuid
    except Exception:
      self.uuid = None
    self.peak, self.peak_src = load_peak()

  def barrier(self):
    if self.world > 1:
      self.dist.barrier()
    self.torch.cuda.synchronize()

  def max_over_ranks(self, value):
    t = self.torch.tensor([value], device=self.dev, dtype=self.torch.float64)
    if self.world > 1:
      self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
    return t.item()

  def close(self):
    if self.world > 1:
      self.dist.destroy_process_group()


def bind_to_gpu_numa(local_rank):
  """Run this rank on the CPUs of its GPU's NUMA node, so that page-locked host buffers
  (first-touch placement) sit next to the PCIe root the GPU hangs off.  A no-op on single-node
  hosts (the B200 boxes of this pool expose ONE NUMA node) and when sysfs has no answer."""
  info = {'numa_node': None, 'cpus': None, 'bound': False}
  try:
    import torch
    bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
    dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
    dev_id = torch.cuda.get_device_properties(local_rank).pci_device_id
    path = f'/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0'
    node = int(open(path + '/numa_node').read().strip())
    info['numa_node'] = node
    nodes = [d for d in os.listdir('/sys/devices/system/node') if d.startswith('node')]
    if node >= 0 and len(nodes) > 1:
      cpus = set()
      for part in open(f'/sys/devices/system/node/node{node}/cpulist').read().strip().split(','):
        lo, _, hi = part.partition('-')
        cpus.update(range(int(lo), int(hi or lo) + 1))
      cpus &= os.sched_getaffinity(0)
      if cpus:
        os.sched_setaffinity(0, cpus)
        info.update(cpus=len(cpus), bound=True)
  except Exception as exc:      # best effort: the measurement proceeds unbound
    info['error'] = repr(exc)[:80]
  bind_to_gpu_numa.info = info
  return info


bind_to_gpu_numa.info = {}


def measure_elbo(ctx, args, inp, param_name, K, W, extras):
  """ELBO loss + gradients (fwd_pre -> post value-and-grad with the loss-scalar reduction in its
  epilogue -> bwd_pre) over `rows` examples per GPU for one parameterisation: whole-step
  throughput (CUDA-graph replay, max over ranks) and per-kernel roofline fractions."""
  torch, dist = ctx.torch, ctx.dist
  from mulan_b200 import ops, _lib
  dev, world, rank = ctx.dev, ctx.world, ctx.rank
  rows = args.rows
  param = PARAMS[param_name]
  eps_form = _lib.kernel_param(param) == PARAMS['eps']
  save_w = eps_form and not args.no_save_w       # fwd_pre +4 B, post -8 B per sub-pixel
  desc = ops.Desc(param=param, pdl=not args.no_pdl)
  ws = ops.ElboWorkspace(desc, rows, dev, save_w=save_w)
  gL = torch.full((rows,), 1.0 / (rows * D * math.log(2.0)), device=dev)
  i = inp
  # N > 1: the pmean of the six loss scalars (ldm/experiment.py:347-348) is an all-gather by peer
  # stores in the post kernel's epilogue (mulan_post_bpd_peer) -- no collective call per step
  board = None
  if world > 1 and not args.nccl_scalars and not args.separate_post:
    from mulan_b200.peer import ScalarBoard
    try:
      board = ScalarBoard(dev)
    except RuntimeError as exc:      # raised on EVERY rank together (peer.PeerAllocations)
      if rank == 0:
        print(f'[bench] scalar board unavailable, NCCL all-reduce per step instead: {exc}',
              file=sys.stderr)
  kernels = {'fwd_pre': lambda: ws.fwd_pre(i['x'], i['a'], i['b'], i['c'], i['t'], i['eps0'],
                                            i['eps'])}
  if args.separate_post:
    kernels['fwd_post'] = lambda: ws.fwd_post(i['x'], i['a'], i['b'], i['c'], i['t'], i['eps'],
                                              i['net'])
    kernels['bpd_reduce'] = lambda: ws.bpd_reduce(None)
    kernels['bwd_post'] = lambda: ws.bwd_post(i['x'], i['a'], i['b'], i['c'], i['t'], i['eps'],
                                              i['net'], gL)
  else:
    # value-and-grad: the loss cotangent of a mean is known up front (jax.value_and_grad); the
    # six loss_fn scalars come out of the same launch
    kernels['post_vg'] = lambda: ws.post_bpd(i['x'], i['a'], i['b'], i['c'], i['t'], i['eps'],
                                             i['net'], gL, board=board)
  kernels['bwd_pre'] = lambda: ws.bwd_pre(i['x'], i['a'], i['b'], i['c'], i['t'], i['eps'],
                                          i['net'], i['z_bar'], i['g_bar'], gL)
  names = list(kernels)

  def step():
    for n in names:
      kernels[n]()

  stream = torch.cuda.Stream()
  with torch.cuda.stream(stream):
    for _ in range(max(W, 3)):
      step()
      if world > 1 and board is None:
        dist.all_reduce(ws.scalars, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=stream):
      step()
    graph.replay()
  ctx.barrier()

  sampler = ClockSampler(ctx.local, ctx.uuid) if (rank == 0 and extras) else None
  if sampler:
    sampler.start()
    time.sleep(0.05)
  t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  with torch.cuda.stream(stream):
    ctx.barrier()
    mark0 = sampler.mark() if sampler else 0
    t_start.record(stream)
    works, reduced = [], torch.empty((K, 6), dtype=torch.float32, device=dev)
    for k in range(K):
      graph.replay()
      if world > 1 and board is None:
        # round 1's exchange: one NCCL all-reduce of the six scalars per step, on NCCL's own
        # stream from a snapshot of the scalars
        reduced[k].copy_(ws.scalars)
        works.append(dist.all_reduce(reduced[k], op=dist.ReduceOp.AVG, async_op=True))
    for wk in works:
      wk.wait()                      # the timed region ends only when every pmean has landed
    if board is not None:
      # every step's scalars were stored into every rank's board by the post kernel; the read
      # waits until the LAST step's rows of all ranks have landed here and averages them
      pmean, pstep = board.read()
    t_end.record(stream)
    ctx.barrier()
    mark1 = sampler.mark() if sampler else 0
  elapsed_ms = ctx.max_over_ranks(t_start.elapsed_time(t_end))
  if board is not None:
    bpd = pmean[0].item()
    assert int(pstep.item()) > 0, 'scalar board: a peer row was missing (timed out / lapped)'
  else:
    bpd = (reduced[-1][0] if world > 1 else ws.scalars[0]).item()

  # per-kernel durations, live, CUDA events on the launching stream: each kernel K times back
  # to back (its inputs alone exceed L2, so every launch streams from HBM)
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

  # sustained: the same graph replayed for ~2 s (clocks and power under a long load)
  sustained = None
  if extras and not args.no_sustained:
    n_rep = max(int(2000.0 / (elapsed_ms / K)), K)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
      ctx.barrier()
      m0 = sampler.mark() if sampler else 0
      e0.record(stream)
      for _ in range(n_rep):
        graph.replay()
      e1.record(stream)
      ctx.barrier()
      m1 = sampler.mark() if sampler else 0
    ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    sustained = {'replays': n_rep, 'seconds': ms * 1e-3,
                 'value': world * rows * n_rep / (ms * 1e-3), 'unit': 'samples/s',
                 'ms_per_step': ms / n_rep}
    if sampler:
      sustained['clocks'] = sampler.window(m0, m1)
  clocks = sampler.stop(mark0, max(mark1, mark0 + 1)) if sampler else None

  nsub = rows * D
  ab = dict(ALGO_BYTES['eps' if eps_form else param_name])
  if eps_form and not save_w:
    ab.update(fwd_pre=25, fwd_post=20, post_vg=24, bwd_post=24)   # w recomputed from a, b, c
  ab = {k: v for k, v in ab.items() if k in names}
  kinfo = {}
  for n in names:
    if n not in ab:
      kinfo[n] = {'ms': kern_ms[n]}
      continue
    gbs = ab[n] * nsub / (kern_ms[n] * 1e-3) / 1e9
    kinfo[n] = {'ms': kern_ms[n], 'algo_bytes_per_subpixel': ab[n], 'gbs': gbs,
                'frac_of_measured': gbs / ctx.peak, 'frac_of_8TBs': gbs / 8000.0}
  total_algo = sum(ab.values()) * nsub
  ms_step = elapsed_ms / K
  dom = max(ab, key=lambda n: kern_ms[n])
  if board is not None:
    del graph
    board.close()
  return {
      'param': param_name, 'loss_form': 'eps' if eps_form else 'velocity', 'saved_w': save_w,
      'value': world * rows * K / (elapsed_ms * 1e-3), 'unit': 'samples/s',
      'ms_per_step': ms_step, 'steps': K, 'rows_per_gpu': rows, 'bpd': bpd,
      'launches_per_step': len(names), 'kernels': kinfo, 'dominant': dom,
      'step_hbm': {'algo_bytes_per_step': total_algo,
                   'gbs': total_algo / (ms_step * 1e-3) / 1e9,
                   'frac_of_measured': total_algo / (ms_step * 1e-3) / 1e9 / ctx.peak,
                   'sum_kernel_ms': sum(kern_ms.values())},
      'clocks': clocks, 'sustained': sustained, 'algo_bytes': ab,
      'scalar_pmean': ('none (1 GPU)' if world == 1 else
                       'peer-memory board written by the post kernel (mulan_post_bpd_peer), read '
                       'once after the timed steps' if board is not None else
                       'one 24-byte NCCL all-reduce per step'),
  }


def measure_latency(ctx, args, inp):
  """The literal per-GPU batch of configs[1] (128 rows): GPU time per step from a CUDA graph that
  holds 20 consecutive steps (a one-step graph replayed back to back measures the host's
  graph-launch rate, ~18 us, not the device)."""
  torch = ctx.torch
  from mulan_b200 import ops
  dev = ctx.dev
  out = {}
  for label, pdl in (('plain', False), ('pdl', True)):
    sm = {k: (v[:GROUP].contiguous()) for k, v in inp.items()}
    gLs = torch.full((GROUP,), 1.0 / (GROUP * D * math.log(2.0)), device=dev)
    wss = ops.ElboWorkspace(ops.Desc(pdl=pdl), GROUP, dev)

    def small_step():
      wss.fwd_pre(sm['x'], sm['a'], sm['b'], sm['c'], sm['t'], sm['eps0'], sm['eps'])
      wss.post_bpd(sm['x'], sm['a'], sm['b'], sm['c'], sm['t'], sm['eps'], sm['net'], gLs)
      wss.bwd_pre(sm['x'], sm['a'], sm['b'], sm['c'], sm['t'], sm['eps'], sm['net'], sm['z_bar'],
                  sm['g_bar'], gLs)
    stream = torch.cuda.Stream()
    per_graph = 20
    with torch.cuda.stream(stream):
      for _ in range(3):
        small_step()
      torch.cuda.synchronize()
      g2 = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g2, stream=stream):
        for _ in range(per_graph):
          small_step()
      for _ in range(3):
        g2.replay()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(stream)
      for _ in range(20):
        g2.replay()
      e1.record(stream)
      torch.cuda.synchronize()
    out[label] = e0.elapsed_time(e1) * 1000 / (20 * per_graph)
  us = out['plain']
  return {'rows': GROUP, 'us_per_step': us, 'us_per_step_with_pdl': out['pdl'],
          'samples_per_s': GROUP / (us * 1e-6), 'launches_per_step': 3,
          'how': 'CUDA graph of 20 consecutive steps (3 plain launches each: fwd_pre, post '
                 'value-and-grad + loss scalars, bwd_pre; 768 threads per row), 20 replays, '
                 'L2-resident'}


def measure_e2e(ctx, args, inp, K):
  """The same loss + gradients through the host-buffer C-ABI call (mulan_elbo_host): H2D of every
  operand and D2H of every result inside the timed region, from / to page-locked buffers; next to
  it a COPY-ONLY leg moving exactly the same bytes with no kernel, i.e. what the platform's PCIe
  path allows (e2e is judged as a fraction of it)."""
  torch, dist = ctx.torch, ctx.dist
  from mulan_b200 import host, _lib
  rows, dev, world = args.rows, ctx.dev, ctx.world
  param = PARAMS[args.param]
  pin = lambda v: v.cpu().pin_memory()
  h = {k: pin(inp[k]) for k in ('x', 'a', 'b', 'c', 't', 'eps0', 'eps', 'net')}
  out = host.HostOutputs(rows, D, want_grad=True, pinned=True)
  call = lambda: host.elbo_host(h['x'], h['a'], h['b'], h['c'], h['t'], h['eps0'], h['eps'],
                                h['net'], param=param, want_grad=True, out=out)
  ke = max(2, min(K, 5))

  def timed(fn):
    fn(); fn()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(ke):
      r = fn()           # synchronises before returning
    return ctx.max_over_ranks(time.perf_counter() - t0), r
  te, r = timed(call)
  h2d = rows * D * (1 + 6 * 4) + rows * 4
  d2h = rows * D * 4 * 4 + (3 * rows + 6) * 4
  e2e = {'value': world * rows * ke / te, 'unit': 'samples/s', 'steps': ke,
         'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
         'h2d_gbs_per_gpu': h2d * ke / te / 1e9, 'd2h_gbs_per_gpu': d2h * ke / te / 1e9,
         'api': 'mulan_elbo_host (C ABI, pinned host buffers)', 'bpd': float(r['scalars'][0]),
         'numa': bind_to_gpu_numa.info}
  # the same call with eps_0 / eps drawn on the device from their threefry keys (what
  # VDM.__call__ itself does with its rng): 8 of the 25 H2D bytes per sub-pixel stay home
  callk = lambda: host.elbo_host(h['x'], h['a'], h['b'], h['c'], h['t'], None, None, h['net'],
                                 param=param, want_grad=True, out=out,
                                 jax_keys=((1234, ctx.rank), (5678, ctx.rank)))
  tk, rk = timed(callk)
  e2e['device_draws'] = {
      'value': world * rows * ke / tk, 'unit': 'samples/s',
      'h2d_bytes_per_step': rows * D * (1 + 4 * 4) + rows * 4 + 16,
      'd2h_bytes_per_step': d2h, 'api': 'mulan_elbo_host_keyed (eps_0, eps from JAX keys)',
      'bpd': float(rk['scalars'][0])}
  _lib.load().mulan_host_workspace_release()
  # ---- copy-only: the same pinned buffers, the same bytes, both directions at once, no kernel
  d_in = {k: torch.empty_like(inp[k]) for k in h}
  d_out = [torch.empty((rows, D), dtype=torch.float32, device=dev) for _ in range(4)]
  h_out = [torch.from_numpy(v) for v in (out.a_bar, out.b_bar, out.c_bar, out.n_bar)]
  s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

  def copy_only():
    with torch.cuda.stream(s_in):
      for k in h:
        d_in[k].copy_(h[k], non_blocking=True)
    with torch.cuda.stream(s_out):
      for src, dst in zip(d_out, h_out):
        dst.copy_(src, non_blocking=True)
    s_in.synchronize(); s_out.synchronize()

  def h2d_only():
    with torch.cuda.stream(s_in):
      for k in h:
        d_in[k].copy_(h[k], non_blocking=True)
    s_in.synchronize()
  tc, _ = timed(copy_only)
  th, _ = timed(h2d_only)
  copy_value = world * rows * ke / tc
  e2e['copy_only'] = {
      'value': copy_value, 'unit': 'samples/s',
      'h2d_gbs_per_gpu': h2d * ke / tc / 1e9, 'd2h_gbs_per_gpu': d2h * ke / tc / 1e9,
      'h2d_alone_gbs_per_gpu': h2d * ke / th / 1e9,
      'how': 'cudaMemcpyAsync of the same pinned buffers in both directions on two streams, no '
             'kernels: the PCIe / host-memory ceiling of this box for these bytes'}
  e2e['frac_of_copy_only'] = e2e['value'] / copy_value
  return e2e


def measure_dense(ctx, args, inp, K, W):
  """BASELINE.json configs[4]: dense-VLB forward (recon + prior + diffusion terms; the velocity
  model the CIFAR-10 config ships), images sharded over the ranks -- every rank evaluates
  rows / 128 images x 128 timesteps per step, one (sum, count) all-reduce per step.  Launch
  shapes: the reference-shaped 16 images x 128 timesteps = 2048 rows per launch on ONE stream,
  the same over 8 streams, and the driver's own sizing (dist.dense_images_per_launch: all 128
  images of the step in one launch) with eps_0 / eps broadcast instead of tiled."""
  torch, dist = ctx.torch, ctx.dist
  from mulan_b200 import ops, _lib
  from mulan_b200.dist import dense_images_per_launch
  dev, world = ctx.dev, ctx.world
  rows = args.rows
  param_name = 'vel'
  param = PARAMS[param_name]
  eps_form = _lib.kernel_param(param) == PARAMS['eps']
  out = {}
  T = GROUP
  noise0, noise = inp['eps0'][:T].contiguous(), inp['eps'][:T].contiguous()

  def build(lrows, n_streams, broadcast):
    desc = ops.Desc(param=param, pdl=not args.no_pdl, noise_rows=T if broadcast else 0)
    chunks = []
    for s0 in range(0, rows, lrows):
      ci = {k: v[s0:s0 + lrows] for k, v in inp.items()}
      chunks.append((ops.ElboWorkspace(desc, lrows, dev, save_w=eps_form), ci))
    side = [torch.cuda.Stream() for _ in range(min(len(chunks), n_streams) - 1)]

    def k_pre(w_, i):
      e0, e = (noise0, noise) if broadcast else (i['eps0'], i['eps'])
      w_.fwd_pre(i['x'], i['a'], i['b'], i['c'], i['t'], e0, e)

    def k_post(w_, i):
      e = noise if broadcast else i['eps']
      w_.post_bpd(i['x'], i['a'], i['b'], i['c'], i['t'], e, i['net'], None)

    def one(w_, i):
      k_pre(w_, i)
      k_post(w_, i)

    def step():
      cur = torch.cuda.current_stream()
      if not side:
        for ch in chunks:
          one(*ch)
        return
      fork = torch.cuda.Event()
      fork.record(cur)
      lanes = [cur] + side
      for s in side:
        s.wait_event(fork)
      for j, ch in enumerate(chunks):
        with torch.cuda.stream(lanes[j % len(lanes)]):
          one(*ch)
      for s in side:
        join = torch.cuda.Event()
        join.record(s)
        cur.wait_event(join)
    return step, chunks, (k_pre, k_post)

  acc = torch.zeros(2, dtype=torch.float64, device=dev)
  drv = min(dense_images_per_launch(T) * T, rows)
  shapes = [('ref_2048_rows_1_stream', min(args.launch_rows, rows), 1, False),
            ('ref_2048_rows_8_streams', min(args.launch_rows, rows), 8, False),
            ('driver_sized_tiled_noise_1_stream', drv, 1, False),
            ('driver_sized_1_stream', drv, 1, True)]
  for label, lrows, n_streams, broadcast in shapes:
    step, chunks, kfn = build(lrows, n_streams, broadcast)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
      for _ in range(max(W, 3)):
        step()
      torch.cuda.synchronize()
      graph = torch.cuda.CUDAGraph()
      with torch.cuda.graph(graph, stream=stream):
        step()
      graph.replay()
      ctx.barrier()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(stream)
      for _ in range(K):
        graph.replay()
      if world > 1:        # the ONE exchange of the evaluation: a final (sum, count) all-reduce
        acc[0] = chunks[0][0].scalars[0].double() * K  # (ldm/notebook_utils.py:191: np.mean(bpds))
        acc[1] = K * rows / T
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
      e1.record(stream)
      ctx.barrier()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1)) / K
    fwd_pre_b = (25 + (4 if eps_form else 0)) - (8.0 * (1 - T / lrows) if broadcast else 0.0)
    post_b = (12 if eps_form else 21) - (4.0 * (1 - T / lrows) if broadcast else 0.0)
    algo = (fwd_pre_b + post_b) * rows * D
    out[label] = {'rows_per_launch': lrows, 'streams': n_streams, 'broadcast_noise': broadcast,
                  'value': world * rows / (ms * 1e-3), 'unit': 'rows/s', 'ms_per_step': ms,
                  'images_per_s': world * rows / T / (ms * 1e-3),
                  'algo_bytes_per_subpixel': fwd_pre_b + post_b,
                  'gbs': algo / (ms * 1e-3) / 1e9,
                  'frac_of_measured': algo / (ms * 1e-3) / 1e9 / ctx.peak}
    if n_streams == 1:
      # each kernel alone, launch after launch on the one stream (CUDA events)
      for kname, fn, nb in (('fwd_pre', kfn[0], fwd_pre_b), ('fwd_post', kfn[1], post_b)):
        with torch.cuda.stream(stream):
          for ch in chunks:
            fn(*ch)
          torch.cuda.synchronize()
          a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          a0.record(stream)
          for _ in range(K):
            for ch in chunks:
              fn(*ch)
          a1.record(stream)
          torch.cuda.synchronize()
        kms = a0.elapsed_time(a1) / K
        out[label][kname] = {'ms': kms, 'algo_bytes_per_subpixel': nb,
                             'frac_of_measured': nb * rows * D / (kms * 1e-3) / 1e9 / ctx.peak}
    del graph, chunks
  best = max(out, key=lambda k: out[k]['value'])
  return {'workload': 'eval_bpd dense VLB forward (mulan_velocity: recon + prior + diffusion), '
                      f'{rows // T} images x {T} timesteps per rank per step, images sharded over '
                      f'{world} rank(s), denoiser output supplied',
          'value': out[best]['value'], 'unit': 'rows/s', 'best_shape': best, 'shapes': out}


def run_native(args):
  ctx = Ctx(args)
  torch = ctx.torch
  rows, K, W = args.rows, args.steps, args.warmup
  inp = make_inputs(rows, ctx.dev, seed=1234 + ctx.rank)
  if args.workload == 'dense_vlb':
    dense = measure_dense(ctx, args, inp, K, W)
    if ctx.rank == 0:
      emit({'metric': 'mulan_dense_vlb_rows_per_s', 'value': dense['value'], 'unit': 'rows/s',
            'n_gpus': ctx.world, 'steps': K, 'warmup': max(W, 3), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': dense['workload']}, 'dense_vlb': dense})
    ctx.close()
    return

  main = measure_elbo(ctx, args, inp, args.param, K, W, extras=True)
  e2e = measure_e2e(ctx, args, inp, K) if not args.no_e2e else None
  lat = measure_latency(ctx, args, inp) if ctx.rank == 0 else None

  configs = None
  if not args.no_configs:
    configs = {}
    kc = max(10, min(K, 30))
    for label, pn in (('cfg3_mulan_velocity', 'vel'),
                      ('cfg4_mulan_velocity_from_epsilon', 'vel_from_eps')):
      if pn == args.param:
        continue
      r = measure_elbo(ctx, args, inp, pn, kc, W, extras=False)
      configs[label] = {k: r[k] for k in ('value', 'unit', 'ms_per_step', 'steps', 'rows_per_gpu',
                                          'loss_form', 'saved_w', 'kernels', 'step_hbm', 'bpd')}
    configs['cfg5_dense_vlb'] = measure_dense(ctx, args, inp, kc, W)
  del inp
  torch.cuda.empty_cache()
  if configs is not None and not args.no_train_step:
    configs['train_step'] = measure_train_steps(ctx, args)

  # ---- cpu baseline (oracle port on host cores; rank 0, N=1 only) ----
  cpu = None
  if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu_baseline:
    cpu = cpu_reference(args, steps=0, warmup=1, min_seconds=8.0)
    cpu['single_thread'] = cpu_reference(args, steps=0, warmup=1, min_seconds=4.0, threads=1)
    cpu['config1_b8'] = cpu_reference(args, steps=0, warmup=1, min_seconds=3.0, rows=8)

  if ctx.rank != 0:
    ctx.close()
    return
  world = ctx.world
  dom = main['dominant']
  kinfo = main['kernels']
  ab = main['algo_bytes']
  nsub = rows * D
  traffic = load_traffic(dom, rows)
  line = {
      'metric': 'mulan_elbo_train_samples_per_s',
      'value': main['value'],
      'unit': 'samples/s', 'n_gpus': world, 'steps': K, 'warmup': max(W, 3),
      'ms_per_step': main['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': workload_name(args), 'rows_per_gpu': rows, 'dim': D,
                 'param': args.param, 'loss_form': main['loss_form'],
                 'saved_w': main['saved_w'], 'l2': 'inputs larger than L2 (%.2f GB of HBM traffic '
                 'per step)' % (main['step_hbm']['algo_bytes_per_step'] / 1e9),
                 'parallelism': f'dp{world} (rows sharded)',
                 'timed_loop': 'CUDA-graph replay of one step (%d launches, programmatic '
                               'dependent launch)' % main['launches_per_step'],
                 'scalar_pmean': main['scalar_pmean']},
      'roofline': {'bound': 'hbm', 'kernel': dom, 'achieved': kinfo[dom]['gbs'],
                   'peak': ctx.peak, 'peak_source': ctx.peak_src, 'unit': 'GB/s',
                   'frac': kinfo[dom]['gbs'] / ctx.peak, 'traffic': traffic,
                   'algo_bytes_per_launch': ab[dom] * nsub,
                   'how': 'CUDA events around %d back-to-back launches on the launch stream' % K},
      'step_hbm': main['step_hbm'],
      'kernels': kinfo, 'gpu_launches': main['launches_per_step'] * K, 'clocks': main['clocks'],
      'sustained': main['sustained'], 'e2e': e2e,
      'latency_b128': lat, 'cpu_baseline': cpu, 'bpd': main['bpd'], 'configs': configs,
  }
  emit(line)
  ctx.close()


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
def cpu_reference(args, steps, warmup, min_seconds=0.0, threads=None, rows=None):
  import torch
  from oracle import mulan_oracle as O   # CPU baseline leg: allowed to execute oracle/
  cores = threads or len(os.sched_getaffinity(0)) or os.cpu_count() or 1
  torch.set_num_threads(cores)
  B = rows or args.ref_rows
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
          'rows': B,
          'sample': f'{steps} steps of {B} rows ({"one per-GPU batch of the workload" if B == GROUP else "BASELINE.json configs[0]" if B == 8 else "sample"}), oracle '
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


def train_step_leg(ctx, args, net_config, param_name, B, modes, K, W):
  """Whole train step around the kernels (Experiment.train_step, ldm/experiment.py:335-356):
  STAND-IN encoder and U-Net on cuDNN/cuBLAS float32 (mulan_b200/standin.py -- the reference
  keeps them on the framework path; same layer structure and parameter count), the ELBO kernels,
  the gradient exchange and the fused AdamW+EMA update, for each way of running the exchange
  (mulan_b200/optim.py: 'allreduce' after backward, 'overlap' = bucketed NCCL during backward,
  'peer' = mulan_adamw_ema_peer over NVLink peer memory during backward)."""
  torch, dist = ctx.torch, ctx.dist
  from mulan_b200.model import VDM, VDMConfig
  from mulan_b200.optim import FlatTrainState, train_step
  from mulan_b200.standin import ScoreUNet, UnetEncoder
  dev, world, rank = ctx.dev, ctx.world, ctx.rank
  torch.backends.cuda.matmul.allow_tf32 = args.tf32      # reference: matmul precision float32
  torch.backends.cudnn.allow_tf32 = args.tf32
  torch.backends.cudnn.benchmark = True
  n_embd = 128 if net_config == 'cifar10' else 256
  cfg = VDMConfig(vdm_type='mulan_epsilon' if param_name == 'eps' else 'mulan_velocity',
                  velocity_from_epsilon=(param_name == 'vel_from_eps'))
  host_images = torch.randint(0, 256, (B, 32, 32, 3), dtype=torch.uint8).pin_memory()
  res = {'workload': f'{net_config} {cfg.vdm_type}'
                     f'{"(velocity_from_epsilon)" if cfg.velocity_from_epsilon else ""} train '
                     f'step, per-GPU batch {B}, global {B * world}; STAND-IN encoder / U-Net on '
                     f'cuDNN/cuBLAS float32, ELBO kernels + gradient exchange + fused AdamW+EMA in '
                     f'libmulan_b200', 'sm_n_embd': n_embd, 'per_gpu_batch': B, 'modes': {}}
  for mode in modes:
    torch.manual_seed(1234)                                 # same init on every rank
    model = VDM(cfg, UnetEncoder(n_embd, 4), ScoreUNet(n_embd, 32)).to(dev)
    model.train()
    try:
      state = FlatTrainState(model.named_parameters(), comm=mode)
    except RuntimeError as exc:      # peer memory unavailable: raised on every rank together
      res['modes'][mode] = {'unavailable': str(exc)[:200]}
      del model
      continue
    gen = torch.Generator(device=dev).manual_seed(100 + rank)

    def one_step():
      images = host_images.to(dev, non_blocking=True)       # H2D of the step's inputs
      return train_step(model, state, {'images': images}, generator=gen)
    for _ in range(W):
      sc = one_step()
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(K):
      sc = one_step()
      bpd = float(sc['bpd'])                                # D2H read of the step's result
    e1.record()
    ctx.barrier()
    wall = time.perf_counter() - t0
    ms = ctx.max_over_ranks(e0.elapsed_time(e1)) / K
    wall_ms = ctx.max_over_ranks(wall * 1e3) / K
    m = {'ms_per_step': ms, 'value': world * B / (ms * 1e-3), 'unit': 'samples/s',
         'e2e_value': world * B / (wall_ms * 1e-3), 'bpd': bpd, 'steps': K,
         'exchange_ranges': len(state.ranges)}
    # ---- the pieces alone (CUDA events, 5 repetitions each)
    def alone(fn, reps=5):
      fn()
      ctx.barrier()
      a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a0.record()
      for _ in range(reps):
        fn()
      a1.record()
      ctx.barrier()
      return ctx.max_over_ranks(a0.elapsed_time(a1)) / reps
    nbytes = 4 * state.n
    if mode == 'allreduce':
      if world > 1:
        t_ar = alone(lambda: dist.all_reduce(state.grads, op=dist.ReduceOp.SUM))
        m['nccl_allreduce_ms'] = t_ar
        m['nccl_allreduce_busbw_gbs'] = 2 * (world - 1) / world * nbytes / (t_ar * 1e-3) / 1e9
      t_up = alone(lambda: state.apply_gradients())
      m['adamw_ema_ms'] = t_up
      m['adamw_ema_frac_of_measured'] = 36.0 * state.n / (t_up * 1e-3) / 1e9 / ctx.peak
    elif mode == 'peer':
      def fused():
        state._reset_ranges()
        for i in range(len(state.ranges)):
          state._fire(i)
        state.finish_exchange()
      t_f = alone(fused)
      m['fused_exchange_update_ms'] = t_f

      def fused_whole():            # the whole bucket in ONE call (as the all-reduce leg does)
        state._reset_ranges()
        state.peer_update_range(0, state.n)
      t_w = alone(fused_whole)
      m['fused_exchange_update_single_call_ms'] = t_w
      m['nvlink_gbs_per_direction_single_call'] = (2 * (world - 1) / world * nbytes
                                                   / (t_w * 1e-3) / 1e9)
      # NVLink bytes per rank and DIRECTION: in = (world-1)/world of the bucket read from peers +
      # as much of new parameters stored here by the peers; out = the mirror image
      m['nvlink_gbs_per_direction'] = 2 * (world - 1) / world * nbytes / (t_f * 1e-3) / 1e9
      m['timed_out'] = state.peer.timed_out()
      m['in_switch_reduction'] = state.peer.multicast   # multimem.ld_reduce / multimem.st
    res['modes'][mode] = m
    res['parameters'] = state.n
    res['grad_bucket_bytes'] = nbytes
    if state.peer is not None:
      state.peer.close()
    for h_ in state._hooks:
      h_.remove()
    del model, state
    torch.cuda.empty_cache()
  ran = {k: v for k, v in res['modes'].items() if 'ms_per_step' in v}
  best = min(ran, key=lambda k: ran[k]['ms_per_step'])
  res.update(value=res['modes'][best]['value'], unit='samples/s', best_mode=best,
             ms_per_step=res['modes'][best]['ms_per_step'])
  return res


def measure_train_steps(ctx, args):
  """The train-step legs of the bench line: configs[1]/[2] (CIFAR-10 networks, the velocity model
  the config ships, per-GPU batch 128: global 1024 at 8 GPUs, 285 MB gradient bucket) at every N,
  and configs[3] (ImageNet-32 networks, v-from-eps, 256 per GPU = global 2048, 682 MB) at N = 8."""
  modes = ['allreduce'] if ctx.world == 1 else ['allreduce', 'overlap', 'peer']
  K = max(2, min(args.train_steps, 10))
  out = {'cfg2_cfg3_cifar10_velocity': train_step_leg(ctx, args, 'cifar10', 'vel', 128, modes,
                                                      K, 2)}
  if ctx.world == 8 and not args.no_imagenet:
    out['cfg4_imagenet32_vfe'] = train_step_leg(ctx, args, 'imagenet32', 'vel_from_eps', 256,
                                                modes, 2, 1)
  return out


def run_train_step(args):
  """--workload train_step: one train-step leg on its own (see train_step_leg)."""
  ctx = Ctx(args)
  B = args.global_batch // ctx.world if args.global_batch else args.batch
  modes = [args.comm] if args.comm != 'all' else (
      ['allreduce'] if ctx.world == 1 else ['allreduce', 'overlap', 'peer'])
  K, W = args.steps, max(args.warmup, 2)
  leg = train_step_leg(ctx, args, args.net_config, args.param, B, modes, K, W)
  if ctx.rank == 0:
    emit({'metric': 'mulan_train_step_samples_per_s', 'value': leg['value'], 'unit': 'samples/s',
          'n_gpus': ctx.world, 'steps': K, 'warmup': W, 'ms_per_step': leg['ms_per_step'],
          'higher_is_better': True, 'scaling': 'strong' if args.global_batch else 'weak',
          'vs_baseline': None, 'dtype': 'tf32' if args.tf32 else 'f32', 'data': 'synthetic',
          'config': {'workload': leg['workload'], 'parallelism': f'dp{ctx.world}'},
          'train_step': leg,
          'e2e': {'value': leg['modes'][leg['best_mode']]['e2e_value'], 'unit': 'samples/s',
                  'h2d_bytes_per_step': B * D, 'd2h_bytes_per_step': 4,
                  'api': 'optim.train_step(model, state, batch) with host uint8 images'}})
  ctx.close()


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
