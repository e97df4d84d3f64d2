"""The gradient exchange fused with the optimizer: host-side logic on CPU
(gloo, world_size 2) and the NVLink peer-memory kernel on >= 2 GPUs.

Reference statements: grads = jax.lax.pmean(grads, 'batch') (ldm/experiment.py:341) then
state.apply_gradients (ldm/experiment.py:344 -> ldm/train_state.py:70-102)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HP = dict(lr=2e-4, b1=0.9, b2=0.99, eps=1e-8, weight_decay=0.01, ema_rate=0.9999)


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def test_shard_ranges_tile_the_range():
  from mulan_b200.optim import shard_range
  for lo, hi in ((0, 0), (0, 4), (8, 40), (0, 100), (16, 71153852 // 4 * 4)):
    for world in (1, 2, 4, 8):
      cuts = [shard_range(lo, hi, world, r) for r in range(world)]
      assert cuts[0][0] == lo and cuts[-1][1] == hi
      for (a0, a1), (b0, b1) in zip(cuts, cuts[1:]):
        assert a1 == b0 and a0 <= a1
      assert all(a % 4 == 0 and b % 4 == 0 for a, b in cuts)


def test_plan_buckets_covers_the_buffer_on_parameter_boundaries():
  from mulan_b200.optim import plan_buckets
  layout, off = [], 0
  rng = np.random.default_rng(0)
  for i in range(40):
    k = int(rng.integers(1, 5000))
    layout.append((f'p{i}', off, k))
    off += (k + 3) // 4 * 4
  ranges, members = plan_buckets(layout, off, 8192)
  assert ranges[0][0] == 0 and ranges[-1][1] == off
  for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
    assert a1 == b0
  starts = {o for _, o, _ in layout} | {off}
  assert all(a in starts and b in starts for a, b in ranges)
  assert sorted(n for m in members for n in m) == sorted(n for n, _, _ in layout)
  assert all(b - a >= 8192 for a, b in ranges[:-1])


def _gloo_worker(rank, world, port, q):
  """Sharded update on CPU: reduce each range's shard in rank order, update it with the oracle,
  all-gather the new parameters -- against all-reduce + full-size update."""
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                    WORLD_SIZE=str(world))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from mulan_b200.optim import plan_buckets, shard_range
  from oracle import adamw_oracle as AO
  torch.manual_seed(0)
  n, n_decay = 4096, 3000
  p0, mu0, nu0 = torch.randn(n), 0.01 * torch.randn(n), 0.001 * torch.rand(n)
  ema0 = p0.clone()
  g = torch.Generator().manual_seed(100 + rank)
  grad = torch.randn(n, generator=g)
  mask = torch.arange(n) < n_decay
  # reference: all-reduce (sum in rank order), full update with grad_scale = 1/world
  everyone = [torch.empty(n) for _ in range(world)]
  dist.all_gather(everyone, grad)
  total = everyone[0].clone()
  for r in range(1, world):
    total = total + everyone[r]
  want = AO.adamw_ema_step(p0, total, mu0, nu0, ema0, 3, decay_mask=mask, grad_scale=1.0 / world,
                           **HP)
  # sharded: every range, my shard only
  layout = [(f'p{i}', 512 * i, 512) for i in range(8)]
  ranges, _ = plan_buckets(layout, n, 1000)
  p, mu, nu, ema = p0.clone(), mu0.clone(), nu0.clone(), ema0.clone()
  for lo, hi in ranges:
    a, b = shard_range(lo, hi, world, rank)
    red = everyone[0][a:b].clone()
    for r in range(1, world):
      red = red + everyone[r][a:b]
    out = AO.adamw_ema_step(p[a:b], red, mu[a:b], nu[a:b], ema[a:b], 3, decay_mask=mask[a:b],
                            grad_scale=1.0 / world, **HP)
    p[a:b], mu[a:b], nu[a:b], ema[a:b] = out
    # all-gather of the new parameters: every rank broadcasts its shard
    for r in range(world):
      ra, rb = shard_range(lo, hi, world, r)
      piece = p[ra:rb].clone() if r == rank else torch.empty(rb - ra)
      dist.broadcast(piece, src=r)
      p[ra:rb] = piece
  ok_p = torch.equal(p, want[0])
  # optimizer state is sharded: only my shards hold the new moments
  ok_state = all(torch.equal(mu[a:b], want[1][a:b]) and torch.equal(nu[a:b], want[2][a:b])
                 and torch.equal(ema[a:b], want[3][a:b])
                 for a, b in (shard_range(lo, hi, world, rank) for lo, hi in ranges))
  q.put((rank, ok_p, ok_state, len(ranges)))
  dist.destroy_process_group()


def test_sharded_update_equals_allreduce_plus_full_update_gloo():
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=300) for _ in procs]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  for rank, ok_p, ok_state, n_ranges in res:
    assert ok_p and ok_state and n_ranges >= 2, (rank, ok_p, ok_state, n_ranges)


# ------------------------------------------------------------------------------------------
# GPU: mulan_adamw_ema_peer over CUDA-IPC peer memory, one process per GPU
# ------------------------------------------------------------------------------------------

def _gpu_worker(rank, world, port, q):
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                    WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
  torch.cuda.set_device(rank)
  dev = torch.device(f'cuda:{rank}')
  dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
  import ctypes as C
  from mulan_b200 import _lib
  from mulan_b200.optim import FlatTrainState
  out = {}
  try:
    def make_model():
      torch.manual_seed(7)
      return torch.nn.Sequential(torch.nn.Linear(300, 517), torch.nn.SiLU(),
                                 torch.nn.Linear(517, 1023), torch.nn.SiLU(),
                                 torch.nn.Linear(1023, 64)).to(dev)
    results = {}
    for comm in ('allreduce', 'overlap', 'peer_mc', 'peer'):
      model = make_model()
      st = FlatTrainState(model.named_parameters(), comm=comm.split('_')[0], bucket_mb=0.5,
                          num_steps_lr_warmup=2, multicast=(comm == 'peer_mc'))   # not 'auto'
      if comm == 'peer_mc':
        out['multicast'] = st.peer.multicast       # False where the fabric has no multicast
      gen = torch.Generator(device=dev).manual_seed(50 + rank)
      snaps = []
      for step in range(3):
        st.zero_grad()
        xb = torch.randn((32, 300), generator=gen, device=dev)
        loss = model(xb).square().mean() * (1.0 + 0.1 * rank)
        loss.backward()
        st.tail[:1] = loss.detach()
        st.all_reduce()
        st.apply_gradients()
        torch.cuda.synchronize()
        snaps.append(float(st.tail[0]))
      st.gather_sharded_state()     # 'peer': mu / nu / ema shards from their owners
      torch.cuda.synchronize()
      results[comm] = [t.clone() for t in (st.params, st.ema)] + [snaps, st.step, len(st.ranges)]
      results[comm + '_moments'] = [st.mu.clone(), st.nu.clone()]
      if comm.startswith('peer'):
        assert not st.peer.timed_out()
      results[comm + '_ranges'] = list(st.ranges)
      if comm == 'peer':
        assert not st.peer.multicast
        peer_state = st
      elif st.peer is not None:
        for h_ in st._hooks:
          h_.remove()
        st.peer.close()
    for comm in ('overlap', 'peer_mc', 'peer'):
      for a_, b_ in zip(results[comm][:2], results['allreduce'][:2]):
        err = ((a_ - b_).abs().max() / b_.abs().max()).item()
        assert err < 2e-6, (comm, err)
      for a_, b_ in zip(results[comm + '_moments'], results['allreduce_moments']):
        err = ((a_ - b_).abs().max() / b_.abs().max()).item()
        assert err < 1e-5, (comm, 'moments', err)       # every shard, after the gather
      assert results[comm][3] == results['allreduce'][3] == 3
      assert np.allclose(results[comm][2], results['allreduce'][2], rtol=1e-6)
    # parameters identical on every rank after the fused all-gather
    mine = results['peer'][0]
    other = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(other, mine)
    for o in other:
      assert torch.equal(o, mine)
    # ---- bit-exactness of ONE fused update against its single-process emulation: sum the ranks'
    #      gradients in rank order, then the plain full-size kernel.  mulan_adamw_ema_peer must
    #      reproduce the parameters (on every rank) and this rank's shards of mu / nu / ema.
    st = peer_state
    lib = _lib.load()
    n = st.n
    gen = torch.Generator(device=dev).manual_seed(900 + rank)
    st.grads[:n] = torch.randn(n, generator=gen, device=dev)
    torch.cuda.synchronize()
    dist.barrier()
    everyone = [torch.empty(n, device=dev) for _ in range(world)]
    dist.all_gather(everyone, st.grads[:n].clone())
    total = everyone[0].clone()
    for r in range(1, world):
      total = total + everyone[r]
    ref = [t.clone() for t in (st.params, st.mu, st.nu, st.ema)]
    ptr = lambda t: C.c_void_p(t.data_ptr())
    d = _lib.MulanAdamwDesc(n, st.n_decay, 9, 0, 1.5e-4, 0.9, 0.99, 1e-8, 0.01, 0.9999, 1.0 / world,
                            0.0, None)
    cur = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.mulan_adamw_ema(C.byref(d), ptr(ref[0]), ptr(total), ptr(ref[1]), ptr(ref[2]),
                                   ptr(ref[3]), cur))
    from mulan_b200.optim import shard_range
    for lo, hi in st.ranges:
      pd = st.peer.desc()
      _lib.check(lib.mulan_adamw_ema_peer(C.byref(d), C.byref(pd), lo, hi, ptr(st.mu), ptr(st.nu),
                                          ptr(st.ema), cur))
    torch.cuda.synchronize()
    assert not st.peer.timed_out()
    # the optimizer state is SHARDED: rank r's emulation is right on r's shards only (elsewhere
    # its mu / nu are stale), so the expected parameters are assembled from the owners
    refs = [torch.empty_like(ref[0]) for _ in range(world)]
    dist.all_gather(refs, ref[0])
    expect = torch.empty_like(ref[0])
    for lo, hi in st.ranges:
      for r in range(world):
        a, b = shard_range(lo, hi, world, r)
        expect[a:b] = refs[r][a:b]
    assert torch.equal(st.params, expect)
    for lo, hi in st.ranges:
      a, b = shard_range(lo, hi, world, rank)
      for got, want in ((st.mu, ref[1]), (st.nu, ref[2]), (st.ema, ref[3])):
        assert torch.equal(got[a:b], want[a:b])
    dist.barrier()
    st.peer.close()
    # ---- the six loss scalars' pmean through the peer-memory board (mulan_post_bpd_peer)
    import math
    from mulan_b200 import ops
    from mulan_b200.peer import ScalarBoard
    from oracle import mulan_oracle as O
    rows = 300
    inp = O.synth_inputs(rows, 40 + rank)
    g = {k: v.to(dev).contiguous() for k, v in inp.items()}
    ws = ops.ElboWorkspace(ops.Desc(), rows, dev)
    gL = torch.full((rows,), 1.0 / (rows * 3072 * math.log(2.0)), device=dev)
    board = ScalarBoard(dev)
    for rep in range(70):            # more steps than the ring has slots
      ws.fwd_pre(g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'], g['eps'])
      ws.post_bpd(g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'], gL, board=board)
      if rep % 23 == 0:
        dist.barrier()               # keep the ranks within the ring
    mean, step = board.read()
    torch.cuda.synchronize()
    assert int(step.item()) == 70
    everyone = [torch.empty(6, device=dev) for _ in range(world)]
    dist.all_gather(everyone, ws.scalars.clone())
    want = everyone[0].clone()
    for r in range(1, world):
      want = want + everyone[r]
    want = want / world
    assert torch.equal(mean, want), (mean, want)
    dist.barrier()
    board.close()
    out['ok'] = True
  except Exception as exc:      # reported to the parent
    import traceback
    out['ok'] = False
    out['err'] = traceback.format_exc()
  q.put((rank, out))
  dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize('world', [2, 4, 8])
def test_peer_fused_update_matches_allreduce_and_is_bit_exact(cuda_device, world):
  if torch.cuda.device_count() < world:
    pytest.skip(f'needs {world} GPUs')
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_gpu_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = [q.get(timeout=600) for _ in procs]
  for p in procs:
    p.join(timeout=120)
  for rank, out in res:
    assert out.get('ok'), (rank, out.get('err'))
