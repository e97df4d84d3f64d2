"""Fused AdamW + EMA update (mulan_adamw_ema) against the CPU oracle of the optax chain, and the
oracle against torch.optim.AdamW (independent implementation of the same published formulas)."""
import math

import numpy as np
import pytest
import torch

from oracle import adamw_oracle as AO


def test_oracle_matches_torch_adamw():
  """optax.adamw and torch.optim.AdamW differ only in where lr multiplies the decay term
  (torch: p *= 1 - lr*wd before the Adam step; optax: p -= lr*(u + wd*p)) -- identical to
  O(lr^2 wd). Check the oracle against torch in float64 on a few steps."""
  torch.manual_seed(0)
  p0 = torch.randn(1000, dtype=torch.float64)
  p_t = p0.clone().requires_grad_(True)
  opt = torch.optim.AdamW([p_t], lr=2e-4, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.01)
  p, mu, nu, ema = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0), p0.clone()
  for step in range(1, 6):
    g = torch.randn(1000, dtype=torch.float64)
    p_t.grad = g.clone()
    opt.step()
    p, mu, nu, ema = AO.adamw_ema_step(p, g, mu, nu, ema, step, 2e-4)
    assert (p - p_t.detach()).abs().max().item() < 1e-9       # lr^2 * wd * |p| ~ 4e-10 per step
  assert (ema - p0).abs().max().item() < 1e-3 and not torch.equal(ema, p0)


def test_lr_schedule_and_mask():
  from mulan_b200.optim import decay_mask, lr_schedule
  assert lr_schedule(0) == 0.0 and abs(lr_schedule(50) - 1e-4) < 1e-12
  assert lr_schedule(100) == 2e-4 and lr_schedule(10**6) == 2e-4
  assert decay_mask('score.conv_in.weight') and decay_mask('score.norm_out.weight')
  assert not decay_mask('score.conv_in.bias')
  # lr_decay=True: join_schedules([warm-up, linear decay to 0], boundaries=[warm-up])
  for step in (0, 1, 50, 99, 100, 101, 500, 999, 1000, 5000):
    for decay in (False, True):
      want = AO.lr_schedule(step, 2e-4, 100, decay, 1000)
      got = lr_schedule(step, 2e-4, 100, decay, 1000)
      assert abs(got - want) < 1e-18, (step, decay)
  assert lr_schedule(550, 2e-4, 100, True, 1000) == pytest.approx(1e-4)
  assert lr_schedule(1000, 2e-4, 100, True, 1000) == 0.0


def test_oracle_clip_by_global_norm():
  g = [torch.tensor([3.0, 0.0]), torch.tensor([[4.0]])]
  out, norm = AO.clip_by_global_norm(g, 10.0)
  assert norm.item() == 5.0 and all(torch.equal(a, b) for a, b in zip(out, g))
  out, _ = AO.clip_by_global_norm(g, 1.0)
  assert torch.allclose(out[0], torch.tensor([0.6, 0.0])) and torch.allclose(out[1], torch.tensor([[0.8]]))


@pytest.mark.gpu
@pytest.mark.parametrize('clip', [0.5, 1e6])
def test_adamw_ema_clip_by_global_norm(cuda_device, clip):
  """gradient_clip_norm: mulan_grad_sumsq + the clipped update == the oracle's
  clip_by_global_norm -> adamw -> ema on the same bucket (clip active / inactive)."""
  import ctypes as C
  from mulan_b200 import _lib
  dev = cuda_device
  rng = np.random.default_rng(1)
  n = 1_000_004
  p = torch.from_numpy(rng.standard_normal(n).astype(np.float32))
  g = torch.from_numpy((rng.standard_normal(n) * 1e-2).astype(np.float32))
  mu, nu, ema = torch.zeros(n), torch.zeros(n), p.clone()
  gp, gg, gmu, gnu, gema = (t.clone().to(dev) for t in (p, g, mu, nu, ema))
  ptr = lambda t: C.c_void_p(t.data_ptr())
  sumsq = torch.zeros(1, device=dev)
  scratch = torch.empty(_lib.MULAN_SUMSQ_SCRATCH, dtype=torch.float64, device=dev)
  lib = _lib.load()
  _lib.check(lib.mulan_grad_sumsq(n, ptr(gg), ptr(scratch), ptr(sumsq), None))
  want_ss = (g.double() ** 2).sum().item()
  assert abs(sumsq.item() - want_ss) < 1e-6 * want_ss
  first = sumsq.item()
  _lib.check(lib.mulan_grad_sumsq(n, ptr(gg), ptr(scratch), ptr(sumsq), None))
  assert sumsq.item() == first                                      # deterministic
  d = _lib.MulanAdamwDesc(n, n, 3, 0, 2e-4, 0.9, 0.99, 1e-8, 0.01, 0.9999, 0.5, clip,
                          sumsq.data_ptr())
  _lib.check(lib.mulan_adamw_ema(C.byref(d), ptr(gp), ptr(gg), ptr(gmu), ptr(gnu), ptr(gema), None))
  torch.cuda.synchronize()
  wp, wmu, wnu, wema = AO.adamw_ema_step(p, g, mu, nu, ema, 3, 2e-4, grad_scale=0.5,
                                         clip_norm=clip)
  clipped = 0.5 * math.sqrt(want_ss) >= clip
  assert clipped == (clip == 0.5)
  for got, want in ((gp, wp), (gmu, wmu), (gnu, wnu), (gema, wema)):
    tol = 2e-6 * want.abs() + 2e-7 * want.abs().max()
    assert torch.all((got.cpu() - want).abs() <= tol)
  # clip_norm without the device scalar is refused
  bad = _lib.MulanAdamwDesc(n, n, 3, 0, 2e-4, 0.9, 0.99, 1e-8, 0.01, 0.9999, 0.5, clip, None)
  assert lib.mulan_adamw_ema(C.byref(bad), ptr(gp), ptr(gg), ptr(gmu), ptr(gnu), ptr(gema),
                             None) == -1


@pytest.mark.gpu
def test_adamw_ema_kernel_parity(cuda_device):
  import ctypes as C
  from mulan_b200 import _lib
  dev = cuda_device
  rng = np.random.default_rng(0)
  n, n_decay = 40_000, 25_000
  n_decay = n_decay // 4 * 4
  p = torch.from_numpy(rng.standard_normal(n).astype(np.float32))
  mu, nu, ema = torch.zeros(n), torch.zeros(n), p.clone()
  mask = torch.arange(n) < n_decay
  gp, gmu, gnu, gema = (t.clone().to(dev) for t in (p, mu, nu, ema))
  p64, mu64, nu64, ema64 = p.double(), mu.double(), nu.double(), ema.double()
  ptr = lambda t: C.c_void_p(t.data_ptr())
  for step in range(1, 8):
    g = torch.from_numpy((rng.standard_normal(n) * 10 ** rng.uniform(-4, 0, n)).astype(np.float32))
    lr = 2e-4 * min(step, 100) / 100
    p, mu, nu, ema = AO.adamw_ema_step(p, g, mu, nu, ema, step, lr, decay_mask=mask, grad_scale=0.5)
    p64, mu64, nu64, ema64 = AO.adamw_ema_step(p64, g.double(), mu64, nu64, ema64, step, lr,
                                               decay_mask=mask, grad_scale=0.5)
    d = _lib.MulanAdamwDesc(n, n_decay, step, 0, lr, 0.9, 0.99, 1e-8, 0.01, 0.9999, 0.5)
    gg = g.to(dev)
    _lib.check(_lib.load().mulan_adamw_ema(C.byref(d), ptr(gp), ptr(gg), ptr(gmu), ptr(gnu),
                                           ptr(gema), None))
    torch.cuda.synchronize()
  for got, want, w64 in ((gp, p, p64), (gmu, mu, mu64), (gnu, nu, nu64), (gema, ema, ema64)):
    got = got.cpu()
    # same op order; only 1 - b^t (host powf vs torch pow) can differ by an ulp, and sums of
    # opposite-sign terms cancel -> compare against the tensor's scale, not element-wise only
    tol = 2e-6 * want.abs() + 2e-7 * want.abs().max()
    assert torch.all((got - want).abs() <= tol), ((got - want).abs() / tol).max()
    err = (got.double() - w64).abs().max().item()
    ref = (want.double() - w64).abs().max().item()
    assert err <= 4 * ref + 1e-9
  # argument validation
  bad = _lib.MulanAdamwDesc(n + 1, n_decay, 1, 0, 1e-4, 0.9, 0.99, 1e-8, 0.01, 0.9999, 1.0)
  assert _lib.load().mulan_adamw_ema(C.byref(bad), ptr(gp), ptr(gg), ptr(gmu), ptr(gnu),
                                     ptr(gema), None) == -2


@pytest.mark.gpu
def test_flat_train_state_step(cuda_device):
  """FlatTrainState re-homes a module's parameters / gradients into flat buffers; one
  train_step updates them exactly like the oracle applied tensor by tensor."""
  from mulan_b200.optim import FlatTrainState, decay_mask
  dev = cuda_device
  torch.manual_seed(0)
  net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.GroupNorm(1, 5),
                            torch.nn.Linear(5, 3)).to(dev)
  before = {n: p.detach().clone().cpu() for n, p in net.named_parameters()}
  state = FlatTrainState(net.named_parameters())
  assert state.n % 4 == 0 and state.n_decay % 4 == 0
  for n, p in net.named_parameters():
    assert torch.equal(p.detach().cpu(), before[n])               # values survived the move
  x = torch.randn(4, 7, device=dev)
  state.zero_grad()
  net(x).square().mean().backward()
  grads = {n: p.grad.detach().clone().cpu() for n, p in net.named_parameters()}
  lr = state.apply_gradients()
  assert lr == 0.0                                                # first step of the warm-up
  state.zero_grad()
  net(x).square().mean().backward()
  grads2 = {n: p.grad.detach().clone().cpu() for n, p in net.named_parameters()}
  lr2 = state.apply_gradients()
  assert abs(lr2 - 2e-6) < 1e-12
  ema = state.ema_state_dict()
  for n, p in net.named_parameters():
    q, mu, nu, e = before[n], torch.zeros_like(before[n]), torch.zeros_like(before[n]), before[n]
    m = torch.full_like(q, decay_mask(n), dtype=torch.bool)
    q, mu, nu, e = AO.adamw_ema_step(q, grads[n], mu, nu, e, 1, 0.0, decay_mask=m)
    q, mu, nu, e = AO.adamw_ema_step(q, grads2[n], mu, nu, e, 2, 2e-6, decay_mask=m)
    assert torch.allclose(p.detach().cpu(), q, rtol=1e-6, atol=1e-9), n
    assert torch.allclose(ema[n].cpu().view_as(e), e, rtol=1e-6, atol=1e-9), n


@pytest.mark.gpu
def test_flat_train_state_gradient_clip(cuda_device):
  """config.gradient_clip_norm: the clip uses the GLOBAL norm over every parameter tensor."""
  from mulan_b200.optim import FlatTrainState, decay_mask
  dev = cuda_device
  torch.manual_seed(1)
  net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3)).to(dev)
  before = {n: p.detach().clone().cpu() for n, p in net.named_parameters()}
  state = FlatTrainState(net.named_parameters(), num_steps_lr_warmup=0, gradient_clip_norm=0.05)
  x = torch.randn(4, 7, device=dev)
  state.zero_grad()
  (10 * net(x).square().mean()).backward()
  grads = {n: p.grad.detach().clone().cpu() for n, p in net.named_parameters()}
  names = list(grads)
  clipped, norm = AO.clip_by_global_norm([grads[n] for n in names], 0.05)
  assert norm.item() > 0.05
  assert abs(state.grad_global_norm().item() - norm.item()) < 1e-6 * norm.item()
  lr = state.apply_gradients()
  assert lr == 2e-4
  for n, gc in zip(names, clipped):
    q = before[n]
    m = torch.full_like(q, decay_mask(n), dtype=torch.bool)
    q, _, _, _ = AO.adamw_ema_step(q, gc, torch.zeros_like(q), torch.zeros_like(q), q, 1, 2e-4,
                                   decay_mask=m)
    got = dict(net.named_parameters())[n].detach().cpu()
    assert torch.allclose(got, q, rtol=1e-6, atol=1e-8), n
