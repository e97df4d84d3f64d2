"""Fused AdamW + EMA update (mulan_adamw_ema) against the CPU oracle of the optax chain, and the
oracle against torch.optim.AdamW (independent implementation of the same published formulas)."""
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
