"""Probability-flow ODE driver: the device-resident RK45 integrator, the
exact-likelihood function and the ODE sampler (ldm/notebook_utils.py:264-448).

The integrator the reference calls is scipy.integrate.solve_ivp(method='RK45'); scipy is
installed, so it is the pin: the oracle restatement must reproduce it bit for bit, and the
product's host control flow / CUDA state must reproduce the same step sequence.
"""
import math
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from scipy.integrate import RK45, solve_ivp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import golden_inputs as GI  # noqa: E402
from oracle import mulan_oracle as O  # noqa: E402
from oracle import rk45_oracle as R  # noqa: E402


def host_fun(w):
  """A smooth non-autonomous linear test system whose derivative is rounded to float32, like
  the reference's device round trip (_to_flattened_numpy of a float32 drift)."""
  def fun(t, y):
    y32 = np.asarray(y, dtype=np.float32)
    return (np.sin(np.float32(5 * t) + w) * y32 + np.float32(np.cos(3 * t)) * w).astype(
        np.float32).astype(np.float64)
  return fun


def problem(n=501, seed=0):
  rng = np.random.default_rng(seed)
  return rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)


def test_tableau_is_scipys():
  from mulan_b200 import ode
  for mine in (R, None):
    A = R.A if mine else np.array([list(r) + [0.0] * (5 - len(r)) for r in ode.RK45_A])
    B = R.B if mine else np.array(ode.RK45_B)
    Cc = R.C if mine else np.array(ode.RK45_C)
    E = R.E if mine else np.array(ode.RK45_E)
    assert np.array_equal(A, RK45.A) and np.array_equal(B, RK45.B)
    assert np.array_equal(Cc, RK45.C) and np.array_equal(E, RK45.E)
  assert (ode.SAFETY, ode.MIN_FACTOR, ode.MAX_FACTOR) == (0.9, 0.2, 10.0)


@pytest.mark.parametrize('span', [(0.0, 1.0), (1.0, 0.0)])
@pytest.mark.parametrize('tol', [1e-3, 1e-5, 1e-8])
def test_oracle_rk45_is_scipy_bit_for_bit(span, tol):
  w, y0 = problem()
  fun = host_fun(w)
  s = solve_ivp(fun, span, y0, rtol=tol, atol=tol, method='RK45')
  r = R.solve_rk45(fun, span, y0, rtol=tol, atol=tol)
  assert r.status == 0 and s.status == 0
  assert r.nfev == s.nfev and r.n_steps == len(s.t) - 1
  assert np.array_equal(np.array(r.ts), s.t)
  assert np.array_equal(r.y, s.y[:, -1])


class CpuState:
  """Stand-in for mulan_b200.ode.DeviceState on torch-CPU float64: lets the host control flow of
  solve_ivp_rk45 run without a GPU.  Lives in tests only."""

  def __init__(self, y0, group=None):
    self.n = y0.numel()
    self.y = y0.reshape(-1).to(torch.float64).clone()
    self.y_new = torch.empty_like(self.y)
    self.K = torch.zeros((7, self.n), dtype=torch.float32)
    self.y32 = torch.empty(self.n, dtype=torch.float32)
    self.group, self.n_total = group, self.n
    if group is not None:
      cnt = torch.tensor([float(self.n)], dtype=torch.float64)
      dist.all_reduce(cnt, group=group)
      self.n_total = int(cnt.item())

  def _comb(self, n_k, coef, h):
    if n_k == 0:
      return torch.zeros_like(self.y)
    return (self.K[:n_k].double().T @ torch.tensor(coef[:n_k], dtype=torch.float64)) * h

  def stage(self, n_k, coef, h, want_new=False):
    v = self.y + self._comb(n_k, coef, h)
    self.y32.copy_(v.to(torch.float32))
    if want_new:
      self.y_new.copy_(v)

  def rms(self, n_k, coef, h, rtol, atol, of_y=False, with_new=False):
    mag = self.y.abs()
    if with_new:
      mag = torch.maximum(mag, self.y_new.abs())
    v = self.y if of_y else self._comb(n_k, coef, h)
    out = ((v / (atol + mag * rtol)) ** 2).sum().reshape(1)
    if self.group is not None:
      dist.all_reduce(out, group=self.group)
    return math.sqrt(out.item() / self.n_total)

  def accept(self):
    self.y, self.y_new = self.y_new, self.y
    self.K[0].copy_(self.K[6])

  def k_row(self, j):
    return self.K[j]


def torch_fun(w):
  hf = host_fun(w.numpy())

  def fun(t, y32, out):
    out.copy_(torch.from_numpy(hf(t, y32.numpy()).astype(np.float32)))
  return fun


@pytest.mark.parametrize('span', [(0.0, 1.0), (1.0, 0.0)])
@pytest.mark.parametrize('tol', [1e-3, 1e-5, 1e-8])
def test_host_control_flow_matches_scipy(span, tol):
  """mulan_b200.ode.solve_ivp_rk45's accept/reject/step-size logic == scipy's."""
  from mulan_b200 import ode
  w, y0 = problem()
  s = solve_ivp(host_fun(w), span, y0, rtol=tol, atol=tol, method='RK45')
  yt = torch.from_numpy(y0)
  sol = ode.solve_ivp_rk45(torch_fun(torch.from_numpy(w)), span, yt, rtol=tol, atol=tol,
                           _state=CpuState(yt))
  assert sol.status == 0 and sol.nfev == s.nfev and sol.n_steps == len(s.t) - 1
  # the error estimate is a cancelling sum (E sums to 0): summation order moves it ~1e-12 rel
  assert np.allclose(np.array(sol.ts), s.t, rtol=1e-9, atol=0)
  assert np.allclose(sol.y.numpy(), s.y[:, -1], rtol=1e-9, atol=1e-11)


def test_device_integrator_refuses_cpu_tensors():
  from mulan_b200 import ode
  with pytest.raises(TypeError):
    ode.solve_ivp_rk45(lambda t, y, o: None, (0, 1), torch.zeros(4))


def test_bpd_offset_and_helpers():
  from mulan_b200 import ode
  assert ode._get_bpd_offset('uniform', 1) == 7.0
  for num_is in (1, 4):
    assert abs(ode._get_bpd_offset('tn', num_is) - R.bpd_offset('tn', num_is)) < 1e-15
  # -(0.5 (1 + log 2 pi) - 0.01522 + 0.5 (gt - softplus(gt))) / log 2 at gt = -13.3
  assert abs(ode._get_bpd_offset('tn', 1) - 7.5687) < 1e-3
  lg = torch.tensor([[float(i) for i in range(50)]])
  e = ode.logits_to_embeddings(lg, 15)
  assert e.sum().item() == 15 and e[0, 35:].min().item() == 1.0
  assert torch.equal(e, R.logits_to_embeddings(lg, 15))
  with pytest.raises(ValueError):
    ode._get_bpd_offset('gaussian', 1)


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _ode_worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                    WORLD_SIZE=str(world))
  torch.set_num_threads(1)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from mulan_b200 import ode
  w, y0 = problem(n=400)
  sl = slice(rank * 200, (rank + 1) * 200)
  yt = torch.from_numpy(y0[sl].copy())
  sol = ode.solve_ivp_rk45(torch_fun(torch.from_numpy(w[sl].copy())), (0.0, 1.0), yt, rtol=1e-6,
                           atol=1e-6, _state=CpuState(yt, group=dist.group.WORLD))
  q.put((rank, sol.nfev, sol.ts, sol.y.numpy().copy()))
  dist.barrier()
  dist.destroy_process_group()


def test_world_size_2_sharded_error_norm_gloo():
  """Rows sharded over 2 ranks: the all-reduced error norm gives every rank the step sequence of
  the single-process solve over the concatenated state (what the reference's host solver sees)."""
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_ode_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = sorted([q.get(timeout=300) for _ in procs], key=lambda r: r[0])
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  w, y0 = problem(n=400)
  s = solve_ivp(host_fun(w), (0.0, 1.0), y0, rtol=1e-6, atol=1e-6, method='RK45')
  for rank, nfev, ts, y in res:
    assert nfev == s.nfev
    assert np.allclose(np.array(ts), s.t, rtol=1e-9, atol=0)
    assert np.allclose(y, s.y[rank * 200:(rank + 1) * 200, -1], rtol=1e-9, atol=1e-11)


# ---------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize('n', [1, 9219, 3073 * 128])
def test_cuda_rk45_kernels_match_numpy(cuda_device, n):
  from mulan_b200 import _lib, ops
  rng = np.random.default_rng(n)
  y = rng.standard_normal(n)
  y_new = y + 0.01 * rng.standard_normal(n)
  n_pad = (n + 3) // 4 * 4
  K = np.zeros((7, n_pad), np.float32)
  K[:, :n] = rng.standard_normal((7, n)).astype(np.float32)
  dev = cuda_device
  yd, ynd, Kd = (torch.from_numpy(v).to(dev) for v in (y, y_new, K))
  y32 = torch.empty(n, dtype=torch.float32, device=dev)
  yo = torch.empty(n, dtype=torch.float64, device=dev)
  h = 0.0371
  for s in range(0, 7):
    coef = rng.standard_normal(7)
    ops.rk45_stage(s, coef, h, yd, Kd, y_stage=y32, y_out=yo)
    want = y + (K[:s, :n].astype(np.float64).T @ coef[:s]) * h if s else y
    assert np.allclose(yo.cpu().numpy(), want, rtol=1e-14, atol=1e-15)
    assert np.array_equal(y32.cpu().numpy(), yo.cpu().numpy().astype(np.float32))
  scratch = torch.empty(_lib.MULAN_RK45_SCRATCH, dtype=torch.float64, device=dev)
  out = torch.empty(1, dtype=torch.float64, device=dev)
  rtol, atol = 1e-5, 1e-6
  ops.rk45_norm(7, R.E, h, rtol, atol, yd, ynd, Kd, False, scratch, out)
  err = (K[:, :n].astype(np.float64).T @ R.E) * h
  want = np.sum((err / (atol + np.maximum(np.abs(y), np.abs(y_new)) * rtol)) ** 2)
  assert abs(out.item() - want) < 1e-12 * want
  first = out.item()
  ops.rk45_norm(7, R.E, h, rtol, atol, yd, ynd, Kd, False, scratch, out)
  assert out.item() == first                                   # fixed-order reduction
  ops.rk45_norm(0, (), 0.0, rtol, atol, yd, None, Kd, True, scratch, out)
  want = np.sum((y / (atol + np.abs(y) * rtol)) ** 2)
  assert abs(out.item() - want) < 1e-12 * want
  with pytest.raises(_lib.MulanError):
    ops.rk45_norm(0, (), 0.0, rtol, atol, yd, None, Kd, False, scratch, out)


@pytest.mark.gpu
@pytest.mark.parametrize('span', [(0.0, 1.0), (1.0, 0.0)])
@pytest.mark.parametrize('tol', [1e-3, 1e-6])
def test_cuda_integrator_matches_scipy(cuda_device, span, tol):
  """Same step sequence and final state as scipy driving the same float32 derivative through
  host round trips (what the reference does)."""
  from mulan_b200 import ode
  w, y0 = problem(n=9219)
  wd = torch.from_numpy(w).to(cuda_device)

  def dfun(t, y32, out):
    torch.add(torch.sin(np.float32(5 * t) + wd) * y32, float(np.float32(np.cos(3 * t))) * wd,
              out=out)

  def hfun(t, y):
    y32 = torch.from_numpy(np.asarray(y, np.float32)).to(cuda_device)
    out = torch.empty_like(y32)
    dfun(t, y32, out)
    return out.cpu().numpy().astype(np.float64)

  s = solve_ivp(hfun, span, y0, rtol=tol, atol=tol, method='RK45')
  sol = ode.solve_ivp_rk45(dfun, span, torch.from_numpy(y0).to(cuda_device), rtol=tol, atol=tol)
  assert sol.status == 0 and sol.nfev == s.nfev and sol.n_steps == len(s.t) - 1
  # Not bit-identical: a 1e-16 difference in the float64 stage sum (summation order) can flip
  # the float32 rounding of a stage input, and the error estimate - a cancelling sum - amplifies
  # that 1-ulp change of a derivative (measured: 1e-14 at tol 1e-3, 3e-6 at tol 1e-6).
  assert np.allclose(np.array(sol.ts), s.t, rtol=1e-4, atol=0)
  assert np.allclose(sol.y.cpu().numpy(), s.y[:, -1], rtol=tol, atol=tol)


def _likelihood_setup(dev, B=4, seed=21):
  from mulan_b200 import model as M
  g = torch.Generator().manual_seed(seed)
  conv = torch.nn.Conv2d(3, 3, 3, padding=1)
  with torch.no_grad():
    conv.weight.copy_(0.15 * torch.randn(conv.weight.shape, generator=g))
    conv.bias.copy_(0.05 * torch.randn(3, generator=g))
  We = 0.3 * torch.randn(256, 50, generator=g)
  Wc = 0.02 * torch.randn(50, generator=g)

  def make(device):
    cv = torch.nn.Conv2d(3, 3, 3, padding=1).to(device)
    cv.load_state_dict(conv.state_dict())
    we, wc = We.to(device), Wc.to(device)
    for p in cv.parameters():
      p.requires_grad_(False)

    def score(z, g_t, cond, det=True):
      h = torch.tanh(cv(z.permute(0, 3, 1, 2)).permute(0, 2, 3, 1))
      return 0.6 * z + 0.4 * h + (0.02 * g_t + cond @ wc).reshape(-1, 1, 1, 1)

    def encoder(orig_f, det=True):
      return orig_f.reshape(orig_f.shape[0], -1)[:, :256] @ we
    return encoder, score

  enc, score = make(dev)
  vdm = M.VDM(M.VDMConfig(), enc, score).to(dev)
  W = GI.mlp_weights(seed + 1000)
  vdm.gamma.load_flax(W)
  enc_c, score_c = make('cpu')
  W32 = {k: torch.from_numpy(v) for k, v in W.items()}
  ocfg = O.OracleConfig()

  def oracle_value_div(x, emb, t, noise, hp=False):
    a, b, c = O.compute_coefficients(W32, emb)
    td = torch.full((x.shape[0], 1), float(t), dtype=torch.float32)
    g_t = O.eval_polynomial(a, b, c, td, ocfg).reshape(x.shape)
    w = O.eval_polynomial_dt(a, b, c, td, ocfg).reshape(x.shape)
    g_in = g_t.mean(dim=(1, 2, 3))
    return O.value_div(x, g_t, w, lambda z: score_c(z, g_in, emb), noise, hp)

  rng = np.random.default_rng(seed)
  data = torch.from_numpy(rng.integers(0, 256, (B, 32, 32, 3), dtype=np.uint8))
  return vdm, enc_c, oracle_value_div, data, rng


@pytest.mark.gpu
@pytest.mark.parametrize('deq', ['tn', 'uniform'])
def test_cuda_likelihood_fn_matches_oracle(cuda_device, deq):
  """likelihood_fn (notebook_utils.py:303-371), device-resident, against the oracle driving the
  reference's algorithm on the host in float32 with the same draws."""
  from mulan_b200 import ode
  dev = cuda_device
  B = 4
  vdm, enc_c, oracle_vd, data, rng = _likelihood_setup(dev, B)
  if deq == 'tn':
    u = np.clip(rng.standard_normal((B, 32, 32, 3)), -3, 3).astype(np.float32)
  else:
    u = rng.random((B, 32, 32, 3), dtype=np.float32)
  v = (rng.integers(0, 2, (B, 32, 32, 3)) * 2 - 1).astype(np.float32)
  rtol = atol = 1e-4
  fn = ode.get_ode_likelihood_fn(vdm, 'Rademacher', rtol=rtol, atol=atol, dequantization=deq)
  draws = {'u': torch.from_numpy(u).to(dev), 'hutchinson': torch.from_numpy(v).to(dev)}
  log_p, log_q, aux = fn(None, data.to(dev), deterministic_noise=True, draws=draws)
  sol = fn.last_solution
  encode_then = lambda images_int: enc_c(2 * ((images_int.round() + .5) / 256) - 1)
  w_log_p, w_log_q, w_aux, w_sol = R.likelihood(
      data, torch.from_numpy(u), torch.from_numpy(v), encode_then, oracle_vd, deq, rtol, atol)
  assert sol.status == 0 and w_sol.status == 0
  assert sol.nfev == w_sol.nfev, (sol.nfev, w_sol.nfev)
  # float32 drifts differ by ~1 ulp between the CUDA kernels and torch-CPU; the step-size
  # controller (err ** -0.2 of a cancelling sum) turns that into ~3e-4 relative in the grid
  assert np.allclose(np.array(sol.ts), np.array(w_sol.ts), rtol=2e-3)
  assert torch.allclose(aux.cpu(), w_aux, rtol=1e-5, atol=1e-6)
  if deq == 'tn':
    assert torch.allclose(log_q.cpu(), w_log_q, rtol=1e-6)
  else:
    assert log_q is None and w_log_q is None
  # two solves whose grids differ agree to the integrator's own tolerance, atol + rtol |y|
  # (log_p ~ -4e3 per row)
  gap = (log_p.cpu() - w_log_p).abs()
  assert torch.all(gap < 0.5 * (atol + rtol * w_log_p.abs())), (log_p.cpu(), w_log_p)
  z = sol.y[:B * 3072].cpu().numpy()
  wz = w_sol.y[:B * 3072]
  assert np.all(np.abs(z - wz) < 0.5 * (atol + rtol * np.abs(wz)))


@pytest.mark.gpu
def test_cuda_ode_sample_fn_matches_oracle(cuda_device):
  """sample_fn (notebook_utils.py:397-431): t from 1 to 0, drift only."""
  from mulan_b200 import ode
  dev = cuda_device
  B = 3                                              # odd: state rows are not 16-byte multiples
  vdm, _, oracle_vd, _, rng = _likelihood_setup(dev, B)
  logits = torch.from_numpy(rng.standard_normal((B, 50)).astype(np.float32))
  prior = torch.from_numpy(rng.standard_normal((B, 32, 32, 3)).astype(np.float32))
  fn = ode.get_sample_fn(vdm, rtol=1e-4, atol=1e-4)
  z, nfev = fn(None, sample_size=B, device=dev,
               draws={'logits': logits.to(dev), 'prior': prior.to(dev)})
  emb = R.logits_to_embeddings(logits)
  zeros = torch.zeros(B, 32, 32, 3)

  def ode_func(t, x):
    xt = torch.from_numpy(np.asarray(x, np.float32)).reshape(B, 32, 32, 3)
    drift, _ = oracle_vd(xt, emb, t, zeros)
    return drift.detach().double().numpy().reshape(-1)

  w = R.solve_rk45(ode_func, (1.0, 0.0), prior.numpy().reshape(-1), rtol=1e-4, atol=1e-4)
  assert nfev == w.nfev
  assert z.shape == (B, 32, 32, 3)
  assert np.abs(z.cpu().numpy().reshape(-1) - w.y).max() < 1e-4 * np.abs(w.y).max()


@pytest.mark.gpu
def test_cuda_eval_bpd_ode_runs(cuda_device):
  """eval_bpd_ode (notebook_utils.py:451-531): importance-weighted and single-sample paths."""
  from mulan_b200 import ode
  vdm, _, _, data, _ = _likelihood_setup(cuda_device, 4)
  gen = torch.Generator(device=cuda_device).manual_seed(0)
  for num_is in (1, 2):
    bpd = ode.eval_bpd_ode(vdm, [data.to(cuda_device)], True, 'Rademacher', 'tn', num_is=num_is,
                           rtol=1e-3, atol=1e-3, generator=gen)
    assert math.isfinite(bpd)


@pytest.mark.gpu
def test_cuda_rk45_unaligned_vectors_take_the_scalar_path(cuda_device):
  """The 4-wide kernels need 16-byte aligned vectors; an 8-byte offset view must give the same
  numbers through the scalar kernels."""
  from mulan_b200 import _lib, ops
  n = 1001
  rng = np.random.default_rng(5)
  dev = cuda_device
  base = torch.from_numpy(rng.standard_normal(n + 1)).to(dev)
  K = torch.from_numpy(rng.standard_normal((7, 1004)).astype(np.float32)).to(dev)
  coef = rng.standard_normal(7)
  outs = []
  for y in (base[1:], base[1:].clone()):        # 8-byte offset view, then an aligned copy
    yo = torch.empty(n + 1, dtype=torch.float64, device=dev)[1:] if y.data_ptr() % 16 else \
        torch.empty(n, dtype=torch.float64, device=dev)
    y32 = torch.empty(n, dtype=torch.float32, device=dev)
    ops.rk45_stage(6, coef, 0.02, y, K, y_stage=y32, y_out=yo)
    scratch = torch.empty(_lib.MULAN_RK45_SCRATCH, dtype=torch.float64, device=dev)
    out = torch.empty(1, dtype=torch.float64, device=dev)
    ops.rk45_norm(7, R.E, 0.02, 1e-5, 1e-6, y, yo, K, False, scratch, out)
    outs.append((yo.cpu().numpy().copy(), y32.cpu().numpy().copy(), out.item()))
  assert base[1:].data_ptr() % 16 == 8
  assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
  assert abs(outs[0][2] - outs[1][2]) < 1e-13 * abs(outs[1][2])
