#!/bin/bash
# compute-sanitizer over a small end-to-end exercise of every kernel (memcheck, racecheck,
# initcheck, synccheck).  Summaries -> gpurun_out/sanitizer_*.txt
cat > /tmp/san_drive.py <<'PY'
import sys, math, torch, numpy as np
sys.path.insert(0, '.')
from mulan_b200 import ops, model as M, host
from oracle import mulan_oracle as O
dev = torch.device('cuda:0')
B = 5
inp = O.synth_inputs(B, 3)
g = {k: v.to(dev).contiguous() for k, v in inp.items()}
gL = torch.full((B,), 1e-6, device=dev)
for mode in (0, 1, 2):
  for gt in (0, 1):
    d = ops.Desc(param=mode, gt_mode=gt)
    r = ops.fwd_pre(d, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps_0'], g['eps'], save_w=(mode == 0))
    diff, nb = ops.fwd_bwd_post(d, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'], r['w'], gL)
    ops.fwd_post(d, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'], r['w'])
    ops.bwd_post(d, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'], r['w'], gL)
    gb = torch.zeros((B,) if gt == 0 else (B, 3072), device=dev)
    ops.bwd_pre(d, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'], torch.zeros_like(g['a']), gb, gL)
    ops.bpd_reduce(d, r['loss_recon'], r['loss_klz_prior'], None, diff, r['var_sums'])
    ops.bpd_reduce(d, r['loss_recon'], r['loss_klz_prior'], None, diff, r['var_sums'], ws=None)
    ops.post_bpd(d, g['x'], g['a'], g['b'], g['c'], g['t'], g['eps'], g['net'], r['w'], gL,
                 r['loss_recon'], r['loss_klz_prior'], None, r['var_sums'])
    ops.scale_rows(nb, gL * 2, gL)
# ABI v2: pre-activation c, broadcast noise rows, programmatic dependent launch; the three launch
# shapes (768 threads per row up to one row per SM, 256 up to four, 128 beyond) and a multi-group
# fused reduction
for rows_ in (6, 300, 1300):
  big = O.synth_inputs(rows_, 5)
  gb_ = {k: v.to(dev).contiguous() for k, v in big.items()}
  gLb = torch.full((rows_,), 1e-6, device=dev)
  for mode in (0, 1):
    d2 = ops.Desc(param=mode, c_raw=True, pdl=True, noise_rows=3)
    e0, e1 = gb_['eps_0'][:3].contiguous(), gb_['eps'][:3].contiguous()
    r = ops.fwd_pre(d2, gb_['x'], gb_['a'], gb_['b'], gb_['c'], gb_['t'], e0, e1, save_w=(mode == 0))
    ops.post_bpd(d2, gb_['x'], gb_['a'], gb_['b'], gb_['c'], gb_['t'], e1, gb_['net'], r['w'], gLb,
                 r['loss_recon'], r['loss_klz_prior'], None, r['var_sums'])
    ops.bwd_pre(d2, gb_['x'], gb_['a'], gb_['b'], gb_['c'], gb_['t'], e1, gb_['net'],
                torch.zeros_like(gb_['a']), torch.zeros(rows_, device=dev), gLb)
dT = ops.Desc(n_timesteps=10)
tT = O.sample_t(0.3, B, O.OracleConfig(sm_n_timesteps=10)).to(dev)
r = ops.fwd_pre(dT, g['x'], g['a'], g['b'], g['c'], tT, g['eps_0'], g['eps'])
ops.bwd_pre(dT, g['x'], g['a'], g['b'], g['c'], tT, g['eps'], g['net'], None, None, gL)
lg = torch.randn(B, 50, device=dev); G = torch.rand(10, B, 50, device=dev)
e, k = ops.aux_topk_fwd(lg, G, 15); ops.aux_topk_bwd(lg, G, 15, e, k)
d = ops.Desc()
t = torch.full((B,), 0.5, device=dev); s = torch.full((B,), 0.499, device=dev)
ops.sample_gamma(d, g['a'][:1], g['b'][:1], g['c'][:1], t)
zs = ops.sample_step(d, g['a'], g['b'], g['c'], t, s, g['eps'], g['net'], g['eps_0'])
ops.generate_x(d, zs)
# broadcast coefficients: persistent factor-table kernel, more rows than resident CTAs, a run of
# equal (t, s), a change of t mid-batch, and per-row times; eps and velocity models
Bs = 1600
zb_, nb_, eb_ = (torch.randn(Bs, 3072, device=dev) for _ in range(3))
for tt_ in (torch.full((Bs,), 0.5, device=dev),
            torch.where(torch.arange(Bs, device=dev) < 900, 0.5, 0.8).float(),
            torch.rand(Bs, device=dev) * 0.9 + 0.05):
  for mode in (0, 1):
    ops.sample_step(ops.Desc(param=mode), g['a'][:1], g['b'][:1], g['c'][:1], tt_, tt_ - 0.001,
                    zb_, nb_, eb_)
# in-kernel draws (row-pair CTAs) and the loss-scalar board (one rank: the peer stores land in
# this process's own board; mulan_adamw_ema_peer needs >= 2 processes and is covered by
# tests/test_peer.py)
big6 = O.synth_inputs(6, 7); g6 = {k: v.to(dev).contiguous() for k, v in big6.items()}
for mode in (0, 1):
  rk = ops.fwd_pre_keyed(ops.Desc(param=mode), (1, 2), (3, 4), g6['x'], g6['a'], g6['b'], g6['c'],
                         g6['t'], save_w=(mode == 0), want_eps=True, want_eps0=True)
import os, torch.distributed as dist
os.environ.setdefault('MASTER_ADDR', '127.0.0.1'); os.environ.setdefault('MASTER_PORT', '29791')
dist.init_process_group('gloo', rank=0, world_size=1)
from mulan_b200.peer import ScalarBoard
try:
  board = ScalarBoard(dev)
  for rows_ in (6, 300):
    bb = O.synth_inputs(rows_, 9); gb_ = {k: v.to(dev).contiguous() for k, v in bb.items()}
    wsb = ops.ElboWorkspace(ops.Desc(), rows_, dev)
    wsb.fwd_pre(gb_['x'], gb_['a'], gb_['b'], gb_['c'], gb_['t'], gb_['eps_0'], gb_['eps'])
    wsb.post_bpd(gb_['x'], gb_['a'], gb_['b'], gb_['c'], gb_['t'], gb_['eps'], gb_['net'],
                 torch.full((rows_,), 1e-6, device=dev), board=board)
    m_, e_ = board.read()
  torch.cuda.synchronize()
  print('board ok', m_.tolist(), int(e_))
  board.close()
except Exception as exc:     # CUDA IPC may be unavailable under the sanitizer
  print('board skipped:', exc)
dist.destroy_process_group()
# JAX-compatible draws
for shp in ((7,), (4096,), (3, 1001)):
  ops.rng_bits((1, 2), shp, device=dev); ops.rng_uniform((3, 4), shp, device=dev)
  ops.rng_normal((5, 6), shp, device=dev)
ops.ode_drift(d, g['a'], g['b'], g['c'], t, g['eps'], g['net'], g['eps_0'], True)
ops.row_dot(g['eps'], g['net'], gL)
npy = {k: v.numpy() for k, v in inp.items()}
host.elbo_host(npy['x'], npy['a'], npy['b'], npy['c'], npy['t'], npy['eps_0'], npy['eps'], npy['net'])
import ctypes as C
from mulan_b200 import _lib
n = 4096
bufs = [torch.randn(n, device=dev) for _ in range(5)]
dd = _lib.MulanAdamwDesc(n, 2048, 1, 0, 1e-4, 0.9, 0.99, 1e-8, 0.01, 0.9999, 1.0)
_lib.check(_lib.load().mulan_adamw_ema(C.byref(dd), *[C.c_void_p(b.data_ptr()) for b in bufs], None))
# aux-latent variants
nz = torch.randn(B, 50, device=dev); kb = torch.randn(B, device=dev)
lgr = lg.clone().requires_grad_(True)
for fn, extra in ((ops.aux_topk_add, 15), (ops.aux_gumbel, 0.7)):
  e, k = fn(lgr, nz, extra); ((e * nz).sum() + (k * kb).sum()).backward()
mu_ = torch.randn(B, 50, device=dev, requires_grad=True)
var_ = (torch.rand(B, 50, device=dev) + 0.1).requires_grad_(True)
e, k = ops.aux_gaussian(mu_, var_, nz); ((e * nz).sum() + (k * kb).sum()).backward()
# RK45 state kernels (odd n: rows not 16-byte multiples) and a short device-resident solve
from mulan_b200 import ode
nn_ = 3 * 3073
y = torch.randn(nn_, dtype=torch.float64, device=dev); yn = y + 0.01
K = torch.randn(7, (nn_ + 3) // 4 * 4, device=dev)
y32 = torch.empty(nn_, device=dev); yo = torch.empty_like(y)
scr = torch.empty(_lib.MULAN_RK45_SCRATCH, dtype=torch.float64, device=dev)
out1 = torch.empty(1, dtype=torch.float64, device=dev)
for s_ in range(7):
  ops.rk45_stage(s_, ode.RK45_E, 0.1, y, K, y_stage=y32, y_out=yo)
ops.rk45_norm(7, ode.RK45_E, 0.1, 1e-5, 1e-5, y, yn, K, False, scr, out1)
ops.rk45_norm(0, (), 0.0, 1e-5, 1e-5, y, None, K, True, scr, out1)
# 8-byte offset views: the scalar (non-vectorised) RK45 kernels
yb = torch.randn(1002, dtype=torch.float64, device=dev)
ops.rk45_stage(6, ode.RK45_E, 0.1, yb[1:], K, y_stage=torch.empty(1001, device=dev),
               y_out=torch.empty(1002, dtype=torch.float64, device=dev)[1:])
ops.rk45_norm(7, ode.RK45_E, 0.1, 1e-5, 1e-5, yb[1:], None, K, False, scr, out1)
ode.solve_ivp_rk45(lambda t_, yy, o: torch.mul(yy, -1.0, out=o), (0.0, 1.0), y32, 1e-3, 1e-3)
# global-norm clip
ss = torch.zeros(1, device=dev)
scr2 = torch.empty(_lib.MULAN_SUMSQ_SCRATCH, dtype=torch.float64, device=dev)
_lib.check(_lib.load().mulan_grad_sumsq(n, C.c_void_p(bufs[1].data_ptr()), C.c_void_p(scr2.data_ptr()),
                                        C.c_void_p(ss.data_ptr()), None))
dd = _lib.MulanAdamwDesc(n, 2048, 2, 0, 1e-4, 0.9, 0.99, 1e-8, 0.01, 0.9999, 1.0, 0.5, ss.data_ptr())
_lib.check(_lib.load().mulan_adamw_ema(C.byref(dd), *[C.c_void_p(b.data_ptr()) for b in bufs], None))
torch.cuda.synchronize()
print('drive ok')
PY
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  MULAN_FWD_PRE_TMA=0 timeout 600 compute-sanitizer --tool $tool --kernel-regex kns=mulan --print-limit 5 python /tmp/san_drive.py > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|drive ok|board ok|board skipped' gpurun_out/sanitizer_$tool.txt | tr '\n' ' ')"
done
# initcheck must see torch's own initialising kernels (a kernel filter makes every torch-written
# input look uninitialised) and needs the caching allocator off
MULAN_FWD_PRE_TMA=0 PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 900 compute-sanitizer --tool initcheck --print-limit 5 python /tmp/san_drive.py > gpurun_out/sanitizer_initcheck.txt 2>&1
echo "== initcheck: $(grep -E 'ERROR SUMMARY|drive ok|board ok|board skipped' gpurun_out/sanitizer_initcheck.txt | tr '\n' ' ')"
# (the opt-in TMA fwd_pre under racecheck: see profiles/r1_sanitizer.md; rerun with
#  MULAN_FWD_PRE_TMA=1 compute-sanitizer --tool racecheck --kernel-regex kns=mulan python /tmp/san_drive.py)
