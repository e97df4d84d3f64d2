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
    ops.scale_rows(nb, gL * 2, gL)
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
torch.cuda.synchronize()
print('drive ok')
PY
for tool in memcheck racecheck initcheck synccheck; do
  MULAN_FWD_PRE_TMA=0 compute-sanitizer --tool $tool --kernel-regex kns=mulan --print-limit 5 python /tmp/san_drive.py > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|drive ok' gpurun_out/sanitizer_$tool.txt | tr '\n' ' ')"
done
MULAN_FWD_PRE_TMA=1 compute-sanitizer --tool racecheck --kernel-regex kns=mulan --print-limit 5 python /tmp/san_drive.py > gpurun_out/sanitizer_racecheck_tma.txt 2>&1
echo "== racecheck (TMA fwd_pre): $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|drive ok' gpurun_out/sanitizer_racecheck_tma.txt | tr '\n' ' ')"
