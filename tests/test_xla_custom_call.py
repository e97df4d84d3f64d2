"""XLA legacy custom-call targets (include/mulan_b200_xla.h): the binding surface of the jaxlib
the reference pins.  CPU: every declared target is exported, the opaque is validated, failures
reach XlaCustomCallStatusSetFailure when the hosting process provides it.  GPU: driven exactly
as XLA would (stream, void** buffers = operands then results, opaque bytes), each target
reproduces the direct C-ABI call bit for bit."""
import ctypes as C
import math
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'mulan_b200_xla.h')


@pytest.fixture(scope='module')
def lib():
  from mulan_b200.build import build_library
  build_library()
  from mulan_b200 import _lib
  return _lib


def test_library_exports_every_declared_target(lib):
  src = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
  syms = sorted(set(re.findall(r'^\s*void\s+(mulan_xla_\w+)\s*\(', src, flags=re.M)))
  assert len(syms) == 8, syms
  h = lib.load()
  for s in syms:
    assert hasattr(h, s), s
  assert set(syms) == set(lib.XLA_SIGNATURES)
  assert C.sizeof(lib.MulanXlaOpaque) == 56 and lib.MulanXlaOpaque.absent_mask.offset == 48
  assert C.sizeof(lib.MulanXlaAuxOpaque) == 16


def test_python_side_opaque_packing_matches_the_struct(lib):
  """jax_binding/mulan_jax_legacy.py packs the opaque with struct.pack('<6i2dIi2I', ...)."""
  import struct
  src = open(os.path.join(ROOT, 'jax_binding', 'mulan_jax_legacy.py')).read()
  assert "struct.pack('<6i2dIi2I'" in src
  op = lib.MulanXlaOpaque(desc=lib.make_desc(rows=7, dim=3072, vocab=256, param=2, gt_mode=1,
                                             n_timesteps=0, gamma_min=-13.3, gamma_max=5.0),
                          absent_mask=0x240, reserved=0)
  assert bytes(op) == struct.pack('<6i2dIi2I', 7, 3072, 256, 2, 1, 0, -13.3, 5.0, 0, 0, 0x240, 0)


def test_opaque_is_validated_without_gpu(lib, capfd):
  h = lib.load()
  bufs = (C.c_void_p * 13)()
  h.mulan_xla_fwd_pre(None, bufs, b'\0' * 7, 7, None)
  assert b'opaque must be a mulan_xla_opaque (56 bytes)' in h.mulan_last_error()
  assert 'mulan_xla_fwd_pre' in capfd.readouterr().err
  # a well-formed opaque with a bad descriptor: the entry point's own message comes through
  op = lib.MulanXlaOpaque(desc=lib.make_desc(rows=2, dim=3070))
  h.mulan_xla_bwd_pre(None, bufs, bytes(op), C.sizeof(op), None)
  assert b'not a multiple of 4' in h.mulan_last_error()
  # rows == 0 is a no-op that touches no buffer
  op = lib.MulanXlaOpaque(desc=lib.make_desc(rows=0))
  h.mulan_xla_fwd_post(None, bufs, bytes(op), C.sizeof(op), None)
  h.mulan_xla_aux_topk_fwd(None, bufs, bytes(lib.MulanXlaAuxOpaque(0, 50, 15, 0)), 16, None)
  h.mulan_xla_aux_topk_fwd(None, bufs, b'\0' * 12, 12, None)
  assert b'mulan_xla_aux_opaque' in h.mulan_last_error()


def test_failure_reaches_xla_status_when_the_host_process_exports_it(lib):
  """In a JAX process xla_extension exports XlaCustomCallStatusSetFailure; emulate that with
  a five-line shared object loaded RTLD_GLOBAL in a child process."""
  code = r'''
import ctypes as C, os, subprocess, sys, tempfile
sys.path.insert(0, %r)
d = tempfile.mkdtemp()
src = os.path.join(d, 's.c')
open(src, 'w').write("""
#include <string.h>
char got[512]; void* got_status;
void XlaCustomCallStatusSetFailure(void* status, const char* m, size_t n) {
  got_status = status; memcpy(got, m, n < 511 ? n : 511); }
""")
so = os.path.join(d, 'libstub.so')
subprocess.run(['gcc', '-shared', '-fPIC', '-o', so, src], check=True)
stub = C.CDLL(so, mode=C.RTLD_GLOBAL)
from mulan_b200 import _lib
h = _lib.load()
op = _lib.MulanXlaOpaque(desc=_lib.make_desc(rows=2, param=9))
bufs = (C.c_void_p * 13)()
status = C.c_void_p(0x1234)
h.mulan_xla_fwd_pre(None, bufs, bytes(op), C.sizeof(op), status)
msg = C.string_at(C.addressof(C.c_char.in_dll(stub, 'got')))
assert C.c_void_p.in_dll(stub, 'got_status').value == 0x1234
assert msg.startswith(b'mulan_xla_fwd_pre: ') and b'not a mulan_param' in msg, msg
print('ok')
''' % ROOT
  out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
  assert out.returncode == 0 and 'ok' in out.stdout, out.stderr[-2000:]


# ------------------------------------------------------------------------------------------
# GPU: call the targets the way XLA does
# ------------------------------------------------------------------------------------------
def _call(lib, name, operands, results, opaque, absent=()):
  """operands / results: CUDA tensors (or None for an operand marked absent)."""
  h = lib.load()
  tensors = list(operands) + list(results)
  bufs = (C.c_void_p * len(tensors))()
  mask = 0
  for i, tsr in enumerate(tensors):
    if tsr is None:
      mask |= 1 << i
      bufs[i] = 0xdead0000          # XLA never passes NULL: an absent operand is a dummy
    else:
      bufs[i] = tsr.data_ptr()
  opaque.absent_mask = mask
  stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
  getattr(h, name)(stream, bufs, bytes(opaque), C.sizeof(opaque), None)
  torch.cuda.synchronize()


@pytest.mark.gpu
@pytest.mark.parametrize('mode', ['eps', 'vel', 'vel_from_eps'])
def test_targets_reproduce_the_c_abi(cuda_device, lib, mode):
  from mulan_b200 import ops
  from oracle import mulan_oracle as O
  dev = cuda_device
  B, D = 24, 3072
  param = {'eps': 0, 'vel': 1, 'vel_from_eps': 2}[mode]
  inp = O.synth_inputs(B, 33)
  g = {k: v.to(dev).contiguous() for k, v in inp.items()}
  desc = ops.Desc(param=param)
  ws = ops.ElboWorkspace(desc, B, dev)
  gL = torch.full((B,), 1.0 / (B * D * math.log(2.0)), device=dev)
  rng = np.random.default_rng(2)
  zb = torch.from_numpy(1e-4 * rng.standard_normal((B, D)).astype(np.float32)).to(dev)
  gb = torch.from_numpy(1e-3 * rng.standard_normal(B).astype(np.float32)).to(dev)
  args = (g['x'], g['a'], g['b'], g['c'], g['t'])
  ws.fwd_pre(*args, g['eps_0'], g['eps'])
  ws.fwd_bwd_post(*args, g['eps'], g['net'], gL)
  ws.bpd_reduce(None)
  ws.bwd_pre(*args, g['eps'], g['net'], zb, gb, gL)
  torch.cuda.synchronize()

  f = lambda *s: torch.full(s, float('nan'), dtype=torch.float32, device=dev)
  op = lambda: lib.MulanXlaOpaque(desc=lib.make_desc(rows=B, param=param))
  z_t, g_net, w, rec, klz, vs = f(B, D), f(B), (f(B, D) if ws.w is not None else None), f(B), f(B), f(B, 2)
  _call(lib, 'mulan_xla_fwd_pre', [*args, g['eps_0'], g['eps']], [z_t, g_net, w, rec, klz, vs], op())
  assert torch.equal(z_t, ws.z_t) and torch.equal(g_net, ws.g_net)
  assert torch.equal(rec, ws.loss_recon) and torch.equal(klz, ws.loss_klz_prior)
  assert torch.equal(vs, ws.var_sums)
  if w is not None:
    assert torch.equal(w, ws.w)

  diff, n_bar = f(B), f(B, D)
  _call(lib, 'mulan_xla_fwd_bwd_post', [*args, g['eps'], g['net'], w, gL], [diff, n_bar], op())
  assert torch.equal(diff, ws.loss_diff) and torch.equal(n_bar, ws.n_bar)
  diff2, n_bar2 = f(B), f(B, D)
  _call(lib, 'mulan_xla_fwd_post', [*args, g['eps'], g['net'], w], [diff2], op())
  _call(lib, 'mulan_xla_bwd_post', [*args, g['eps'], g['net'], w, gL], [n_bar2], op())
  assert torch.equal(diff2, ws.loss_diff) and torch.equal(n_bar2, ws.n_bar)

  ab, bb, cb = f(B, D), f(B, D), f(B, D)
  _call(lib, 'mulan_xla_bwd_pre', [*args, g['eps'], g['net'], zb, gb, gL], [ab, bb, cb], op())
  assert torch.equal(ab, ws.a_bar) and torch.equal(bb, ws.b_bar) and torch.equal(cb, ws.c_bar)
  # the vjp of mulan_pre alone: no loss cotangent, net absent
  ab2 = f(B, D)
  _call(lib, 'mulan_xla_bwd_pre', [*args, g['eps'], None, zb, gb, None], [ab2, bb, cb], op())
  want = ops.bwd_pre(desc, *args, g['eps'], None, zb, gb, None)[0]
  assert torch.equal(ab2, want)

  scalars, klz_tot = f(6), f(B)
  _call(lib, 'mulan_xla_bpd_reduce', [rec, klz, None, diff, vs], [scalars, klz_tot], op())
  assert torch.equal(scalars, ws.scalars) and torch.equal(klz_tot, ws.loss_klz)


@pytest.mark.gpu
def test_aux_targets_reproduce_the_c_abi(cuda_device, lib):
  from mulan_b200 import ops
  dev = cuda_device
  B, L, K = 12, 50, 15
  rng = np.random.default_rng(8)
  logits = torch.from_numpy(rng.standard_normal((B, L)).astype(np.float32)).to(dev)
  draw = torch.from_numpy(rng.gamma(1.0 / K, size=(10, B, L)).astype(np.float32) + 1e-20).to(dev)
  emb_bar = torch.from_numpy(rng.standard_normal((B, L)).astype(np.float32)).to(dev)
  klz_bar = torch.from_numpy(rng.standard_normal(B).astype(np.float32)).to(dev)
  emb_w, klz_w = ops.aux_topk_fwd(logits, draw, K)
  lb_w = ops.aux_topk_bwd(logits, draw, K, emb_bar, klz_bar)
  f = lambda *s: torch.full(s, float('nan'), dtype=torch.float32, device=dev)
  emb, klz, lb = f(B, L), f(B), f(B, L)
  _call(lib, 'mulan_xla_aux_topk_fwd', [logits, draw], [emb, klz], lib.MulanXlaAuxOpaque(B, L, K, 0))
  _call(lib, 'mulan_xla_aux_topk_bwd', [logits, draw, emb_bar, klz_bar], [lb],
        lib.MulanXlaAuxOpaque(B, L, K, 0))
  assert torch.equal(emb, emb_w) and torch.equal(klz, klz_w) and torch.equal(lb, lb_w)
