"""CPU-side checks: the C-ABI library loads and exports every symbol include/mulan_b200.h
declares (no compute calls without a GPU), argument validation works, the bindings refuse
CPU tensors, and the host-side logic (t sampling, sharding, flat gradient bucket over gloo
with world_size 2, dense-eval sharding) is right."""
import ctypes as C
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'mulan_b200.h')


@pytest.fixture(scope='module')
def lib():
  from mulan_b200.build import build_library
  build_library()
  from mulan_b200 import _lib
  return _lib


def header_symbols():
  src = open(HEADER).read()
  src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
  return sorted(set(re.findall(r'^\s*(?:const\s+char\s*\*|int|void|size_t)\s+(mulan_\w+)\s*\(', src,
                               flags=re.M)))


def test_library_exports_every_declared_symbol(lib):
  syms = header_symbols()
  assert len(syms) >= 11, syms
  handle = lib.load()
  for s in syms:
    assert hasattr(handle, s), f'libmulan_b200.so does not export {s}'
  assert set(syms) == set(lib.SIGNATURES), (set(syms) ^ set(lib.SIGNATURES))
  assert handle.mulan_abi_version() == lib.MULAN_ABI_VERSION == 2


def test_desc_layout_matches_header(lib):
  # 6 x int32 + 2 x double, no padding surprises
  assert C.sizeof(lib.MulanDesc) == 48
  assert lib.MulanDesc.gamma_min.offset == 24 and lib.MulanDesc.gamma_max.offset == 32
  assert lib.MulanDesc.flags.offset == 40 and lib.MulanDesc.noise_rows.offset == 44


def test_argument_validation_without_gpu(lib):
  h = lib.load()
  assert h.mulan_fwd_pre(None, *([None] * 14)) == -1
  assert b'desc is NULL' in h.mulan_last_error()
  d = lib.make_desc(rows=2, dim=3070)
  assert h.mulan_fwd_pre(C.byref(d), *([None] * 14)) == -2
  d = lib.make_desc(rows=2, vocab=1)
  assert h.mulan_fwd_post(C.byref(d), *([None] * 10)) == -1
  d = lib.make_desc(rows=2, param=7)
  assert h.mulan_bwd_post(C.byref(d), *([None] * 11)) == -1
  d = lib.make_desc(rows=2, n_timesteps=1000, param=lib.MULAN_PARAM_VEL)
  assert h.mulan_bwd_pre(C.byref(d), *([None] * 14)) == -3     # velocity model asserts T == 0
  assert b'discrete time' in h.mulan_last_error()
  d = lib.make_desc(rows=2, gamma_min=5.0, gamma_max=-13.3)
  assert h.mulan_fwd_pre(C.byref(d), *([None] * 14)) == -1
  d = lib.make_desc(rows=2)
  assert h.mulan_fwd_pre(C.byref(d), *([None] * 14)) == -1     # x is NULL
  assert b'x is NULL' in h.mulan_last_error()
  assert h.mulan_aux_topk_fwd(4, 65, 15, None, None, None, None, None) == -1
  assert h.mulan_aux_topk_fwd(4, 50, 51, None, None, None, None, None) == -1
  # rows == 0 is a no-op that needs no pointers and no device
  d = lib.make_desc(rows=0)
  assert h.mulan_fwd_pre(C.byref(d), *([None] * 14)) == 0
  assert h.mulan_aux_topk_fwd(0, 50, 15, None, None, None, None, None) == 0


def test_keyed_host_entry_validates_without_gpu(lib):
  h = lib.load()
  d = lib.make_desc(rows=2)
  assert h.mulan_elbo_host_keyed(C.byref(d), *([None] * 8), lib.DENOISER_FN(), None, 0,
                                 *([None] * 6)) == -1
  assert b'key is NULL' in h.mulan_last_error()
  assert h.mulan_elbo_host(C.byref(d), *([None] * 8), lib.DENOISER_FN(), None, 0,
                           *([None] * 6)) == -1
  assert b'eps0 / eps is NULL' in h.mulan_last_error()


def test_shipped_configuration_selects_the_specialised_fwd_pre(lib):
  """The immediates baked into the specialised fwd_pre kernel are compared bit for bit with the
  constants THIS host's libm produces; a mismatch would silently fall back to the generic-
  constant kernel (same results, 4 % slower).  Other configurations take the other variants."""
  h = lib.load()
  assert h.mulan_fwd_pre_variant(C.byref(lib.make_desc(rows=1))) == 2
  assert h.mulan_fwd_pre_variant(C.byref(lib.make_desc(rows=1, gamma_max=4.0))) == 1
  assert h.mulan_fwd_pre_variant(C.byref(lib.make_desc(rows=1, gamma_min=-6.0))) == 0   # W > 1
  assert h.mulan_fwd_pre_variant(C.byref(lib.make_desc(rows=1, vocab=100))) == 0
  assert h.mulan_fwd_pre_variant(C.byref(lib.make_desc(rows=1, vocab=1))) == -1


def test_end_constants_can_be_supplied_by_the_caller(lib):
  """mulan_fwd_pre_consts: host-side plumbing (no device needed).  The library's own constants
  are the correctly rounded ones quoted in DESIGN.md section 2; supplying them back selects the
  same (immediates) kernel, a one-ulp different exp(-g_0/2) -- what numpy's float32 exp returns --
  selects the parameter-bank kernel."""
  h = lib.load()
  d = lib.make_desc(rows=1)
  k = lib.MulanEndConsts()
  assert h.mulan_host_end_consts(C.byref(d), C.byref(k)) == 0
  assert float(k.exp_neg_half_g0).hex() == '0x1.8264680000000p+9'
  assert float(k.exp_half_g0).hex() == '0x1.5338580000000p-10'
  assert abs(k.sigmoid_g1 - 1 / (1 + np.exp(-5.0))) < 1e-7
  assert abs(k.log_sigmoid_g1 - np.log(k.sigmoid_g1)) < 1e-9
  assert h.mulan_fwd_pre_variant_consts(C.byref(d), None) == 2
  assert h.mulan_fwd_pre_variant_consts(C.byref(d), C.byref(k)) == 2
  k2 = lib.MulanEndConsts.from_buffer_copy(k)
  k2.exp_neg_half_g0 = float(np.nextafter(np.float32(k.exp_neg_half_g0), np.float32(0)))
  assert float(k2.exp_neg_half_g0).hex() == '0x1.8264660000000p+9'
  assert h.mulan_fwd_pre_variant_consts(C.byref(d), C.byref(k2)) == 1
  # validation
  bad = lib.MulanEndConsts.from_buffer_copy(k)
  bad.sigmoid_g1 = 1.5
  assert h.mulan_fwd_pre_consts(C.byref(d), C.byref(bad), *([None] * 14)) == -1
  assert b'end constants out of range' in h.mulan_last_error()
  assert h.mulan_fwd_pre_consts(C.byref(d), C.byref(k), *([None] * 14)) == -1
  assert b'x is NULL' in h.mulan_last_error()
  assert h.mulan_fwd_pre_consts(C.byref(d), None, *([None] * 14)) == -1      # == mulan_fwd_pre
  assert b'mulan_fwd_pre: x is NULL' in h.mulan_last_error()
  dT = lib.make_desc(rows=1, n_timesteps=10)
  assert h.mulan_fwd_pre_consts(C.byref(dT), C.byref(k), *([None] * 14)) == -3


def test_ops_refuse_cpu_tensors(lib):
  from mulan_b200 import ops
  z = torch.zeros(2, 3072)
  with pytest.raises(TypeError, match='no CPU path'):
    ops.fwd_pre(ops.Desc(), torch.zeros(2, 3072, dtype=torch.uint8), z, z, z, torch.zeros(2), z, z)
  with pytest.raises(TypeError):
    ops.aux_topk_fwd(torch.zeros(2, 50), None, 15)


def test_product_path_does_not_import_oracle():
  for fn in os.listdir(os.path.join(ROOT, 'mulan_b200')):
    if fn.endswith('.py'):
      src = open(os.path.join(ROOT, 'mulan_b200', fn)).read()
      assert 'import oracle' not in src and 'from oracle' not in src, fn


def test_sample_t_matches_oracle():
  from mulan_b200.model import VDMConfig, sample_t
  from oracle import mulan_oracle as O
  for B in (1, 2, 8, 127, 128):
    for t0 in (0.0, 0.123456, 0.999):
      got = sample_t(torch.tensor(t0), B, VDMConfig())
      want = O.sample_t(t0, B, O.OracleConfig())
      assert torch.equal(got, want)
  got = sample_t(torch.tensor(0.37), 8, VDMConfig(sm_n_timesteps=10))
  assert torch.equal(got, O.sample_t(0.37, 8, O.OracleConfig(sm_n_timesteps=10)))


def test_non_antithetic_time_sampling():
  """ldm/model_mulan_epsilon.py:291-297: without antithetic sampling t is one uniform draw PER
  EXAMPLE, and it is discretised like the antithetic branch when sm_n_timesteps > 0."""
  from mulan_b200.model import VDM, VDMConfig, sample_t
  cfg = VDMConfig(antithetic_time_sampling=False, sm_n_timesteps=10)
  f = lambda *a, **k: None
  model = VDM(cfg, f, f)
  draws = model.make_draws(5, 'cpu', torch.Generator().manual_seed(0))
  assert tuple(draws['t0'].shape) == (5,)
  t = sample_t(draws['t0'], 5, cfg)
  assert tuple(t.shape) == (5,)
  assert torch.equal(t, torch.ceil(draws['t0'] * 10) / 10)
  assert len(set(t.tolist())) > 1
  cfg0 = VDMConfig(antithetic_time_sampling=False)
  assert torch.equal(sample_t(draws['t0'], 5, cfg0), draws['t0'])
  # antithetic: still the scalar draw
  assert tuple(VDM(VDMConfig(), f, f).make_draws(5, 'cpu')['t0'].shape) == ()


def test_model_rejects_off_path_configs():
  from mulan_b200.model import VDM, VDMConfig
  f = lambda *a, **k: None
  with pytest.raises(NotImplementedError):
    VDM(VDMConfig(gamma_type='learnable_nnet'), f, f)
  with pytest.raises(NotImplementedError):
    VDM(VDMConfig(latent_type='vq'), f, f)
  with pytest.raises(NotImplementedError):
    VDM(VDMConfig(topk_noise_type='laplace'), f, f)
  with pytest.raises(NotImplementedError):
    VDM(VDMConfig(vdm_type='vdm'), f, f)
  for lt in ('topk', 'gumbel', 'gaussian'):
    VDM(VDMConfig(latent_type=lt), f, f)


def test_schedule_head_matches_oracle_coefficients():
  """The cuBLAS-side MLP (torch Linear) restates _compute_coefficients; check on CPU."""
  from mulan_b200.model import NoiseSchedule_polynomial_fixedend, VDMConfig
  from oracle import mulan_oracle as O
  sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
  import golden_inputs as GI
  W = GI.mlp_weights(5)
  head = NoiseSchedule_polynomial_fixedend(VDMConfig())
  assert torch.count_nonzero(head.dense_out_a.weight) == 0      # zero-init like the reference
  head.load_flax(W)
  emb = torch.randn(3, 50)
  a, b, c = head._compute_coefficients(emb)
  oa, ob, oc = O.compute_coefficients({k: torch.from_numpy(v) for k, v in W.items()}, emb)
  for u, v in ((a, oa), (b, ob), (c, oc)):
    assert (u - v).abs().max().item() < 2e-4 * v.abs().max().item()
  assert torch.all(c > 1e-3)


def test_shard_rows():
  from mulan_b200.dist import shard_rows
  for n in (0, 1, 7, 128, 1000):
    for w in (1, 2, 3, 8):
      got = []
      for r in range(w):
        s = shard_rows(n, r, w)
        got += list(range(n))[s]
      assert got == list(range(n))


def test_dense_launches_are_sized_in_whole_images():
  """The dense-VLB driver sizes its launches itself (whole images, as many as fit 16384 rows;
  ldm/notebook_utils.py:176-191 evaluates 16 x 128 rows per call)."""
  from mulan_b200.dist import dense_images_per_launch
  assert dense_images_per_launch(128) == 128
  assert dense_images_per_launch(1000) == 16
  assert dense_images_per_launch(128, max_rows=2048) == 16       # the reference's shape
  assert dense_images_per_launch(100000) == 1                     # never less than one image


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


class _CpuStubVDM(torch.nn.Module):
  """VDM-shaped CPU model built on the ORACLE (test-only), to exercise the multi-process
  host logic without a GPU."""

  def __init__(self):
    super().__init__()
    from mulan_b200.model import VDMConfig
    self.config = VDMConfig()
    self.w = torch.nn.Parameter(torch.tensor(0.5))
    self.head = torch.nn.Linear(4, 3, bias=False)

  def make_draws(self, n, device, gen):
    return dict(t0=torch.rand((), generator=gen), G=torch.zeros(10, n, 50),
                eps_0=torch.randn((n, 32, 32, 3), generator=gen),
                eps=torch.randn((n, 32, 32, 3), generator=gen))

  def forward(self, images, labels=None, conditioning=None, step=0, deterministic=True,
              draws=None, generator=None):
    from mulan_b200.model import VDMOutput, sample_t
    from oracle import mulan_oracle as O
    B = images.shape[0]
    if draws is None:
      draws = self.make_draws(B, images.device, generator)
    t = draws['t'].reshape(B) if 't' in draws else sample_t(draws['t0'], B, self.config)
    pix = images.reshape(B, -1).float() / 255.0
    a = pix * self.head.weight[0, 0] + 0.3
    b = pix * self.head.weight[1, 1] - 0.2
    c = 1e-3 + torch.nn.functional.softplus(pix * self.head.weight[2, 2])
    # the dense-VLB driver hands over ONE key's draws for several images (broadcast rows)
    tile = lambda v: v if v.shape[0] == B else v.repeat(B // v.shape[0], 1, 1, 1)
    out = O.elbo_terms(images.reshape(B, -1), a, b, c, t, tile(draws['eps_0']).reshape(B, -1),
                       tile(draws['eps']).reshape(B, -1), lambda z, g: self.w * z, O.MODE_EPS,
                       O.OracleConfig())
    return VDMOutput(*out)


def _worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                    WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
  torch.set_num_threads(2)
  from mulan_b200 import dist as md
  r, w, _ = md.init_distributed('gloo')
  assert (r, w) == (rank, world)
  torch.manual_seed(0)
  model = _CpuStubVDM()
  rng = np.random.default_rng(0)
  images = torch.from_numpy(rng.integers(0, 256, (8, 32, 32, 3), dtype=np.uint8))
  # --- train step: each rank sees its shard; gradients and scalars are pmean'ed
  bucket = md.FlatGradBucket(model.parameters(), extra=6)
  opt = torch.optim.SGD(model.parameters(), lr=0.0)
  sl = md.shard_rows(8, rank, world)
  gen = torch.Generator().manual_seed(100 + rank)
  draws = model.make_draws(sl.stop - sl.start, 'cpu', gen)
  scalars = md.train_step(model, opt, bucket, {'images': images[sl]}, 0, draws=draws)
  # --- dense eval, sharded by image
  mean_bpd, mine = md.eval_bpd_dense_sampling(model, images[:4], n_timesteps=4,
                                              images_per_launch=2, seed=3)
  q.put((rank, bucket.flat.numpy().copy(), {k: float(v) for k, v in scalars.items()}, mean_bpd,
         mine.numpy().copy(), {k: v.numpy().copy() for k, v in draws.items()}))
  dist.barrier()
  dist.destroy_process_group()


def test_world_size_2_gloo():
  """N>1 host path on CPU: flat-bucket gradient pmean + scalar pmean (ldm/experiment.py:341,
  347) and the example-sharded dense evaluation reduction."""
  from mulan_b200 import dist as md
  from mulan_b200.model import loss_fn
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = sorted([q.get(timeout=300) for _ in procs], key=lambda r: r[0])
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  (r0, flat0, sc0, bpd0, mine0, d0), (r1, flat1, sc1, bpd1, mine1, d1) = res
  flat0, flat1 = torch.from_numpy(flat0), torch.from_numpy(flat1)
  d0 = {k: torch.from_numpy(v) for k, v in d0.items()}
  d1 = {k: torch.from_numpy(v) for k, v in d1.items()}
  assert torch.equal(flat0, flat1)                      # all-reduced bucket is replicated
  assert sc0 == sc1 and bpd0 == bpd1
  # single-process reference of the same computation
  torch.manual_seed(0)
  model = _CpuStubVDM()
  rng = np.random.default_rng(0)
  images = torch.from_numpy(rng.integers(0, 256, (8, 32, 32, 3), dtype=np.uint8))
  grads, bpds = [], []
  for rank, draws in ((0, d0), (1, d1)):
    sl = md.shard_rows(8, rank, 2)
    model.zero_grad()
    for p_ in model.parameters():
      p_.grad = None
    bpd, _ = loss_fn(model, {'images': images[sl]}, draws=draws)
    bpd.backward()
    grads.append(torch.cat([p_.grad.reshape(-1) for p_ in model.parameters()]))
    bpds.append(bpd.item())
  want = (grads[0] + grads[1]) / 2
  n = want.numel()
  assert torch.allclose(flat0[:n], want, rtol=1e-5, atol=1e-9)
  assert abs(sc0['bpd'] - (bpds[0] + bpds[1]) / 2) < 1e-5 * abs(sc0['bpd'])
  # dense eval: ranks took images[0::2], images[1::2]; the mean covers all four
  all_bpd, _ = md.eval_bpd_dense_sampling(model, images[:4], n_timesteps=4, images_per_launch=2,
                                          seed=3)
  assert abs(bpd0 - all_bpd) < 1e-6 * abs(all_bpd)
  assert mine0.size == 2 and mine1.size == 2


def test_bench_reference_arm_contract():
  """`bench.py --impl reference` prints exactly ONE JSON line with the contract's keys (the CPU
  oracle port is the reference arm; no GPU needed)."""
  import json
  import subprocess
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                        '--steps', '1', '--warmup', '0', '--ref-rows', '8'],
                       capture_output=True, text=True, timeout=600)
  assert out.returncode == 0, out.stderr[-2000:]
  lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
  assert len(lines) == 1
  d = json.loads(lines[0])
  for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
            'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline',
            'e2e'):
    assert k in d, k
  assert d['impl'] == 'reference' and d['vs_baseline'] is None and d['value'] > 0
  assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
  assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
  assert 'workload' in d['config']


def test_native_bench_fails_loudly_without_gpu():
  """No CPU fallback: without a CUDA device the native arm must exit non-zero and print no
  JSON line."""
  import subprocess
  if torch.cuda.is_available():
    pytest.skip('a GPU is present')
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1',
                        '--no-e2e', '--no-cpu-baseline'], capture_output=True, text=True,
                       timeout=600)
  assert out.returncode != 0
  assert out.stdout.strip() == ''


def test_headers_are_plain_c_and_link_from_c(lib, tmp_path):
  """The boundary is a C ABI: both public headers compile as strict C99 (and C++11), and a C
  program linked against libmulan_b200.so reaches the entry points (argument validation only:
  no device needed)."""
  import shutil
  import subprocess
  if shutil.which('gcc') is None:
    pytest.skip('no gcc')
  src = tmp_path / 'abi.c'
  src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "mulan_b200.h"
#include "mulan_b200_xla.h"
int main(void) {
  mulan_desc d;
  memset(&d, 0, sizeof d);
  d.rows = 2; d.dim = 3070; d.vocab = 256; d.gamma_min = -13.3; d.gamma_max = 5.0;
  int st = mulan_fwd_pre(&d, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
  printf("%d %d %d %d|%s\n", mulan_abi_version(), (int)sizeof(mulan_desc),
         (int)sizeof(mulan_xla_opaque), st, mulan_last_error());
  printf("%d %d\n", mulan_kernel_param(MULAN_PARAM_VEL_FROM_EPS), mulan_kernel_param(MULAN_PARAM_VEL));
  return 0;
}
''')
  inc = os.path.join(ROOT, 'include')
  libdir = os.path.join(ROOT, 'mulan_b200')
  exe = tmp_path / 'abi'
  subprocess.run(['gcc', '-std=c99', '-Wall', '-Wextra', '-Werror', '-pedantic', '-I', inc,
                  str(src), '-o', str(exe), '-L', libdir, '-lmulan_b200',
                  '-Wl,-rpath,' + libdir], check=True, capture_output=True)
  env = {k: v for k, v in os.environ.items() if k != 'MULAN_VFE_LITERAL'}
  out = subprocess.run([str(exe)], check=True, capture_output=True, text=True, env=env).stdout
  first, second = out.strip().split('\n')
  assert first.startswith('2 48 56 -2|') and 'not a multiple of 4' in first
  assert second == '0 1'
  subprocess.run(['g++', '-std=c++11', '-Wall', '-Wextra', '-Werror', '-fsyntax-only', '-I', inc,
                  '-x', 'c++', str(src)], check=True, capture_output=True)
