"""Device-side JAX-compatible random draws: jax.random.{bits,uniform,normal}
for float32 under threefry2x32 (ldm/model_mulan_epsilon.py:287-292, :315, :327).

JAX is not installable here, so the oracle (oracle/jax_rng_oracle.py) is pinned to PUBLISHED
known answers: the Random123 Threefry-2x32 vectors (the ones jax/tests/random_test.py uses) and
the values printed in the JAX documentation for PRNGKey(0) / PRNGKey(42).
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import jax_rng_oracle as J  # noqa: E402


def test_threefry2x32_random123_vectors():
  kat = [((0x0, 0x0), (0x0, 0x0), (0x6b200159, 0x99ba4efe)),
         ((0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x1cb996fc, 0xbb002be7)),
         ((0x13198a2e, 0x03707344), (0x243f6a88, 0x85a308d3), (0xc4923a9c, 0x483df7a0))]
  for key, ctr, want in kat:
    y0, y1 = J.threefry2x32(key, ctr[0], ctr[1])
    assert (int(y0), int(y1)) == want


def test_oracle_reproduces_jax_documentation_values():
  k0 = J.prng_key(0)
  assert k0 == (0, 0) and J.prng_key(42) == (0, 42)
  assert J.uniform(k0, 1)[0] == np.float32(0.41845703)
  assert J.normal(k0, 1)[0] == np.float32(-0.20584226)
  assert np.array_equal(J.normal(k0, 3), np.array([1.8160863, -0.48262316, 0.33988908],
                                                  np.float32))
  assert J.normal(J.prng_key(42), 1)[0] == np.float32(-0.18471177)


def test_oracle_pairing_and_ranges():
  """Element i shares its Threefry block with element i + ceil(n/2); odd n pads a zero counter."""
  key = (123, 456)
  for n in (1, 2, 7, 8):
    b = J.random_bits(key, n)
    half = (n + 1) // 2
    for i in range(half):
      j = i + half
      y0, y1 = J.threefry2x32(key, i, j if j < n else 0)
      assert b[i] == y0 and (j >= n or b[j] == y1)
  u = J.uniform(key, 100001)
  assert u.min() >= 0 and u.max() < 1 and abs(u.mean() - 0.5) < 5e-3
  z = J.normal(key, 200001)
  assert abs(z.mean()) < 1e-2 and abs(z.std() - 1) < 1e-2 and np.isfinite(z).all()
  # the erf_inv tail branch (w >= 5) is exercised and continuous across the switch
  x = np.float32([0.9966, 0.99665, 0.9967, -0.9999999])
  e = J.erf_inv(x)
  from scipy.special import erfinv
  assert np.allclose(e, erfinv(x.astype(np.float64)), rtol=2e-6)
  g = J.gumbel(key, 1001)
  assert np.isfinite(g).all()


@pytest.mark.gpu
@pytest.mark.parametrize('n', [1, 2, 3, 8 * 3072, 8 * 3072 + 1, 1_000_003])
def test_cuda_rng_matches_oracle(cuda_device, n):
  from mulan_b200 import ops
  for key in ((0, 0), (0, 42), (0xDEADBEEF, 0x12345678)):
    bits = ops.rng_bits(key, (n,), cuda_device).cpu().numpy().view(np.uint32)
    assert np.array_equal(bits, J.random_bits(key, n))                   # integers: bit-exact
    u = ops.rng_uniform(key, (n,), -2.0, 3.0, cuda_device).cpu().numpy()
    assert np.array_equal(u, J.uniform(key, n, -2.0, 3.0))              # one mul + add: exact
    z = ops.rng_normal(key, (n,), cuda_device).cpu().numpy()
    want = J.normal(key, n)
    # log1pf / sqrtf of libdevice vs numpy's libm: a few float32 ulps
    assert np.all(np.abs(z - want) <= 4e-7 * np.maximum(np.abs(want), 1.0)), np.abs(z - want).max()
  z = ops.rng_normal((0, 0), (1,), cuda_device).item()
  assert abs(z - (-0.20584226)) < 1e-7                                   # the documented value


@pytest.mark.gpu
def test_vdm_make_draws_with_jax_keys(cuda_device):
  """VDM.make_draws(jax_keys=...) produces t0 / eps_0 / eps as jax.random would for those keys."""
  from mulan_b200.model import VDM, VDMConfig, sample_t
  f = lambda *a, **k: None
  vdm = VDM(VDMConfig(), f, f).to(cuda_device)
  keys = {'t0': (0, 0), 'eps_0': (11, 12), 'eps': (13, 14)}
  d = vdm.make_draws(6, cuda_device, torch.Generator(device=cuda_device).manual_seed(0),
                     jax_keys=keys)
  assert d['t0'].item() == np.float32(0.41845703)            # uniform(PRNGKey(0), ())
  assert d['G'].shape == (10, 6, 50)
  for name in ('eps_0', 'eps'):
    want = J.normal(keys[name], 6 * 3072).reshape(6, 32, 32, 3)
    got = d[name].cpu().numpy()
    assert got.shape == want.shape and np.abs(got - want).max() < 2e-6
  t = sample_t(d['t0'], 6, vdm.config)
  assert t.shape == (6,) and float(t.min()) >= 0 and float(t.max()) < 1


@pytest.mark.gpu
def test_cuda_rng_shapes_and_errors(cuda_device):
  from mulan_b200 import _lib, ops
  eps = ops.rng_normal((7, 9), (4, 32, 32, 3), cuda_device)
  assert eps.shape == (4, 32, 32, 3)
  assert torch.equal(eps.reshape(-1), ops.rng_normal((7, 9), (4 * 3072,), cuda_device))
  out = torch.empty(10, device=cuda_device)
  assert ops.rng_uniform((1, 2), (10,), out=out) is out
  with pytest.raises(_lib.MulanError):
    _lib.check(_lib.load().mulan_rng_normal(0, 0, 2 ** 32, None, None))
  with pytest.raises(_lib.MulanError):
    _lib.check(_lib.load().mulan_rng_uniform(0, 0, 4, 1.0, 0.0, out.data_ptr(), None))
