"""TEST INFRASTRUCTURE ONLY - numpy restatement of the JAX random draws VDM.__call__ makes
(SURVEY.md 8f "next" row 2).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import this file.

Reference call sites: ldm/model_mulan_epsilon.py:287-292 (t0 = jax.random.uniform(rng, ())),
:315 and :327 (eps_0, eps = jax.random.normal(rng, shape)), :213 (jax.random.gumbel).
The arithmetic lives in an UN-VENDORED third-party dependency: jax <= 0.4.23 + jaxlib
(README.md:26 of the reference; no lock file), default PRNG implementation `threefry2x32`,
`jax_threefry_partitionable=False`.  Published algorithm restated here:

  threefry2x32       Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3"
                     (SC'11), 20 rounds, rotation constants (13,15,26,6 | 17,29,16,24), key
                     schedule parity constant 0x1BD11BDA.
  random_bits        jax._src.prng.threefry_random_bits (non-partitionable): counters
                     iota(n) (one zero appended when n is odd) are split in two halves
                     (x0, x1); threefry2x32(key, x0, x1) -> concatenate(y0, y1)[:n].
  uniform            jax._src.random._uniform: (bits >> 9 | 0x3F800000) as float - 1, then
                     max(minval, f * (maxval - minval) + minval).
  normal             sqrt(2) * erf_inv(uniform(minval = nextafter(-1, 0), maxval = 1)).
  erf_inv (float32)  XLA's ErfInv: M. Giles, "Approximating the erfinv function" (2010),
                     w = -log1p(-x x); two degree-8 polynomials in w - 2.5 / sqrt(w) - 3.
  gumbel             -log(-log(uniform(minval = finfo.tiny, maxval = 1))).

Pinning.  JAX cannot be installed here, so the pins are published known-answer values
(tests/test_rng.py): the three Random123 threefry2x32 vectors that JAX's own test-suite uses,
and the outputs the JAX documentation prints - uniform(PRNGKey(0), ()) = 0.41845703,
normal(PRNGKey(0), ()) = -0.20584226, normal(PRNGKey(0), (3,)) =
[1.8160863, -0.48262316, 0.33988908], normal(PRNGKey(42), ()) = -0.18471177 - all
reproduced digit for digit.  What stays unpinned: Flax's `make_rng` key derivation (module
path hashing) - per-draw keys are therefore INPUTS of the kernels - and jax.random.gamma.
"""
from __future__ import annotations

import numpy as np

U32 = np.uint32
ROT = (13, 15, 26, 6, 17, 29, 16, 24)
PARITY = U32(0x1BD11BDA)

ERFINV_LT5 = (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087,
              -0.00125372503, -0.00417768164, 0.246640727, 1.50140941)
ERFINV_GE5 = (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773,
              -0.0076224613, 0.00943887047, 1.00167406, 2.83297682)


def prng_key(seed: int):
  """jax.random.PRNGKey(seed) for the threefry implementation: (high word, low word)."""
  seed = int(seed) & 0xFFFFFFFFFFFFFFFF
  return (seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF


def _rotl(x, r):
  return ((x << U32(r)) | (x >> U32(32 - r))).astype(U32)


def threefry2x32(key, x0, x1):
  """20-round Threefry-2x32 of the counter pair (x0, x1) under key = (k0, k1)."""
  with np.errstate(over='ignore'):
    k0, k1 = U32(key[0]), U32(key[1])
    ks = (k0, k1, U32(k0 ^ k1 ^ PARITY))
    x0 = (np.asarray(x0, U32) + ks[0]).astype(U32)
    x1 = (np.asarray(x1, U32) + ks[1]).astype(U32)
    for i in range(5):
      for r in (ROT[:4] if i % 2 == 0 else ROT[4:]):
        x0 = (x0 + x1).astype(U32)
        x1 = _rotl(x1, r)
        x1 = (x1 ^ x0).astype(U32)
      x0 = (x0 + ks[(i + 1) % 3]).astype(U32)
      x1 = (x1 + ks[(i + 2) % 3] + U32(i + 1)).astype(U32)
  return x0, x1


def random_bits(key, n: int):
  """threefry_random_bits(key, 32, (n,)): uint32 [n]."""
  half = (n + 1) // 2
  c = np.arange(n, dtype=U32)
  if n % 2:
    c = np.concatenate([c, np.zeros(1, U32)])
  y0, y1 = threefry2x32(key, c[:half], c[half:])
  return np.concatenate([y0, y1])[:n]


def uniform(key, n: int, minval=0.0, maxval=1.0):
  f32 = np.float32
  minval, maxval = f32(minval), f32(maxval)
  bits = random_bits(key, n)
  floats = ((bits >> U32(9)) | U32(0x3F800000)).view(f32) - f32(1)
  return np.maximum(minval, (floats * f32(maxval - minval)).astype(f32) + minval).astype(f32)


def erf_inv(x):
  f32 = np.float32
  x = np.asarray(x, f32)
  w = (-np.log1p((-x * x).astype(f32))).astype(f32)
  lt = w < 5
  w = np.where(lt, w - f32(2.5), np.sqrt(w) - f32(3)).astype(f32)
  p = np.where(lt, f32(ERFINV_LT5[0]), f32(ERFINV_GE5[0])).astype(f32)
  for a, b in zip(ERFINV_LT5[1:], ERFINV_GE5[1:]):
    p = (np.where(lt, f32(a), f32(b)) + (p * w).astype(f32)).astype(f32)
  with np.errstate(invalid='ignore'):
    return np.where(np.abs(x) == 1, x * f32(np.inf), (p * x).astype(f32)).astype(f32)


def normal(key, n: int):
  f32 = np.float32
  lo = np.nextafter(f32(-1), f32(0))
  return (f32(np.sqrt(2)) * erf_inv(uniform(key, n, lo, 1.0))).astype(f32)


def gumbel(key, n: int):
  f32 = np.float32
  u = uniform(key, n, np.finfo(f32).tiny, 1.0)
  return (-np.log(-np.log(u))).astype(f32)
