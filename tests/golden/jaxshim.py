"""A minimal torch-backed stand-in for the `jax` / `flax` / `chex` modules, just large enough
to IMPORT AND EXECUTE the reference's own source files
(/root/reference/ldm/model_vdm.py, model_mulan_epsilon.py, model_mulan_velocity.py) on CPU.

Used only by tests/golden/make_golden.py to generate golden vectors in the build container
(JAX itself is not installable here).  Nothing from the reference is copied: the reference
modules are imported from /root/reference at generation time.

What it pins: the reference's expression order, broadcasting, reductions, jvp structure and
RNG call order, evaluated op-by-op in float32 (or float64) by torch-CPU kernels.
What it cannot pin: XLA's own roundings of exp/log/pow.

Random draws are not generated: `jax.random.*` pop pre-supplied arrays from a queue in call
order (the reference calls make_rng('sample') for t0, the gamma noise, eps_0, eps).
"""
from __future__ import annotations

import dataclasses
import sys
import types

import numpy as np
import torch

_DRAWS = []          # queue consumed by jax.random.*


def set_draws(draws):
  _DRAWS[:] = list(draws)


def _pop(kind, shape=None):
  if not _DRAWS:
    raise RuntimeError(f'jaxshim: no draw left for jax.random.{kind}')
  k, v = _DRAWS.pop(0)
  assert k == kind, f'jaxshim: expected a {k} draw, reference asked for {kind}'
  v = torch.as_tensor(v).to(torch.get_default_dtype())
  if shape is not None:
    assert tuple(v.shape) == tuple(shape), (kind, v.shape, shape)
  return v


def _t(x):
  if isinstance(x, torch.Tensor):
    return x
  return torch.as_tensor(x, dtype=torch.get_default_dtype()
                         if isinstance(x, float) or (isinstance(x, np.ndarray) and x.dtype.kind == 'f')
                         or (isinstance(x, (list, tuple)) and any(isinstance(e, float) for e in x))
                         else None)


_DT = {'int32': torch.int32, 'uint8': torch.uint8, 'float32': torch.float32,
       'float64': torch.float64, float: None, int: torch.int64}


def _patch_tensor():
  """jax-array methods the reference uses that torch.Tensor lacks or spells differently."""
  if getattr(torch.Tensor, '_jaxshim', False):
    return
  _round, _transpose = torch.Tensor.round, torch.Tensor.transpose

  def round_(self, *a, **k):          # ints: no-op, like jnp
    return self if not self.is_floating_point() else _round(self, *a, **k)

  def astype(self, dt):
    td = _DT.get(dt, dt)
    return self.to(torch.get_default_dtype() if td is None else td)

  def transpose(self, *a):
    if len(a) == 1 and isinstance(a[0], (list, tuple)):
      return self.permute(*a[0])
    return _transpose(self, *a)

  torch.Tensor.round = round_
  torch.Tensor.astype = astype
  torch.Tensor.transpose = transpose
  torch.Tensor._jaxshim = True


class _Loose(types.ModuleType):
  """Module whose unknown attributes are permissive callables (for import-time decorators,
  initializers etc. that the executed path never evaluates)."""

  def __getattr__(self, name):
    if name.startswith('__'):
      raise AttributeError(name)
    def stub(*a, **k):
      if len(a) == 1 and callable(a[0]) and not k:
        return a[0]
      return stub
    stub.__name__ = name
    return stub


def _axis(k):
  ax = k.pop('axis', None)
  if isinstance(ax, list):
    ax = tuple(ax)
  return ax


def _make_jnp():
  m = _Loose('jax.numpy')
  m.float32, m.ndarray = torch.float32, torch.Tensor
  m.isscalar = lambda x: np.isscalar(x)
  m.zeros = lambda shape, dtype=None: torch.zeros(shape, dtype=_DT.get(dtype, dtype))
  m.ones = lambda shape, dtype=None: torch.ones(shape, dtype=_DT.get(dtype, dtype))
  m.zeros_like, m.ones_like = torch.zeros_like, torch.ones_like
  m.array = lambda v, dtype=None: _t(v)
  m.asarray = m.array

  def arange(start, stop=None, step=None, dtype=None):
    # jnp.arange with float arguments falls back to np.arange (double) + cast
    v = np.arange(start, stop, step) if stop is not None else np.arange(start)
    return torch.from_numpy(v).to(torch.get_default_dtype() if v.dtype.kind == 'f' else torch.int64)
  m.arange = arange
  m.mod = lambda x, y: torch.remainder(_t(x), y)
  m.ceil, m.exp, m.log, m.sqrt, m.square, m.expm1 = (
      lambda x: torch.ceil(_t(x)), lambda x: torch.exp(_t(x)), lambda x: torch.log(_t(x)),
      lambda x: torch.sqrt(_t(x)), lambda x: torch.square(_t(x)), lambda x: torch.expm1(_t(x)))
  m.maximum = lambda x, y: torch.maximum(_t(x), _t(y))
  m.where = lambda c, x, y: torch.where(c, x, y)

  def _red(fn):
    def f(x, **k):
      ax = _axis(k)
      kd = k.pop('keepdims', False)
      return fn(x) if ax is None else fn(x, dim=ax, keepdim=kd)
    return f
  m.sum, m.mean = _red(torch.sum), _red(torch.mean)
  m.reshape = lambda x, shape: x.reshape(shape)
  m.repeat = lambda x, n, axis: torch.repeat_interleave(x, n, dim=axis)
  m.tile = lambda x, reps: x.repeat(*reps) if isinstance(reps, (tuple, list)) else x.repeat(reps)
  m.concatenate = lambda xs, axis=0: torch.cat(list(xs), dim=axis)
  m.argmax = lambda x, axis=None: torch.argmax(x, dim=axis)
  linalg = _Loose('jax.numpy.linalg')
  linalg.norm = lambda x, ord=None, axis=None, keepdims=False: torch.sqrt(
      torch.sum(torch.abs(x) ** 2, dim=axis, keepdim=keepdims))
  m.linalg = linalg
  return m


def _sigmoid(x):        # jax.nn.sigmoid == lax.logistic == 1/(1+exp(-x))
  return 1.0 / (1.0 + torch.exp(-x))


def _log_softmax(x, axis=-1):
  shifted = x - x.max(dim=axis, keepdim=True).values.detach()
  return shifted - torch.log(torch.sum(torch.exp(shifted), dim=axis, keepdim=True))


def _softmax(x, axis=-1):
  un = torch.exp(x - x.max(dim=axis, keepdim=True).values.detach())
  return un / torch.sum(un, dim=axis, keepdim=True)


def _softplus(x):       # jnp.logaddexp(x, 0)
  return torch.logaddexp(x, torch.zeros((), dtype=x.dtype))


def _make_jax(jnp):
  jax = _Loose('jax')
  jax.numpy = jnp
  nn = _Loose('jax.nn')
  nn.sigmoid, nn.softmax, nn.log_softmax, nn.softplus = _sigmoid, _softmax, _log_softmax, _softplus
  nn.one_hot = lambda x, n: torch.nn.functional.one_hot(x.long(), n).to(torch.get_default_dtype())
  nn.initializers = _Loose('jax.nn.initializers')
  jax.nn = nn
  lax = _Loose('jax.lax')
  lax.stop_gradient = lambda x: x.detach()
  lax.top_k = lambda x, k: tuple(torch.topk(x, k, dim=-1))
  lax.Precision = None
  jax.lax = lax
  rnd = _Loose('jax.random')
  rnd.uniform = lambda key, shape=(): _pop('uniform', shape)
  rnd.normal = lambda key, shape=(): _pop('normal', shape)
  rnd.gamma = lambda key, a, shape=(): _pop('gamma', shape)
  rnd.gumbel = lambda key, shape=(): _pop('gumbel', shape)
  jax.random = rnd

  def jvp(fn, primals, tangents):
    return torch.func.jvp(fn, tuple(primals), tuple(tangents))
  jax.jvp = jvp
  jax.checkpoint = lambda f, **k: f
  return jax


class Module:
  """flax.linen.Module stand-in: annotated class attributes become constructor fields,
  `setup()` runs at construction, `make_rng` returns a dummy key."""

  def __init_subclass__(cls, **kw):
    super().__init_subclass__(**kw)
    fields = []
    for klass in reversed(cls.__mro__):
      for n in getattr(klass, '__annotations__', {}):
        if n not in fields and n not in ('name', 'parent'):
          fields.append(n)
    cls._fields = fields

  def __init__(self, *args, **kwargs):
    kwargs.pop('parent', None)
    self.name = kwargs.pop('name', None)
    for n, v in zip(self._fields, args):
      setattr(self, n, v)
    for n in self._fields[len(args):]:
      if n in kwargs:
        setattr(self, n, kwargs.pop(n))
      elif not hasattr(type(self), n):
        raise TypeError(f'{type(self).__name__}: missing field {n}')
    for n, v in kwargs.items():          # e.g. Dense(kernel_init=...)
      setattr(self, n, v)
    if hasattr(self, 'setup'):
      self.setup()

  def make_rng(self, name):
    return ('rng', name)


class Dense(Module):
  """flax nn.Dense: x @ kernel + bias; kernel [in, out] / bias assigned by the generator."""
  features: int
  use_bias: bool = True

  def __call__(self, x):
    y = x @ self.kernel
    return y + self.bias if self.use_bias else y


def _make_flax():
  flax = _Loose('flax')
  struct = _Loose('flax.struct')
  struct.dataclass = dataclasses.dataclass
  flax.struct = struct
  nn = _Loose('flax.linen')
  nn.Module, nn.Dense = Module, Dense
  nn.compact = lambda f: f
  nn.remat = lambda f=None, **k: f
  nn.sigmoid, nn.softplus = _sigmoid, _softplus
  nn.swish = lambda x: x * _sigmoid(x)
  nn.relu = torch.relu
  nn.initializers = _Loose('flax.linen.initializers')
  nn.normalization = _Loose('flax.linen.normalization')
  flax.linen = nn
  return flax, nn, struct


def install(reference_root='/root/reference'):
  """Put the stand-ins into sys.modules and the reference root on sys.path."""
  _patch_tensor()
  jnp = _make_jnp()
  jax = _make_jax(jnp)
  flax, nn, struct = _make_flax()
  chex = _Loose('chex')
  chex.Array = torch.Tensor
  for name, mod in (('jax', jax), ('jax.numpy', jnp), ('jax.nn', jax.nn), ('jax.lax', jax.lax),
                    ('jax.random', jax.random), ('flax', flax), ('flax.linen', nn),
                    ('flax.struct', struct), ('chex', chex)):
    sys.modules[name] = mod
  if reference_root not in sys.path:
    sys.path.insert(0, reference_root)
  return jax, jnp, nn
