"""ctypes binding of libmulan_b200.so (include/mulan_b200.h).

There is NO fallback: if the shared library is missing this module raises at first use, and
the ops refuse CPU tensors.  The product path never touches ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / 'libmulan_b200.so'

MULAN_PARAM_EPS = 0
MULAN_PARAM_VEL = 1
MULAN_PARAM_VEL_FROM_EPS = 2
MULAN_GT_MEAN = 0
MULAN_GT_PIXEL = 1
MULAN_FLAG_C_RAW = 1     # c is the pre-activation of dense_out_c (kernels apply 1e-3 + softplus)
MULAN_FLAG_PDL = 2       # programmatic dependent launch
MULAN_ABI_VERSION = 2
MULAN_RK45_SCRATCH = 2048
MULAN_SUMSQ_SCRATCH = 2048

PARAM_NAMES = {'eps': MULAN_PARAM_EPS, 'vel': MULAN_PARAM_VEL,
               'vel_from_eps': MULAN_PARAM_VEL_FROM_EPS}


class MulanDesc(C.Structure):
  """struct mulan_desc."""
  _fields_ = [('rows', C.c_int32), ('dim', C.c_int32), ('vocab', C.c_int32),
              ('param', C.c_int32), ('gt_mode', C.c_int32), ('n_timesteps', C.c_int32),
              ('gamma_min', C.c_double), ('gamma_max', C.c_double),
              ('flags', C.c_uint32), ('noise_rows', C.c_int32)]


class MulanEndConsts(C.Structure):
  """struct mulan_end_consts: the fixed-end transcendental constants as the caller's framework
  evaluates them in float32 (mulan_fwd_pre_consts)."""
  _fields_ = [('exp_half_g0', C.c_float), ('exp_neg_half_g0', C.c_float),
              ('sigmoid_g0', C.c_float), ('sigmoid_g1', C.c_float),
              ('log_sigmoid_g1', C.c_float)]


class MulanAdamwDesc(C.Structure):
  """struct mulan_adamw_desc."""
  _fields_ = [('n', C.c_int64), ('n_decay', C.c_int64), ('step', C.c_int32),
              ('reserved', C.c_int32), ('lr', C.c_double), ('b1', C.c_double), ('b2', C.c_double),
              ('eps', C.c_double), ('weight_decay', C.c_double), ('ema_rate', C.c_double),
              ('grad_scale', C.c_double), ('clip_norm', C.c_double),
              ('grad_sumsq', C.c_void_p)]


MULAN_PEER_HANDLE_BYTES = 64
MULAN_PEER_MAX = 8
MULAN_PEER_FLAG_WORDS = 32
MULAN_PEER_FLAG_ERR = 17


class MulanPeerDesc(C.Structure):
  """struct mulan_peer_desc."""
  _fields_ = [('world', C.c_int32), ('rank', C.c_int32),
              ('grads', C.c_void_p * MULAN_PEER_MAX), ('params', C.c_void_p * MULAN_PEER_MAX),
              ('flags', C.c_void_p * MULAN_PEER_MAX), ('epoch', C.c_uint32),
              ('reserved', C.c_uint32), ('mc_grads', C.c_void_p), ('mc_params', C.c_void_p)]


class MulanScalarBoard(C.Structure):
  """struct mulan_scalar_board."""
  _fields_ = [('world', C.c_int32), ('rank', C.c_int32), ('boards', C.c_void_p * MULAN_PEER_MAX)]


class MulanError(RuntimeError):
  def __init__(self, status: int, msg: str):
    super().__init__(f'libmulan_b200 status {status}: {msg}')
    self.status = status


DENOISER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                          C.c_void_p)

_P = C.c_void_p
_D = C.POINTER(MulanDesc)

# name -> argtypes; every symbol include/mulan_b200.h declares.
SIGNATURES = {
    'mulan_last_error': ([], C.c_char_p),
    'mulan_abi_version': ([], C.c_int),
    'mulan_kernel_param': ([C.c_int32], C.c_int),
    'mulan_fwd_pre': ([_D] + [_P] * 14, C.c_int),
    'mulan_fwd_pre_variant': ([_D], C.c_int),
    'mulan_fwd_pre_keyed': ([_D] + [_P] * 16, C.c_int),
    'mulan_fwd_pre_consts': ([_D, _P] + [_P] * 14, C.c_int),
    'mulan_host_end_consts': ([_D, _P], C.c_int),
    'mulan_fwd_pre_variant_consts': ([_D, _P], C.c_int),
    'mulan_fwd_post': ([_D] + [_P] * 10, C.c_int),
    'mulan_bwd_post': ([_D] + [_P] * 11, C.c_int),
    'mulan_fwd_bwd_post': ([_D] + [_P] * 12, C.c_int),
    'mulan_scale_rows': ([C.c_int32] * 2 + [_P] * 4, C.c_int),
    'mulan_bwd_pre': ([_D] + [_P] * 14, C.c_int),
    'mulan_aux_topk_fwd': ([C.c_int32] * 3 + [_P] * 5, C.c_int),
    'mulan_aux_topk_bwd': ([C.c_int32] * 3 + [_P] * 6, C.c_int),
    'mulan_aux_topk_add_fwd': ([C.c_int32] * 3 + [_P] * 5, C.c_int),
    'mulan_aux_topk_add_bwd': ([C.c_int32] * 3 + [_P] * 6, C.c_int),
    'mulan_aux_gumbel_fwd': ([C.c_int32] * 2 + [C.c_double] + [_P] * 5, C.c_int),
    'mulan_aux_gumbel_bwd': ([C.c_int32] * 2 + [C.c_double] + [_P] * 6, C.c_int),
    'mulan_aux_gaussian_fwd': ([C.c_int32] * 2 + [_P] * 6, C.c_int),
    'mulan_aux_gaussian_bwd': ([C.c_int32] * 2 + [_P] * 8, C.c_int),
    'mulan_bpd_reduce': ([_D] + [_P] * 9, C.c_int),
    'mulan_reduce_ws_bytes': ([C.c_int32], C.c_size_t),
    'mulan_post_bpd': ([_D] + [_P] * 19, C.c_int),
    'mulan_post_bpd_peer': ([_D] + [_P] * 20, C.c_int),
    'mulan_scalar_board_bytes': ([], C.c_size_t),
    'mulan_scalar_board_read': ([_P, _P, _P, _P], C.c_int),
    'mulan_elbo_host': ([_D] + [_P] * 8 + [DENOISER_FN, _P, C.c_int32] + [_P] * 6, C.c_int),
    'mulan_elbo_host_keyed': ([_D] + [_P] * 8 + [DENOISER_FN, _P, C.c_int32] + [_P] * 6, C.c_int),
    'mulan_sample_gamma': ([_D, C.c_int32] + [_P] * 6, C.c_int),
    'mulan_sample_step': ([_D, C.c_int32] + [_P] * 10, C.c_int),
    'mulan_generate_x': ([_D] + [_P] * 3, C.c_int),
    'mulan_ode_drift': ([_D, C.c_int32] + [_P] * 7 + [C.c_int32] + [_P] * 4, C.c_int),
    'mulan_row_dot': ([C.c_int32] * 2 + [_P] * 5, C.c_int),
    'mulan_rk45_stage': ([C.c_int64, C.c_int32, _P, C.c_double, _P, _P, C.c_int64, _P, _P, _P],
                         C.c_int),
    'mulan_rk45_norm': ([C.c_int64, C.c_int32, _P, C.c_double, C.c_double, C.c_double, _P, _P, _P,
                         C.c_int64, C.c_int32, _P, _P, _P], C.c_int),
    'mulan_adamw_ema': ([C.POINTER(MulanAdamwDesc)] + [_P] * 6, C.c_int),
    'mulan_peer_alloc': ([C.c_size_t, C.POINTER(C.c_void_p), _P], C.c_int),
    'mulan_peer_open': ([_P, C.POINTER(C.c_void_p)], C.c_int),
    'mulan_peer_close': ([_P], C.c_int),
    'mulan_peer_free': ([_P], C.c_int),
    'mulan_adamw_ema_peer': ([C.POINTER(MulanAdamwDesc), C.POINTER(MulanPeerDesc), C.c_int64,
                              C.c_int64, _P, _P, _P, _P], C.c_int),
    'mulan_rng_bits': ([C.c_uint32, C.c_uint32, C.c_int64, _P, _P], C.c_int),
    'mulan_rng_uniform': ([C.c_uint32, C.c_uint32, C.c_int64, C.c_float, C.c_float, _P, _P],
                          C.c_int),
    'mulan_rng_normal': ([C.c_uint32, C.c_uint32, C.c_int64, _P, _P], C.c_int),
    'mulan_grad_sumsq': ([C.c_int64, _P, _P, _P, _P], C.c_int),
    'mulan_host_workspace_release': ([], None),
}

# XLA legacy custom-call targets (include/mulan_b200_xla.h):
#   void target(stream, void** buffers, const char* opaque, size_t opaque_len, status)
_XLA_TARGET = ([_P, C.POINTER(_P), C.c_char_p, C.c_size_t, _P], None)
XLA_SIGNATURES = {name: _XLA_TARGET for name in (
    'mulan_xla_fwd_pre', 'mulan_xla_fwd_post', 'mulan_xla_bwd_post', 'mulan_xla_fwd_bwd_post',
    'mulan_xla_bwd_pre', 'mulan_xla_bpd_reduce', 'mulan_xla_aux_topk_fwd',
    'mulan_xla_aux_topk_bwd')}


class MulanXlaOpaque(C.Structure):
  """mulan_xla_opaque: the custom call's backend_config bytes."""
  _fields_ = [('desc', MulanDesc), ('absent_mask', C.c_uint32), ('reserved', C.c_uint32)]


class MulanXlaAuxOpaque(C.Structure):
  _fields_ = [('rows', C.c_int32), ('latent', C.c_int32), ('k', C.c_int32),
              ('absent_mask', C.c_uint32)]


_lib = None


def load() -> C.CDLL:
  """dlopen the in-tree library; loud failure when it has not been built."""
  global _lib
  if _lib is None:
    if not LIB_PATH.exists():
      raise ImportError(
          f'{LIB_PATH} is missing: build it with `python -m mulan_b200.build` '
          '(or __graft_entry__.build()). There is no CPU / PyTorch fallback.')
    lib = C.CDLL(str(LIB_PATH))
    for name, (argtypes, restype) in {**SIGNATURES, **XLA_SIGNATURES}.items():
      fn = getattr(lib, name)
      fn.argtypes = argtypes
      fn.restype = restype
    _lib = lib
  return _lib


def check(status: int) -> None:
  if status != 0:
    raise MulanError(status, load().mulan_last_error().decode())


def make_desc(rows: int, dim: int = 3072, vocab: int = 256, param: int = MULAN_PARAM_EPS,
              gt_mode: int = MULAN_GT_MEAN, n_timesteps: int = 0,
              gamma_min: float = -13.3, gamma_max: float = 5.0, flags: int = 0,
              noise_rows: int = 0) -> MulanDesc:
  return MulanDesc(rows, dim, vocab, param, gt_mode, n_timesteps, gamma_min, gamma_max, flags,
                   noise_rows)


def kernel_param(param: int) -> int:
  """mulan_kernel_param: the loss formula the post / bwd_pre kernels run for `param`
  (velocity_from_epsilon evaluates the algebraically identical epsilon form)."""
  return int(load().mulan_kernel_param(int(param)))
