"""Host-side mirror of the reference's model interface for the hot path.

Same names, argument meaning and outputs as the reference's Flax modules, so a restated
training / evaluation loop stays drop-in:

  VDMConfig, VDMOutput                      ldm/model_vdm.py:33-92
  EncDec.encode                             ldm/model_vdm.py:274-280
  NoiseSchedule_polynomial_fixedend         ldm/model_mulan_epsilon.py:481-613
  VDM.__call__                              ldm/model_mulan_epsilon.py:280-363,
                                            ldm/model_mulan_velocity.py:188-268
  Experiment_VDM.loss_fn                    ldm/experiment_vdm.py:47-78

The encoder and the denoiser (U-Net) are NOT part of the path: they are
constructor arguments (any torch callables).  The five Dense layers of the schedule head are
cuBLAS GEMMs (torch.nn.Linear, float32 -- the reference runs with
JAX_DEFAULT_MATMUL_PRECISION=float32).  Everything between those networks runs in
libmulan_b200.so; there is no PyTorch fallback for it.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, NamedTuple, Optional

import numpy as np
import torch
from torch import nn

from . import ops
from ._lib import (MULAN_GT_MEAN, MULAN_GT_PIXEL, MULAN_PARAM_EPS, MULAN_PARAM_VEL,
                   MULAN_PARAM_VEL_FROM_EPS)


@dataclass
class VDMConfig:
  """The fields of ldm/model_vdm.py:33-82 that the hot path reads (same names)."""
  vocab_size: int = 256
  antithetic_time_sampling: bool = True
  gamma_type: str = 'poly_fixedend'
  gamma_min: float = -13.3
  gamma_max: float = 5.0
  sm_n_timesteps: int = 0
  latent_size: int = 50
  latent_type: str = 'topk'
  latent_k: int = 15
  topk_noise_type: str = 'gamma'
  reparam_type: str = 'true'
  z_conditioning: bool = True
  velocity_from_epsilon: bool = False
  unet_type: str = 'vdm'
  vdm_type: str = 'mulan_epsilon'     # ldm/experiment_vdm.py:32-38 selects the module by this


class VDMOutput(NamedTuple):
  """ldm/model_vdm.py:86-92."""
  loss_recon: torch.Tensor  # [B]
  loss_klz: torch.Tensor    # [B]
  loss_diff: torch.Tensor   # [B]
  var_0: torch.Tensor       # scalar
  var_1: torch.Tensor       # scalar


class EncDec:
  """ldm/model_vdm.py:265-303.  Only `encode` is needed outside the kernels (it feeds the
  encoder network); decode/logprob live inside mulan_fwd_pre."""

  def __init__(self, config: VDMConfig):
    self.config = config

  def encode(self, x: torch.Tensor) -> torch.Tensor:
    return 2 * ((x.to(torch.float32).round() + .5) / self.config.vocab_size) - 1


class NoiseSchedule_polynomial_fixedend(nn.Module):
  """ldm/model_mulan_epsilon.py:481-613: MLP 50 -> 3072 -> 3072 -> 3 x 3072 producing the
  per-pixel polynomial coefficients.  Parameter names follow the Flax module
  (dense_1, dense_2, dense_out_a/b/c); `dense_out_a` is zero-initialised (:495-500)."""

  def __init__(self, config: VDMConfig, n_features: int = 32 * 32 * 3):
    super().__init__()
    self.config = config
    n_out = 32 * 32 * 3
    self.dense_1 = nn.Linear(config.latent_size, n_features)
    self.dense_2 = nn.Linear(n_features, n_features)
    self.dense_out_a = nn.Linear(n_features, n_out)
    self.dense_out_b = nn.Linear(n_features, n_out)
    self.dense_out_c = nn.Linear(n_features, n_out)
    for lin in (self.dense_1, self.dense_2, self.dense_out_b, self.dense_out_c):
      nn.init.normal_(lin.weight, std=1.0 / math.sqrt(lin.in_features))  # flax lecun_normal
      nn.init.zeros_(lin.bias)
    nn.init.zeros_(self.dense_out_a.weight)
    nn.init.zeros_(self.dense_out_a.bias)

  def load_flax(self, params: dict):
    """params: {'dense_1/kernel': [in,out], 'dense_1/bias': [out], ...} (Flax layout)."""
    with torch.no_grad():
      for name in ('dense_1', 'dense_2', 'dense_out_a', 'dense_out_b', 'dense_out_c'):
        lin = getattr(self, name)
        lin.weight.copy_(torch.as_tensor(params[name + '/kernel']).t())
        lin.bias.copy_(torch.as_tensor(params[name + '/bias']))

  def _compute_coefficients_raw(self, embedding):
    """:531-536 and the GEMM of :537 -- everything but the `1e-3 + softplus` epilogue, which the
    kernels apply themselves when handed the pre-activation (MULAN_FLAG_C_RAW)."""
    h = nn.functional.silu(self.dense_1(embedding))
    h = nn.functional.silu(self.dense_2(h))
    return self.dense_out_a(h), self.dense_out_b(h), self.dense_out_c(h)

  def _compute_coefficients(self, embedding):
    """:531-538.  swish == SiLU; softplus == logaddexp(x, 0)."""
    a, b, c_raw = self._compute_coefficients_raw(embedding)
    return a, b, 1e-3 + nn.functional.softplus(c_raw)


def sample_t(t0: torch.Tensor, n_batch: int, config: VDMConfig) -> torch.Tensor:
  """ldm/model_mulan_epsilon.py:287-297.  antithetic_time_sampling: t from the SCALAR draw t0
  (jnp.arange with float arguments is np.arange in double, cast to float32); otherwise t0 is
  the [n_batch] uniform draw itself.  Both branches are discretised when sm_n_timesteps > 0
  (:294-297)."""
  if config.antithetic_time_sampling:
    ar = torch.from_numpy(np.arange(0., 1., step=1. / n_batch).astype(np.float32)).to(t0.device)
    t = torch.remainder(t0.to(torch.float32).reshape(()) + ar, 1.)
  else:
    t = t0.to(torch.float32).reshape(n_batch)
  T = config.sm_n_timesteps
  if T > 0:
    t = torch.ceil(t * T) / T
  return t.contiguous()


def gamma_draw(shape, k: int, generator: Optional[torch.Generator], device) -> torch.Tensor:
  """jax.random.gamma(key, 1/k, shape) stand-in (ldm/model_mulan_epsilon.py:222-223)."""
  conc = torch.full(shape, 1.0 / k, dtype=torch.float32, device=device)
  return torch._standard_gamma(conc, generator=generator)


class VDM(nn.Module):
  """MuLAN model wrapper: the epsilon (`vdm_type='mulan_epsilon'`) and velocity
  (`'mulan_velocity'`, optionally `velocity_from_epsilon`) parameterisations share this class.

  encoder_model(orig_f[B,32,32,3], deterministic) -> logits [B, latent_size]
  score_model(z_t[B,32,32,3], g_t ([B] or [B,32,32,3]), conditioning[B,latent], deterministic)
      -> network output [B,32,32,3]
  """

  def __init__(self, config: VDMConfig, encoder_model: Callable, score_model: Callable,
               gamma: Optional[nn.Module] = None):
    super().__init__()
    if config.gamma_type != 'poly_fixedend':
      raise NotImplementedError('only gamma_type="poly_fixedend" (both shipped configs) is on '
                                'the kernel path')
    if config.latent_type not in ('topk', 'gumbel', 'gaussian') or config.reparam_type != 'true':
      raise NotImplementedError('latent_type must be topk / gumbel / gaussian with '
                                'reparam_type="true"')
    if config.topk_noise_type not in ('gamma', 'gumbel'):
      raise NotImplementedError('topk_noise_type must be "gamma" or "gumbel"')
    self.config = config
    self.encdec = EncDec(config)
    self.encoder_model = encoder_model
    self.score_model = score_model
    self.gamma = gamma if gamma is not None else NoiseSchedule_polynomial_fixedend(config)
    if config.vdm_type == 'mulan_epsilon':
      param = MULAN_PARAM_EPS
    elif config.vdm_type == 'mulan_velocity':
      param = MULAN_PARAM_VEL_FROM_EPS if config.velocity_from_epsilon else MULAN_PARAM_VEL
    else:
      raise NotImplementedError(f'vdm_type={config.vdm_type!r} is not a MuLAN model')
    gt = MULAN_GT_MEAN if config.unet_type == 'vdm' else MULAN_GT_PIXEL
    self.desc = ops.Desc(dim=32 * 32 * 3, vocab=config.vocab_size, param=param, gt_mode=gt,
                         n_timesteps=config.sm_n_timesteps, gamma_min=config.gamma_min,
                         gamma_max=config.gamma_max)
    self._generator = None
    # fuse d loss_diff / d net into the forward pass of the post kernel when training
    self.fused_value_and_grad = True
    # hand the kernels the pre-activation of dense_out_c: they apply 1e-3 + softplus and its
    # derivative themselves (ldm/model_mulan_epsilon.py:537), saving the framework's
    # elementwise passes (8 B/sub-pixel forward, 12 B/sub-pixel backward)
    self.fused_softplus = True
    # programmatic dependent launch between the kernels of one step (MULAN_FLAG_PDL)
    self.pdl = True

  def apply_encoder(self, images_int):
    """ldm/model_mulan_epsilon.py:178-180."""
    return self.encoder_model(self.encdec.encode(images_int), True)

  @torch.no_grad()
  def apply_gamma(self, t, embedding=None):
    """ldm/model_mulan_epsilon.py:182-193 with the embedding supplied (None -> zeros, as the
    reference does for x_zero=None): per-pixel gamma(embedding, t), [B, 3072]."""
    dev = next(self.gamma.parameters()).device
    t = torch.as_tensor(t, dtype=torch.float32, device=dev).reshape(-1)
    B = t.shape[0]
    if embedding is None:
      embedding = torch.zeros((B, self.config.latent_size), dtype=torch.float32, device=dev)
    a, b, c = (v.contiguous() for v in self.gamma._compute_coefficients(embedding))
    pix = ops.Desc(dim=self.desc.dim, vocab=self.desc.vocab, param=self.desc.param,
                   gt_mode=MULAN_GT_PIXEL, gamma_min=self.desc.gamma_min,
                   gamma_max=self.desc.gamma_max)
    return ops.sample_gamma(pix, a, b, c, t.contiguous())

  def make_draws(self, n_batch: int, device, generator: Optional[torch.Generator] = None,
                 jax_keys: Optional[dict] = None):
    """The four make_rng('sample') draws of __call__, in the reference's order.

    jax_keys = {'t0': (k0, k1), 'eps_0': (k0, k1), 'eps': (k0, k1)}: the raw threefry keys
    Flax's make_rng would hand to jax.random.uniform / normal for those draws; they are then
    generated on the device exactly as JAX would (mulan_rng_uniform / mulan_rng_normal).  The
    latent noise G stays a torch draw (jax.random.gamma is not restated)."""
    g = generator
    cfg = self.config
    L = cfg.latent_size
    # :287-292: a scalar draw for antithetic sampling, one uniform per example otherwise
    t_shape = () if cfg.antithetic_time_sampling else (n_batch,)
    if jax_keys is None:
      t0 = torch.rand(t_shape, generator=g, device=device)
    if cfg.latent_type == 'topk' and cfg.topk_noise_type == 'gamma':
      G = gamma_draw((10, n_batch, L), cfg.latent_k, g, device)          # jax.random.gamma
    elif cfg.latent_type == 'gaussian':
      G = torch.randn((n_batch, L), generator=g, device=device)           # eps_z
    else:                                                                  # jax.random.gumbel
      u = torch.rand((n_batch, L), generator=g, device=device).clamp_min(1e-20)
      G = -torch.log(-torch.log(u))
    if jax_keys is not None:
      shape = (n_batch, 32, 32, 3)
      return dict(t0=ops.rng_uniform(jax_keys['t0'], t_shape, device=device), G=G,
                  eps_0=ops.rng_normal(jax_keys['eps_0'], shape, device=device),
                  eps=ops.rng_normal(jax_keys['eps'], shape, device=device))
    return dict(
        t0=t0, G=G,
        eps_0=torch.randn((n_batch, 32, 32, 3), generator=g, device=device),
        eps=torch.randn((n_batch, 32, 32, 3), generator=g, device=device))

  def _get_embedding_and_kl_z(self, orig_f, step, deterministic, noise):
    """ldm/model_mulan_epsilon.py:257-271; `noise` is the helper's single random draw."""
    cfg = self.config
    noise = noise.contiguous()
    if cfg.latent_type == 'topk':
      logits = self.encoder_model(orig_f, deterministic)
      if cfg.topk_noise_type == 'gamma':
        return ops.aux_topk(logits, noise, cfg.latent_k)
      return ops.aux_topk_add(logits, noise, cfg.latent_k)
    if cfg.latent_type == 'gumbel':
      logits = self.encoder_model(orig_f, deterministic)
      tau = max(0.5, math.exp(-0.00001 * float(step)))                   # :218
      return ops.aux_gumbel(logits, noise, tau)
    mu_z, var_z = self.encoder_model(orig_f, deterministic)               # gaussian
    return ops.aux_gaussian(mu_z, var_z, noise)

  def forward(self, images, labels=None, conditioning=None, step=0, deterministic: bool = True,
              draws: Optional[dict] = None, generator: Optional[torch.Generator] = None):
    cfg = self.config
    x = images.reshape(-1, 32, 32, 3)
    n_batch = x.shape[0]
    dev = x.device
    if draws is None:
      draws = self.make_draws(n_batch, dev, generator)
    D = 32 * 32 * 3
    if 't' in draws:
      # per-row times supplied ready-made (the batched dense-VLB driver tiles one image's
      # antithetic t over several images, dist.eval_bpd_dense_sampling)
      t = draws['t'].to(torch.float32).reshape(n_batch).contiguous()
    else:
      t = sample_t(draws['t0'], n_batch, cfg)

    x_u8 = x.to(torch.uint8).reshape(n_batch, D).contiguous()
    orig_f = self.encdec.encode(x)
    embedding, kl_z = self._get_embedding_and_kl_z(orig_f, step, deterministic, draws['G'])
    # a schedule head whose _compute_coefficients was replaced (tests, other heads) returns the
    # activated c; the built-in head can hand over the pre-activation instead
    raw = (self.fused_softplus and hasattr(self.gamma, '_compute_coefficients_raw')
           and '_compute_coefficients' not in vars(self.gamma)
           and (cfg.vocab_size & (cfg.vocab_size - 1)) == 0)
    if raw:
      a, b, c = self.gamma._compute_coefficients_raw(embedding)
    else:
      a, b, c = self.gamma._compute_coefficients(embedding)

    # eps_0 / eps drawn for FEWER rows than the batch are broadcast by row % noise_rows (the
    # dense-VLB driver evaluates several images per launch with ONE key's draws,
    # ldm/notebook_utils.py:178-185) instead of being tiled in HBM
    n_noise = draws['eps'].shape[0]
    if n_noise != n_batch and (n_batch % n_noise != 0 or draws['eps_0'].shape[0] != n_noise):
      raise ValueError(f'eps_0 / eps have {n_noise} rows for a batch of {n_batch}')
    # (programmatic dependent launch pays from ~2000 rows on; below, the early-resident CTAs of
    # the next kernel cost more than the launch gap they hide -- profiles/r2_latency.md)
    desc = self.desc.replace(c_raw=raw, pdl=self.pdl and n_batch >= 2048,
                             noise_rows=n_noise if n_noise != n_batch else 0)
    tape = ops.ElboTape(desc)
    z_t, g_net, loss_recon, klz_prior, var_sums, link = ops.mulan_pre(
        tape, x_u8, a, b, c, t, draws['eps_0'].reshape(n_noise, D).contiguous(),
        draws['eps'].reshape(n_noise, D).contiguous())
    cond = embedding if cfg.z_conditioning else conditioning[:, None]
    g_in = g_net if cfg.unet_type == 'vdm' else g_net.reshape(n_batch, 32, 32, 3)
    net = self.score_model(z_t.reshape(n_batch, 32, 32, 3), g_in, cond, deterministic)
    net_flat = net.reshape(n_batch, D)
    if torch.is_grad_enabled() and net_flat.requires_grad and self.fused_value_and_grad:
      # value-and-grad in one pass: the cotangent loss_fn will send is 1/(B*D*ln 2)
      # (ldm/experiment_vdm.py:62-66); any other upstream gradient is corrected per row
      hint = torch.full((n_batch,), 1.0 / (n_batch * D * math.log(2.0)), dtype=torch.float32,
                        device=dev)
      loss_diff = ops.mulan_post_fused(tape, net_flat, link, hint)
    else:
      loss_diff = ops.mulan_post(tape, net_flat, link)
    n = float(n_batch * D)
    return VDMOutput(loss_recon=loss_recon, loss_klz=kl_z + klz_prior, loss_diff=loss_diff,
                     var_0=var_sums[:, 0].sum() / n, var_1=var_sums[:, 1].sum() / n)


def _deterministic_embedding(model: VDM, batch_size: int, device):
  """_get_deterministic_embedding (ldm/model_mulan_epsilon.py:365-376): k leading ones for
  'topk', one_hot(1) for 'gumbel', zeros for 'gaussian'."""
  cfg = model.config
  e = torch.zeros((batch_size, cfg.latent_size), dtype=torch.float32, device=device)
  if cfg.latent_type == 'topk':
    e[:, :cfg.latent_k] = 1.0
  elif cfg.latent_type == 'gumbel':
    e[:, 1] = 1.0
  return e


@torch.no_grad()
def sample(model: VDM, i: int, T: int, z_t, conditioning=None, eps=None, generator=None,
           coeffs=None):
  """VDM.sample (ldm/model_mulan_epsilon.py:407-438, ldm/model_mulan_velocity.py:314-347): one
  ancestral step t=(T-i)/T -> s=(T-i-1)/T.  `coeffs` = cached (a, b, c) of the deterministic
  embedding ([1, D]: it is the same for every example and every step)."""
  cfg = model.config
  B = z_t.shape[0]
  dev = z_t.device
  D = 32 * 32 * 3
  if coeffs is None:
    coeffs = tuple(v.contiguous() for v in
                   model.gamma._compute_coefficients(_deterministic_embedding(model, 1, dev)))
  a, b, c = coeffs
  t = torch.full((B,), (T - i) / T, dtype=torch.float32, device=dev)
  s = torch.full((B,), (T - i - 1) / T, dtype=torch.float32, device=dev)
  if eps is None:
    eps = torch.randn((B, 32, 32, 3), generator=generator, device=dev)
  g_net = ops.sample_gamma(model.desc, a, b, c, t)
  embedding = _deterministic_embedding(model, B, dev)
  cond = embedding if cfg.z_conditioning else conditioning[:, None]
  g_in = g_net if cfg.unet_type == 'vdm' else g_net.reshape(B, 32, 32, 3)
  net = model.score_model(z_t.reshape(B, 32, 32, 3), g_in, cond, True)
  z_s = ops.sample_step(model.desc, a, b, c, t, s, z_t.reshape(B, D).contiguous(),
                        net.reshape(B, D).contiguous(), eps.reshape(B, D).contiguous())
  return z_s.reshape(B, 32, 32, 3)


@torch.no_grad()
def conditional_sample(model: VDM, i: int, T: int, z_t, embedding, conditioning=None, eps=None,
                       generator=None, coeffs=None):
  """VDM.conditional_sample (ldm/model_mulan_epsilon.py:377-406,
  ldm/model_mulan_velocity.py:281-312): one ancestral step with a GIVEN per-example embedding
  (the schedule, and with z_conditioning the denoiser, are conditioned on it).  `coeffs` =
  cached (a, b, c) of `embedding` ([B, D]; they do not change along the T steps)."""
  cfg = model.config
  B = z_t.shape[0]
  dev = z_t.device
  D = 32 * 32 * 3
  if coeffs is None:
    coeffs = tuple(v.contiguous() for v in model.gamma._compute_coefficients(embedding))
  a, b, c = coeffs
  t = torch.full((B,), (T - i) / T, dtype=torch.float32, device=dev)
  s = torch.full((B,), (T - i - 1) / T, dtype=torch.float32, device=dev)
  if eps is None:
    eps = torch.randn((B, 32, 32, 3), generator=generator, device=dev)
  g_net = ops.sample_gamma(model.desc, a, b, c, t)
  cond = embedding if cfg.z_conditioning else conditioning[:, None]
  g_in = g_net if cfg.unet_type == 'vdm' else g_net.reshape(B, 32, 32, 3)
  net = model.score_model(z_t.reshape(B, 32, 32, 3), g_in, cond, True)
  z_s = ops.sample_step(model.desc, a, b, c, t, s, z_t.reshape(B, D).contiguous(),
                        net.reshape(B, D).contiguous(), eps.reshape(B, D).contiguous())
  return z_s.reshape(B, 32, 32, 3)


@torch.no_grad()
def generate_x(model: VDM, z_0):
  """VDM.generate_x (ldm/model_mulan_epsilon.py:440-457), sample_softmax=False."""
  B = z_0.shape[0]
  x = ops.generate_x(model.desc, z_0.reshape(B, 32 * 32 * 3).contiguous())
  return x.reshape(B, 32, 32, 3)


@torch.no_grad()
def sample_fn(model: VDM, n: int, T: int = 1000, generator=None, sigma_prior: float = 1.0,
              device=None):
  """Experiment_VDM.sample_fn (ldm/experiment_vdm.py:80-110): z_init ~ N(0, sigma_prior^2),
  T ancestral steps, decode."""
  device = device or next(model.gamma.parameters()).device
  z = sigma_prior * torch.randn((n, 32, 32, 3), generator=generator, device=device)
  coeffs = tuple(v.contiguous() for v in
                 model.gamma._compute_coefficients(_deterministic_embedding(model, 1, device)))
  for i in range(T):
    z = sample(model, i, T, z, generator=generator, coeffs=coeffs)
  return generate_x(model, z)


def value_div_fn(model: VDM, x, embeddings, t, hutchinson_noise, high_precision: bool = False,
                 coeffs=None, out=None):
  """VDM.reverse_ode (ldm/model_mulan_epsilon.py:459-478) and its Hutchinson divergence
  (_get_value_div_fn, ldm/notebook_utils.py:204-216): -> (drift[B,32,32,3], div[B]).
  The denoiser's Jacobian-vector product comes from torch autograd through `score_model`;
  everything else is mulan_sample_gamma / mulan_ode_drift / mulan_row_dot.

  hutchinson_noise=None: drift only (the ODE sampler discards the divergence,
  notebook_utils.py:419), div is None.  coeffs = cached (a, b, c) of `embeddings` (they do not
  change along an ODE solve); out = (drift_out[B,D], div_out[B]) buffers to write into."""
  cfg = model.config
  B = x.shape[0]
  D = 32 * 32 * 3
  drift_out, div_out = out if out is not None else (None, None)
  with torch.no_grad():
    if coeffs is None:
      coeffs = tuple(q.contiguous() for q in model.gamma._compute_coefficients(embeddings))
    a, b, c = coeffs
    if torch.is_tensor(t) and t.numel() == B:      # per-row t (the reference passes vec_t)
      tt = t.to(device=x.device, dtype=torch.float32).reshape(B).contiguous()
    else:
      tt = torch.full((B,), float(t), dtype=torch.float32, device=x.device)
    g_net = ops.sample_gamma(model.desc, a, b, c, tt)
  g_in = g_net if cfg.unet_type == 'vdm' else g_net.reshape(B, 32, 32, 3)
  if hutchinson_noise is None:
    with torch.no_grad():
      x4 = x.detach().reshape(B, 32, 32, 3)
      net = model.score_model(x4, g_in, embeddings, True)
      drift = ops.ode_drift(model.desc, a, b, c, tt, x4.reshape(B, D).contiguous(),
                            net.reshape(B, D).contiguous(), None, high_precision, out=drift_out)
    return drift.reshape(B, 32, 32, 3), None
  xg = x.detach().reshape(B, 32, 32, 3).requires_grad_(True)
  with torch.enable_grad():
    net = model.score_model(xg, g_in, embeddings, True)
  v = hutchinson_noise.reshape(B, D).contiguous()
  drift, net_bar, div_direct = ops.ode_drift(
      model.desc, a, b, c, tt, xg.detach().reshape(B, D).contiguous(),
      net.detach().reshape(B, D).contiguous(), v, high_precision, out=drift_out)
  (x_bar,) = torch.autograd.grad(net, xg, grad_outputs=net_bar.reshape(net.shape))
  div = ops.row_dot(x_bar.reshape(B, D).contiguous(), v, add=div_direct, out=div_out)
  return drift.reshape(B, 32, 32, 3), div


def loss_fn(model: VDM, inputs: dict, step=0, is_train: bool = True, draws=None,
            generator=None):
  """Experiment_VDM.loss_fn (ldm/experiment_vdm.py:47-78): -> (bpd, metrics)."""
  outputs = model(**inputs, step=step, deterministic=not is_train, draws=draws,
                  generator=generator)
  rescale_to_bpd = 1. / (np.prod(inputs['images'].shape[1:]) * np.log(2.))
  bpd_latent = torch.mean(outputs.loss_klz) * rescale_to_bpd
  bpd_recon = torch.mean(outputs.loss_recon) * rescale_to_bpd
  bpd_diff = torch.mean(outputs.loss_diff) * rescale_to_bpd
  bpd = bpd_recon + bpd_latent + bpd_diff
  scalar_dict = {'bpd': bpd, 'bpd_latent': bpd_latent, 'bpd_recon': bpd_recon,
                 'bpd_diff': bpd_diff, 'var0': outputs.var_0, 'var': outputs.var_1}
  metrics = {'scalars': scalar_dict, 'images': {'inputs': inputs['images']}}
  return bpd, metrics
