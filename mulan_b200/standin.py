"""Library-op STAND-INS for the two networks that sit on either side of the hot path.

NOT part of the path (SURVEY.md section 8 keeps the U-Net and the encoder on "the framework's
library path"): these are plain torch.nn restatements on cuDNN / cuBLAS, float32, of

  model_vdm.ScoreUNet          ldm/model_vdm.py:309-389   (ResnetBlock :610-658, AttnBlock
                               :660-702, Base2FourierFeatures :812-829,
                               get_timestep_embedding :391-413)
  model_mulan_epsilon.UnetEncoder   ldm/model_mulan_epsilon.py:103-154

so that the train-step workloads of BASELINE.json configs[1..3] ("cifar10-conditioned /
imagenet32 train step") can be timed end to end around the CUDA kernels with the same
FLOP count, layer structure, zero-initialised layers and parameter count as the reference
(30.6 M / 122 M score parameters for sm_n_embd 128 / 256).  No hand-written kernel here.

Layout: the kernels' [B,32,32,3] NHWC tensors are viewed as NCHW with channels_last strides
(no copy); cuDNN runs channels_last natively.
"""
from __future__ import annotations

import math

import torch
from torch import nn
from torch.nn import functional as F


def get_timestep_embedding(timesteps, embedding_dim: int):
  """ldm/model_vdm.py:391-413."""
  timesteps = timesteps * 1000.
  half = embedding_dim // 2
  emb = math.log(10000) / (half - 1)
  emb = torch.exp(torch.arange(half, dtype=torch.float32, device=timesteps.device) * -emb)
  emb = timesteps.float()[:, None] * emb[None, :]
  emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)
  if embedding_dim % 2 == 1:
    emb = F.pad(emb, (0, 1))
  return emb


class Base2FourierFeatures(nn.Module):
  """ldm/model_vdm.py:812-829 on NCHW: per input channel, sin/cos at 2^f * 2 pi."""

  def __init__(self, start=6, stop=8, step=1):
    super().__init__()
    self.freqs = list(range(start, stop, step))

  def forward(self, x):                     # x [B, C, H, W]
    w = torch.tensor([2. ** f * 2 * math.pi for f in self.freqs], dtype=x.dtype, device=x.device)
    h = x.repeat_interleave(len(self.freqs), dim=1) * w.repeat(x.shape[1])[None, :, None, None]
    return torch.cat([torch.sin(h), torch.cos(h)], dim=1)


class ResnetBlock(nn.Module):
  """ldm/model_vdm.py:610-658 (GroupNorm(32) -> swish -> conv3x3 -> +cond -> GroupNorm ->
  swish -> dropout -> conv3x3(zero init); nin_shortcut when channels change)."""

  def __init__(self, in_ch, out_ch, cond_ch, pdrop):
    super().__init__()
    self.norm1 = nn.GroupNorm(32, in_ch, eps=1e-6)
    self.conv1 = nn.Conv2d(in_ch, out_ch, 3, padding=1)
    self.cond_proj = nn.Linear(cond_ch, out_ch, bias=False)
    self.norm2 = nn.GroupNorm(32, out_ch, eps=1e-6)
    self.drop = nn.Dropout(pdrop)
    self.conv2 = nn.Conv2d(out_ch, out_ch, 3, padding=1)
    nn.init.zeros_(self.cond_proj.weight)
    nn.init.zeros_(self.conv2.weight)
    nn.init.zeros_(self.conv2.bias)
    self.nin_shortcut = nn.Conv2d(in_ch, out_ch, 1) if in_ch != out_ch else None

  def forward(self, x, cond):
    h = self.conv1(F.silu(self.norm1(x)))
    h = h + self.cond_proj(cond)[:, :, None, None]
    h = self.conv2(self.drop(F.silu(self.norm2(h))))
    if self.nin_shortcut is not None:
      x = self.nin_shortcut(x)
    return x + h


class AttnBlock(nn.Module):
  """ldm/model_vdm.py:660-702, single head over the 32x32 positions."""

  def __init__(self, ch):
    super().__init__()
    self.norm = nn.GroupNorm(32, ch, eps=1e-6)
    self.q, self.k, self.v = nn.Conv2d(ch, ch, 1), nn.Conv2d(ch, ch, 1), nn.Conv2d(ch, ch, 1)
    self.proj_out = nn.Conv2d(ch, ch, 1)
    nn.init.zeros_(self.proj_out.weight)
    nn.init.zeros_(self.proj_out.bias)

  def forward(self, x):
    B, C, H, W = x.shape
    h = self.norm(x)
    q, k, v = (m(h).reshape(B, 1, C, H * W).transpose(2, 3) for m in (self.q, self.k, self.v))
    h = F.scaled_dot_product_attention(q, k, v)             # [B, 1, HW, C]
    h = h.transpose(2, 3).reshape(B, C, H, W)
    return x + self.proj_out(h)


class _UNetTrunk(nn.Module):
  """conditioning MLP + conv_in + down blocks + middle (shared by ScoreUNet and UnetEncoder)."""

  def __init__(self, n_embd, n_down, cond_in, pdrop, with_fourier):
    super().__init__()
    self.n_embd = n_embd
    self.dense0 = nn.Linear(n_embd + cond_in, n_embd * 4)
    self.dense1 = nn.Linear(n_embd * 4, n_embd * 4)
    self.fourier = Base2FourierFeatures(6, 8, 1) if with_fourier else None
    in_ch = 3 + (3 * 2 * 2 if with_fourier else 0)
    self.conv_in = nn.Conv2d(in_ch, n_embd, 3, padding=1)
    c4 = n_embd * 4
    self.down = nn.ModuleList([ResnetBlock(n_embd, n_embd, c4, pdrop) for _ in range(n_down)])
    self.mid1 = ResnetBlock(n_embd, n_embd, c4, pdrop)
    self.mid_attn = AttnBlock(n_embd)
    self.mid2 = ResnetBlock(n_embd, n_embd, c4, pdrop)

  def trunk(self, z_nchw, t01, conditioning):
    temb = get_timestep_embedding(t01, self.n_embd)
    cond = torch.cat([temb, conditioning], dim=1)
    cond = F.silu(self.dense1(F.silu(self.dense0(cond))))
    h = z_nchw
    if self.fourier is not None:
      h = torch.cat([h, self.fourier(h)], dim=1)
    h = self.conv_in(h)
    hs = [h]
    for blk in self.down:
      h = blk(hs[-1], cond)
      hs.append(h)
    h = self.mid2(self.mid_attn(self.mid1(hs[-1], cond)), cond)
    return h, hs, cond


class ScoreUNet(_UNetTrunk):
  """ldm/model_vdm.py:309-389.  forward(z[B,32,32,3], g_t[B], conditioning[B,L], deterministic)
  -> [B,32,32,3]; zero-initialised conv_out, so the fresh network returns z (:378-386)."""

  def __init__(self, n_embd=128, n_layer=32, pdrop=0.1, latent_size=50, gamma_min=-13.3,
               gamma_max=5.0, with_fourier=True):
    super().__init__(n_embd, n_layer, latent_size, pdrop, with_fourier)
    self.gamma_min, self.gamma_max = gamma_min, gamma_max
    c4 = n_embd * 4
    self.up = nn.ModuleList([ResnetBlock(2 * n_embd, n_embd, c4, pdrop)
                             for _ in range(n_layer + 1)])
    self.norm_out = nn.GroupNorm(32, n_embd, eps=1e-6)
    self.conv_out = nn.Conv2d(n_embd, 3, 3, padding=1)
    nn.init.zeros_(self.conv_out.weight)
    nn.init.zeros_(self.conv_out.bias)

  def forward(self, z, g_t, conditioning, deterministic=True):
    t01 = (g_t - self.gamma_min) / (self.gamma_max - self.gamma_min)      # :323-326
    zc = z.permute(0, 3, 1, 2)                      # NHWC storage, NCHW view (channels_last)
    h, hs, cond = self.trunk(zc, t01, conditioning)
    for blk in self.up:
      h = blk(torch.cat([h, hs.pop()], dim=1), cond)
    eps_pred = self.conv_out(F.silu(self.norm_out(h))) + zc
    return eps_pred.permute(0, 2, 3, 1)


class UnetEncoder(_UNetTrunk):
  """ldm/model_mulan_epsilon.py:103-154: image -> [B, latent_size] logits."""

  def __init__(self, n_embd=128, forward_n_layer=4, pdrop=0.1, latent_size=50,
               with_fourier=True):
    super().__init__(n_embd, forward_n_layer, 1, pdrop, with_fourier)
    self.norm_out = nn.GroupNorm(32, n_embd, eps=1e-6)
    self.conv_out = nn.Conv2d(n_embd, 1, 3, padding=1)
    nn.init.zeros_(self.conv_out.weight)
    nn.init.zeros_(self.conv_out.bias)
    self.dense_layer_final = nn.Linear(32 * 32, latent_size)

  def forward(self, z, deterministic=True):
    B = z.shape[0]
    zc = z.permute(0, 3, 1, 2)
    t = torch.zeros((B,), dtype=z.dtype, device=z.device)
    h, _, _ = self.trunk(zc, t, torch.zeros((B, 1), dtype=z.dtype, device=z.device))
    h = self.conv_out(F.silu(self.norm_out(h)))
    return self.dense_layer_final(F.silu(h.reshape(B, -1)))
