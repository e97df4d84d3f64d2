"""CPU oracle of the optimizer update (TEST INFRASTRUCTURE ONLY, like mulan_oracle.py).

Restates, op by op, what TrainState.apply_gradients (ldm/train_state.py:70-102) runs with the
optax chain of ldm/experiment.py:132-182:

  optax.adamw(lr, b1, b2, eps, weight_decay, mask) ==
      chain(scale_by_adam(b1, b2, eps, eps_root=0), add_decayed_weights(wd, mask),
            scale_by_learning_rate(lr))
  new_params = params + updates                      (optax.apply_updates)
  new_ema    = ema + (1 - ema_rate) * (new_params - ema)     (train_state.py:91-95)

optax 0.1.x formulas (the un-vendored dependency the reference pins, requirements.txt):
  update_moment(g, m, decay, order) = (1 - decay) * g**order + decay * m
  bias_correction(m, decay, count)  = m / (1 - decay**count)
  scale_by_adam: u = mu_hat / (sqrt(nu_hat + eps_root) + eps)

Parity unpinned against the real optax (not installable here); the formulas are optax's
published ones and are cross-checked against torch.optim.AdamW in tests/test_optim.py.
"""
import torch


def adamw_ema_step(p, g, mu, nu, ema, count, lr, b1=0.9, b2=0.99, eps=1e-8, weight_decay=0.01,
                   decay_mask=None, ema_rate=0.9999, grad_scale=1.0):
  """One update. `count` is the 1-based step. Returns (p, mu, nu, ema)."""
  g = g * grad_scale
  mu = (1 - b1) * g + b1 * mu
  nu = (1 - b2) * (g * g) + b2 * nu
  dt = p.dtype
  bc1 = 1 - torch.tensor(b1, dtype=dt) ** count
  bc2 = 1 - torch.tensor(b2, dtype=dt) ** count
  mu_hat = mu / bc1
  nu_hat = nu / bc2
  u = mu_hat / (torch.sqrt(nu_hat) + eps)
  if decay_mask is None:
    u = u + weight_decay * p
  else:
    u = torch.where(decay_mask, u + weight_decay * p, u)
  p = p + (-lr) * u
  ema = ema + (1. - ema_rate) * (p - ema)
  return p, mu, nu, ema
