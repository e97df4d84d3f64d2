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


def clip_by_global_norm(grads, max_norm):
  """optax.clip_by_global_norm (ldm/experiment.py:176-178) over a list of gradient tensors:
  g_norm = sqrt(sum_leaves sum(g^2)); g if g_norm < max_norm else (g / g_norm) * max_norm."""
  g_norm = torch.sqrt(sum(torch.sum(g * g) for g in grads))
  if g_norm < max_norm:
    return list(grads), g_norm
  return [(g / g_norm) * max_norm for g in grads], g_norm


def linear_schedule(init_value, end_value, transition_steps, count):
  """optax.linear_schedule == polynomial_schedule(power=1)."""
  count = min(max(count, 0), transition_steps)
  frac = 1 - count / transition_steps
  return (init_value - end_value) * frac + end_value


def lr_schedule(count, learning_rate, warmup, lr_decay=False, num_steps_train=0):
  """get_lr_schedule (ldm/experiment.py:106-129): optax.join_schedules([warmup, decay],
  boundaries=[warmup]) - the decay schedule sees count - warmup."""
  if lr_decay and count >= warmup:
    return linear_schedule(learning_rate, 0.0, num_steps_train - warmup, count - warmup)
  return linear_schedule(0.0, learning_rate, warmup, count)


def adamw_ema_step(p, g, mu, nu, ema, count, lr, b1=0.9, b2=0.99, eps=1e-8, weight_decay=0.01,
                   decay_mask=None, ema_rate=0.9999, grad_scale=1.0, clip_norm=None):
  """One update. `count` is the 1-based step. Returns (p, mu, nu, ema).  clip_norm clips by the
  norm of THIS tensor (pass the whole flat bucket to get the global norm)."""
  g = g * grad_scale
  if clip_norm:
    (g,), _ = clip_by_global_norm([g], clip_norm)
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
