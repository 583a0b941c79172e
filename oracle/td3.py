"""Oracle: the TD3 off-policy update, NumPy float32.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Restates ``TD3Agent`` (src/agent.py:12-386) on explicit batches and explicit smoothing noise:
  * ``critic_update`` :164-251  -- target-policy smoothing ``clamp(randn * policy_noise, +-noise_clamp)``,
    ``a' = clamp(pi_t(s') + noise, -1, 1)``, ``y = r + gamma (1 - d) min(Q1_t, Q2_t)`` (no clamp of y),
    smooth-L1 (beta 1) losses, critic 1 UNCLIPPED (:201 is commented out), critic 2 clipped, AdamW
  * ``actor_update`` :149-162   -- ``-mean Q1(s, pi(s))`` through the stepped critic 1
  * ``update_critic`` / ``update_actor`` :117-132 and ``update`` :281-317 -- critic targets Polyak
    every step, actor target only after an actor step.
Pinned by ``tests/golden/td3_*.npz`` (unmodified reference, recorded ``torch.randn_like`` draws).
"""
from __future__ import annotations

import numpy as np

from .ddpg import (AdamState, CosineAnnealingLR, F32, clip_grad_norm_, clone_params, grad_norm_python,
                   mlp_backward, mlp_forward)

ADAMW_WD = 0.01  # torch.optim.AdamW default, src/agent.py:46-48


def smooth_l1(q, y, weights=None):
    """torch.nn.functional.smooth_l1_loss(q, y), beta = 1, mean reduction -> (loss, dloss/dq); with ``weights``
    the prioritised form ``(weights * smooth_l1_loss(q, y, reduction="none")).mean()`` (:193-197)."""
    diff = (q - y).astype(F32)
    ad = np.abs(diff)
    elem = np.where(ad < 1, F32(0.5) * diff * diff, ad - F32(0.5)).astype(F32)
    grad = np.where(ad < 1, diff, np.sign(diff)).astype(F32) / F32(q.shape[0])
    if weights is not None:
        w = np.asarray(weights, F32).reshape(q.shape)
        elem, grad = (w * elem).astype(F32), (w * grad).astype(F32)
    return float(np.mean(elem, dtype=F32)), grad.astype(F32)


class TD3Oracle:
    def __init__(self, actor, critic_1, critic_2, *, gamma, tau, grad_clip, actor_lr, critic_lr,
                 policy_noise, noise_clamp, actor_lr_min=None, critic_lr_min=None, ac_scheduler_steps=1,
                 cr_scheduler_steps=1, ac_update_freq=1):
        self.actor, self.critic_1, self.critic_2 = clone_params(actor), clone_params(critic_1), clone_params(critic_2)
        self.target_actor = clone_params(actor)
        self.target_critic_1, self.target_critic_2 = clone_params(critic_1), clone_params(critic_2)
        self.actor_opt = AdamState(self.actor, ADAMW_WD)
        self.c1_opt, self.c2_opt = AdamState(self.critic_1, ADAMW_WD), AdamState(self.critic_2, ADAMW_WD)
        self.actor_sched = CosineAnnealingLR(actor_lr, ac_scheduler_steps,
                                             actor_lr if actor_lr_min is None else actor_lr_min)
        self.critic_sched = CosineAnnealingLR(critic_lr, cr_scheduler_steps,
                                              critic_lr if critic_lr_min is None else critic_lr_min)
        self.gamma, self.tau, self.grad_clip = gamma, tau, grad_clip
        self.policy_noise, self.noise_clamp, self.ac_update_freq = policy_noise, noise_clamp, ac_update_freq

    def critic_update(self, s, a, r, ns, d, randn, weights=None):         # :164-251
        noise = np.clip((randn * F32(self.policy_noise)).astype(F32), F32(-self.noise_clamp), F32(self.noise_clamp))
        na, _ = mlp_forward(self.target_actor, ns, final_tanh=True)
        na = np.clip((na + noise).astype(F32), F32(-1), F32(1))
        tin = np.concatenate([ns, na], -1)
        tq = np.minimum(mlp_forward(self.target_critic_1, tin, False)[0], mlp_forward(self.target_critic_2, tin, False)[0])
        y = (r + F32(self.gamma) * (F32(1) - d) * tq).astype(F32)
        cin = np.concatenate([s, a], -1)
        q1, acts1 = mlp_forward(self.critic_1, cin, False)
        q2, acts2 = mlp_forward(self.critic_2, cin, False)
        loss1, dq1 = smooth_l1(q1, y, weights)
        g1, _ = mlp_backward(self.critic_1, acts1, dq1, False)
        gn1 = grad_norm_python(g1)                                        # unclipped (:201)
        self.c1_opt.step(self.critic_1, g1, self.critic_sched.lr)
        loss2, dq2 = smooth_l1(q2, y, weights)
        g2, _ = mlp_backward(self.critic_2, acts2, dq2, False)
        clip_grad_norm_(g2, self.grad_clip)
        gn2 = grad_norm_python(g2)
        self.c2_opt.step(self.critic_2, g2, self.critic_sched.lr)
        self.critic_sched.step()
        td = np.maximum(np.abs(q1 - y), np.abs(q2 - y)).astype(F32)       # per sample when prioritised (:232-233)
        if weights is None:
            td = float(np.mean(td, dtype=F32))
        qv = float(np.mean(np.concatenate([q1, q2], -1), dtype=F32))
        return loss1, loss2, td, qv, gn1, gn2

    def actor_update(self, s):                                            # :149-162
        a, a_acts = mlp_forward(self.actor, s, True)
        q, c_acts = mlp_forward(self.critic_1, np.concatenate([s, a], -1), False)
        dq = np.full_like(q, F32(-1.0) / F32(q.shape[0]))
        _, d_in = mlp_backward(self.critic_1, c_acts, dq, False, need_input_grad=True)
        grads, _ = mlp_backward(self.actor, a_acts, d_in[:, s.shape[1]:], True)
        clip_grad_norm_(grads, self.grad_clip)
        gn = grad_norm_python(grads)
        self.actor_opt.step(self.actor, grads, self.actor_sched.lr)
        self.actor_sched.step()
        return float(-np.mean(q, dtype=F32)), gn

    @staticmethod
    def _polyak(tgt, src, tau):
        t, omt = F32(tau), F32(1 - tau)
        for (tw, tb), (w, b) in zip(tgt, src):
            tw[...] = t * w + omt * tw
            tb[...] = t * b + omt * tb

    def update_on_batch(self, step, s, a, r, ns, d, randn, weights=None):  # :281-317
        l1, l2, td, qv, g1, g2 = self.critic_update(s, a, r, ns, d, randn, weights)
        self._polyak(self.target_critic_1, self.critic_1, self.tau)
        self._polyak(self.target_critic_2, self.critic_2, self.tau)
        if step % self.ac_update_freq == 0:
            al, ag = self.actor_update(s)
            self._polyak(self.target_actor, self.actor, self.tau)
            return l1, l2, al, td, qv, g1, g2, ag
        return l1, l2, td, qv, g1, g2
