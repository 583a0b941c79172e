"""Oracle: Actor / Critic MLPs and the DDPG off-policy update, NumPy float32.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Restates with hand-written forward / backward (no autograd):
  * ``Actor``  src/model.py:7-45,  ``Critic`` src/model.py:48-83
  * ``DDPG.critic_update`` src/agent.py:1302-1343
  * ``DDPG.actor_update``  src/agent.py:1288-1300
  * ``DDPG.update_target_network`` src/agent.py:1255-1271
  * ``DDPG.update`` src/agent.py:1378-1404 (HER / uniform branch)
  * ``DDPG.get_gradient_norm`` src/agent.py:1279-1286
and the torch library semantics those lines rely on (torch 2.11, CPU,
single-tensor code paths): ``nn.Linear``, ``LeakyReLU(0.01)``, ``Tanh``,
``mse_loss``, ``clip_grad_norm_``, ``Adam`` / ``AdamW``, ``CosineAnnealingLR``.

Pinned by ``tests/golden/ddpg_*.npz`` produced by the unmodified reference
``DDPG`` class (see tests/golden/make_golden.py); tolerance rel 1e-5.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32
LEAKY_SLOPE = F32(0.01)


# ----------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------
def init_mlp(rng, in_dim, hidden, out_dim, layers):
    """Xavier-uniform weights, bias 0.01 (src/model.py:39-42). rng: np Generator."""
    dims = [in_dim] + [hidden] * layers + [out_dim]
    params = []
    for i in range(len(dims) - 1):
        fan_in, fan_out = dims[i], dims[i + 1]
        bound = math.sqrt(6.0 / (fan_in + fan_out))
        w = rng.uniform(-bound, bound, size=(fan_out, fan_in)).astype(F32)
        b = np.full(fan_out, 0.01, dtype=F32)
        params.append([w, b])
    return params


def clone_params(p):
    return [[w.copy(), b.copy()] for w, b in p]


def zeros_like_params(p):
    return [[np.zeros_like(w), np.zeros_like(b)] for w, b in p]


def state_dict_to_params(sd, prefix):
    """Reference checkpoint naming: ``base_net.{0,2,..}`` (Actor) / ``net.{..}``
    (Critic), weight [out,in] row-major fp32 (src/model.py:24-26,64-65)."""
    idx = sorted({int(k.split(".")[1]) for k in sd if k.startswith(prefix + ".")})
    return [[np.asarray(sd[f"{prefix}.{i}.weight"], F32).copy(),
             np.asarray(sd[f"{prefix}.{i}.bias"], F32).copy()] for i in idx]


# ----------------------------------------------------------------------------
# forward / backward
# ----------------------------------------------------------------------------
def mlp_forward(params, x, final_tanh):
    """Returns (output, cache). Hidden layers Linear->LeakyReLU(0.01); last layer
    Linear (+Tanh for the actor)."""
    acts = [np.asarray(x, F32)]
    h = acts[0]
    n = len(params)
    for i, (w, b) in enumerate(params):
        z = (h @ w.T + b).astype(F32)
        if i < n - 1:
            h = np.where(z > 0, z, z * LEAKY_SLOPE).astype(F32)
        else:
            h = np.tanh(z).astype(F32) if final_tanh else z
        acts.append(h)
    return h, acts


def mlp_backward(params, acts, d_out, final_tanh, need_input_grad=False):
    """d_out = dLoss/d(output).  Returns (grads like params, d_input or None)."""
    n = len(params)
    grads = zeros_like_params(params)
    out = acts[-1]
    dz = (d_out * (F32(1.0) - out * out)).astype(F32) if final_tanh else d_out.astype(F32)
    d_in = None
    for i in range(n - 1, -1, -1):
        w, _ = params[i]
        x = acts[i]
        grads[i][0] = (dz.T @ x).astype(F32)
        grads[i][1] = dz.sum(axis=0).astype(F32)
        if i > 0 or need_input_grad:
            dx = (dz @ w).astype(F32)
            if i > 0:
                # LeakyReLU backward keyed on the (post-activation) sign: x>0 ? 1 : slope
                dz = np.where(x > 0, dx, dx * LEAKY_SLOPE).astype(F32)
            else:
                d_in = dx
    return grads, d_in


# ----------------------------------------------------------------------------
# torch optimiser / clip semantics
# ----------------------------------------------------------------------------
def grad_norm_python(grads):
    """src/agent.py:1279-1286: sqrt(sum_p (||g_p||_2 as float)**2) in Python floats."""
    tot = 0.0
    for w, b in grads:
        for g in (w, b):
            tot += float(np.sqrt(np.sum(np.square(g, dtype=F32), dtype=F32))) ** 2
    return tot ** 0.5


def clip_grad_norm_(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ (L2): coef = max_norm/(total+1e-6), clamped
    to <=1 and ALWAYS multiplied in.  Returns the pre-clip total norm."""
    norms = np.array([np.sqrt(np.sum(np.square(g, dtype=F32), dtype=F32))
                      for pair in grads for g in pair], dtype=F32)
    total = F32(np.sqrt(np.sum(np.square(norms), dtype=F32)))
    coef = F32(max_norm) / (total + F32(1e-6))
    coef = F32(min(coef, F32(1.0)))
    for pair in grads:
        pair[0] *= coef
        pair[1] *= coef
    return float(total)


class AdamState:
    """torch.optim.Adam / AdamW, single-tensor CPU path, betas (0.9, 0.999),
    eps 1e-8.  weight_decay: 0 for Adam (DDPG, src/agent.py:1201-1202); AdamW
    default 0.01 decoupled (TD3/SAC/TQC, src/agent.py:46-48)."""

    def __init__(self, params, decoupled_wd=0.0):
        self.m = zeros_like_params(params)
        self.v = zeros_like_params(params)
        self.t = 0
        self.wd = decoupled_wd
        self.b1, self.b2, self.eps = 0.9, 0.999, 1e-8

    def step(self, params, grads, lr):
        self.t += 1
        bc1 = 1.0 - self.b1 ** self.t
        bc2 = 1.0 - self.b2 ** self.t
        step_size = F32(lr / bc1)
        bc2_sqrt = F32(bc2 ** 0.5)
        for i in range(len(params)):
            for j in range(2):
                p, g = params[i][j], grads[i][j]
                m, v = self.m[i][j], self.v[i][j]
                if self.wd:
                    p *= F32(1.0 - lr * self.wd)
                m += F32(1.0 - self.b1) * (g - m)                      # lerp_
                v *= F32(self.b2)
                v += F32(1.0 - self.b2) * g * g                         # addcmul_
                denom = np.sqrt(v) / bc2_sqrt + F32(self.eps)
                p -= step_size * (m / denom)                            # addcdiv_


class CosineAnnealingLR:
    """torch.optim.lr_scheduler.CosineAnnealingLR, recursive (chainable) form as
    driven by ``scheduler.step()`` once per optimiser step
    (src/agent.py:1203-1212,1298,1334).  ``lr`` is the rate the NEXT optimiser
    step will use."""

    def __init__(self, base_lr, T_max, eta_min):
        self.base_lr, self.T_max, self.eta_min = base_lr, T_max, eta_min
        self.last_epoch = 0
        self.lr = base_lr

    def step(self):
        self.last_epoch += 1
        e, T = self.last_epoch, self.T_max
        if (e - 1 - T) % (2 * T) == 0:
            self.lr = self.lr + (self.base_lr - self.eta_min) * (1 - math.cos(math.pi / T)) / 2
        else:
            self.lr = (1 + math.cos(math.pi * e / T)) / (1 + math.cos(math.pi * (e - 1) / T)) \
                * (self.lr - self.eta_min) + self.eta_min
        return self.lr


# ----------------------------------------------------------------------------
# DDPG
# ----------------------------------------------------------------------------
class DDPGOracle:
    """State + update rule of ``DDPG`` (src/agent.py:1173-1404) on explicit batches."""

    POLYAK_EVERY = 40  # literal at src/agent.py:1397

    def __init__(self, actor, critic, *, gamma, tau, grad_clip, actor_lr, critic_lr,
                 actor_lr_min=None, critic_lr_min=None, ac_scheduler_steps=1,
                 cr_scheduler_steps=1, ac_update_freq=1):
        self.actor = clone_params(actor)
        self.critic = clone_params(critic)
        self.target_actor = clone_params(actor)      # hard sync, :1251-1253
        self.target_critic = clone_params(critic)
        self.actor_opt = AdamState(self.actor)
        self.critic_opt = AdamState(self.critic)
        self.actor_sched = CosineAnnealingLR(actor_lr, ac_scheduler_steps,
                                             actor_lr if actor_lr_min is None else actor_lr_min)
        self.critic_sched = CosineAnnealingLR(critic_lr, cr_scheduler_steps,
                                              critic_lr if critic_lr_min is None else critic_lr_min)
        self.gamma, self.tau, self.grad_clip = gamma, tau, grad_clip
        self.ac_update_freq = ac_update_freq
        self.flip_delta = 0.0            # > 0: actor_update also evaluates _actor_flip_slack
        self.last_actor_flip_slack = 0.0
        # per-ELEMENT allowance for the actor's weights that the near-zero pre-activations seen so far can
        # explain (test tolerance, not reference behaviour; see _flip_track): [[W, b], ...] like the actor
        self.flip_allowance = zeros_like_params([[np.asarray(w, np.float64), np.asarray(b, np.float64)] for w, b in actor])
        self._flip_dm = zeros_like_params(self.flip_allowance)
        self._flip_dv = zeros_like_params(self.flip_allowance)
        self._last_flip_abs = None
        # the same for the critic (its own forward pass in critic_update)
        self.last_critic_flip_slack = 0.0
        self.flip_allowance_critic = zeros_like_params([[np.asarray(w, np.float64), np.asarray(b, np.float64)] for w, b in critic])
        self._flip_dm_c = zeros_like_params(self.flip_allowance_critic)
        self._flip_dv_c = zeros_like_params(self.flip_allowance_critic)

    # src/agent.py:1302-1343
    def critic_update(self, s, a, r, ns, d, weights=None):
        """``weights`` [B, 1]: the prioritised-replay branch (:1322-1324, :1338-1340) -- loss = mean(w * (q - y)^2)
        and the TD errors come back per sample instead of as their mean."""
        g = F32(self.gamma)
        na, _ = mlp_forward(self.target_actor, ns, final_tanh=True)
        tq, _ = mlp_forward(self.target_critic, np.concatenate([ns, na], -1), final_tanh=False)
        y = (r + g * (F32(1.0) - d) * tq).astype(F32)
        y = np.clip(y, F32(-1.0 / (1.0 - self.gamma)), F32(0.0)).astype(F32)
        q, acts = mlp_forward(self.critic, np.concatenate([s, a], -1), final_tanh=False)
        B = q.shape[0]
        diff = (q - y).astype(F32)
        if weights is None:
            loss = float(np.mean(diff * diff, dtype=F32))
            td = float(np.mean(np.abs(y - q), dtype=F32))
            dq = (F32(2.0) * diff / F32(B)).astype(F32)
        else:
            w = np.asarray(weights, F32).reshape(B, 1)
            loss = float(np.mean(w * (diff * diff), dtype=F32))
            td = np.abs(y - q).astype(F32)
            dq = (F32(2.0) * diff * w / F32(B)).astype(F32)
        grads, _ = mlp_backward(self.critic, acts, dq, final_tanh=False)
        if self.flip_delta:
            self.last_critic_flip_slack, flip_abs = self._critic_flip_slack(acts, dq, grads)
        pre_norm = grad_norm_python(grads)
        if self.grad_clip is not None:
            clip_grad_norm_(grads, self.grad_clip)
        if self.flip_delta:
            coef = 1.0 if self.grad_clip is None else min(1.0, self.grad_clip / (pre_norm + 1e-6))
            self._flip_track(self.critic_opt, grads, coef, self.critic_sched.lr, flip_abs, self.last_critic_flip_slack,
                             self._flip_dm_c, self._flip_dv_c, self.flip_allowance_critic)
        gnorm = grad_norm_python(grads)
        self.critic_opt.step(self.critic, grads, self.critic_sched.lr)
        self.critic_sched.step()
        self.last_critic_grads = grads
        self.last_y = y
        return loss, td, float(np.mean(q, dtype=F32)), gnorm

    # src/agent.py:1288-1300
    def actor_update(self, s):
        a, a_acts = mlp_forward(self.actor, s, final_tanh=True)
        q, c_acts = mlp_forward(self.critic, np.concatenate([s, a], -1), final_tanh=False)
        B = q.shape[0]
        loss = float(-np.mean(q, dtype=F32))
        dq = np.full_like(q, F32(-1.0) / F32(B))
        _, d_in = mlp_backward(self.critic, c_acts, dq, final_tanh=False, need_input_grad=True)
        d_a = d_in[:, s.shape[1]:]
        grads, _ = mlp_backward(self.actor, a_acts, d_a, final_tanh=True)
        if self.flip_delta:
            self.last_actor_flip_slack = self._actor_flip_slack(s, a_acts, c_acts, dq, grads)
        pre_norm = grad_norm_python(grads)
        if self.grad_clip is not None:
            clip_grad_norm_(grads, self.grad_clip)
        if self.flip_delta:
            coef = 1.0 if self.grad_clip is None else min(1.0, self.grad_clip / (pre_norm + 1e-6))
            self._flip_track(self.actor_opt, grads, coef, self.actor_sched.lr, self._last_flip_abs,
                             self.last_actor_flip_slack, self._flip_dm, self._flip_dv, self.flip_allowance)
        self.actor_opt.step(self.actor, grads, self.actor_sched.lr)
        self.actor_sched.step()
        self.last_actor_grads = grads
        return loss, grad_norm_python(grads)

    def _actor_flip_slack(self, s, a_acts, c_acts, dq, grads):
        """Conditioning of the actor gradient norm (test tolerance, not reference behaviour).

        LeakyReLU' is discontinuous at 0: a hidden pre-activation that is zero to within fp32
        rounding (|z| < flip_delta, a few ulps of the O(1) terms summed) takes slope 1 or 0.01 depending on summation order, and the
        gradient jumps by a finite amount.  For every such unit of the actor-phase forward
        passes, recompute the (pre-clip) gradient norm with that unit's sign flipped; the sum of
        the absolute changes bounds what any correctly rounded fp32 implementation may deviate
        by on this batch.  Returns that bound relative to the norm (0.0 when no unit is
        near zero)."""
        base = grad_norm_python(grads)
        slack = 0.0
        D = s.shape[1]
        flip_abs = zeros_like_params([[np.asarray(w, np.float64), np.asarray(b, np.float64)] for w, b in grads])
        self._last_flip_abs = flip_abs

        def row_grads(a_row, c_row, r_):
            _, d_in = mlp_backward(self.critic, c_row, dq[r_:r_ + 1], final_tanh=False, need_input_grad=True)
            g, _ = mlp_backward(self.actor, a_row, d_in[:, D:], final_tanh=True)
            return g

        for which, acts in (("actor", a_acts), ("critic", c_acts)):
            for li in range(1, len(acts) - 1):
                h = acts[li]      # post-activation: z for z > 0, 0.01 z otherwise
                rows, units = np.nonzero(np.abs(np.where(h > 0, h, h / LEAKY_SLOPE)) < self.flip_delta)
                for r_, u_ in zip(rows, units):
                    # only row r_'s contribution to the batch-summed gradient changes
                    a_row = [x[r_:r_ + 1].copy() for x in a_acts]
                    c_row = [x[r_:r_ + 1].copy() for x in c_acts]
                    g_old = row_grads(a_row, c_row, r_)
                    tgt = a_row if which == "actor" else c_row
                    tgt[li][0, u_] = F32(-1e-30) if tgt[li][0, u_] > 0 else F32(1e-30)
                    g_new = row_grads(a_row, c_row, r_)
                    g2 = [[w + (nw - ow), b + (nb - ob)]
                          for (w, b), (nw, nb), (ow, ob) in zip(grads, g_new, g_old)]
                    slack += abs(grad_norm_python(g2) - base)
                    for fa, gn, go in zip(flip_abs, g_new, g_old):
                        fa[0] += np.abs(np.asarray(gn[0], np.float64) - go[0])
                        fa[1] += np.abs(np.asarray(gn[1], np.float64) - go[1])
        return slack / max(base, 1e-30)

    def _critic_flip_slack(self, acts, dq, grads):
        """_actor_flip_slack for critic_update's own forward pass q = critic([s, a]): hidden units whose
        pre-activation is within flip_delta of zero, each flipping ONE batch row's contribution to the critic's
        gradient.  Returns (bound on the change of the gradient norm, relative; per-element |change| like grads)."""
        base = grad_norm_python(grads)
        slack = 0.0
        flip_abs = zeros_like_params([[np.asarray(w, np.float64), np.asarray(b, np.float64)] for w, b in grads])
        for li in range(1, len(acts) - 1):
            h = acts[li]
            rows, units = np.nonzero(np.abs(np.where(h > 0, h, h / LEAKY_SLOPE)) < self.flip_delta)
            for r_, u_ in zip(rows, units):
                row = [x[r_:r_ + 1].copy() for x in acts]
                g_old, _ = mlp_backward(self.critic, row, dq[r_:r_ + 1], final_tanh=False)
                row[li][0, u_] = F32(-1e-30) if row[li][0, u_] > 0 else F32(1e-30)
                g_new, _ = mlp_backward(self.critic, row, dq[r_:r_ + 1], final_tanh=False)
                g2 = [[w + (nw - ow), b + (nb - ob)] for (w, b), (nw, nb), (ow, ob) in zip(grads, g_new, g_old)]
                slack += abs(grad_norm_python(g2) - base)
                for fa, gn, go in zip(flip_abs, g_new, g_old):
                    fa[0] += np.abs(np.asarray(gn[0], np.float64) - go[0])
                    fa[1] += np.abs(np.asarray(gn[1], np.float64) - go[1])
        return slack / max(base, 1e-30), flip_abs

    def _flip_track(self, opt, grads, coef, lr, flip_abs, slack, flip_dm, flip_dv, allowance):
        """Per-element bound on how far the actor's weights may legitimately differ from this oracle's because of
        the near-zero hidden pre-activations seen so far (test tolerance, not reference behaviour).

        A pre-activation that is zero to fp32 rounding takes LeakyReLU slope 1 or 0.01 depending on the summation
        order, which changes ONE batch row's contribution to the gradient: element e of the (clipped) gradient
        moves by at most d_e = coef (sum over such units of |g_row_new - g_row_old|_e) + |g_e| slack, where slack
        bounds the change of the clip coefficient.  Adam carries the perturbation in both moments,
            dm <- b1 dm + (1 - b1) d,      dv <- b2 dv + (1 - b2) (2 |g| d + d^2),
        and the step lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps) is monotone in m and in v separately, so its extreme
        values over the box [m +- dm] x [v +- dv] are at the corners.  The allowance accumulates the largest corner
        deviation of every step, capped at Adam's hard bound 2 lr.  Elements that no flipped row touches get 0."""
        t = opt.t + 1
        bc1, bc2s = 1.0 - opt.b1 ** t, (1.0 - opt.b2 ** t) ** 0.5
        for i in range(len(grads)):
            for j in range(2):
                g = np.abs(np.asarray(grads[i][j], np.float64))
                d = coef * flip_abs[i][j] + g * slack
                dm, dv = flip_dm[i][j], flip_dv[i][j]
                dm *= opt.b1
                dm += (1.0 - opt.b1) * d
                dv *= opt.b2
                dv += (1.0 - opt.b2) * (2.0 * g * d + d * d)
                # the moments this step will use
                m = np.asarray(opt.m[i][j], np.float64) * opt.b1 + (1.0 - opt.b1) * np.asarray(grads[i][j], np.float64)
                v = np.asarray(opt.v[i][j], np.float64) * opt.b2 + (1.0 - opt.b2) * g * g
                nom = m / (np.sqrt(v) / bc2s + opt.eps)
                dev = np.zeros_like(nom)
                for mm in (m - dm, m + dm):
                    for vv in (np.maximum(v - dv, 0.0), v + dv):
                        dev = np.maximum(dev, np.abs(mm / (np.sqrt(vv) / bc2s + opt.eps) - nom))
                allowance[i][j] += np.minimum(2.0 * lr, (lr / bc1) * dev)

    # src/agent.py:1255-1271
    def soft_update(self, tau):
        t = F32(tau)
        omt = F32(1 - tau)
        for tgt, src in ((self.target_actor, self.actor), (self.target_critic, self.critic)):
            for (tw, tb), (w, b) in zip(tgt, src):
                tw[...] = t * w + omt * tw
                tb[...] = t * b + omt * tb

    # src/agent.py:1378-1404
    def update_on_batch(self, step, s, a, r, ns, d, weights=None):
        closs, td, qv, cg = self.critic_update(s, a, r, ns, d, weights)
        if step % self.POLYAK_EVERY == 0:
            self.soft_update(self.tau)
        if step % self.ac_update_freq == 0:
            aloss, agn = self.actor_update(s)
            return closs, aloss, td, qv, cg, agn
        return closs, td, qv, cg
