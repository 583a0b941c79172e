import sys, os, subprocess
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "goal-conditioned-rl-framework_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import ddpg as OD
from tests.test_ddpg_gpu import make_config
from gcrl_b200 import DDPG
from gcrl_b200.agent import NET_ACTOR, NET_CRITIC
B,H,L,D,A = [int(x) for x in sys.argv[1:6]]
rng = np.random.default_rng(B * 7 + H)
cfg = make_config(hidden_dim=H, layer_count=L, batch_size=B, grad_clip=0.5, tau=0.05)
ag = DDPG(D, A, cfg, None, 1, 40)
actor0, critic0 = OD.init_mlp(rng, D, H, A, L), OD.init_mlp(rng, D + A, H, 1, L)
ag._set_layers(NET_ACTOR, actor0); ag._set_layers(NET_CRITIC, critic0); ag.update_target_network()
orc = OD.DDPGOracle(actor0, critic0, gamma=cfg.gamma, tau=cfg.tau, grad_clip=cfg.grad_clip, actor_lr=cfg.actor_lr, critic_lr=cfg.critic_lr)
print("fused env", os.environ.get("GCRL_B200_NO_FUSED"))
for si, step in enumerate((39, 40, 41, 42)):
    s = rng.standard_normal((B, D)).astype(np.float32)
    ns = (s + 0.1 * rng.standard_normal((B, D))).astype(np.float32)
    a = rng.uniform(-1, 1, (B, A)).astype(np.float32)
    r = -(rng.random((B, 1)) > 0.3).astype(np.float32)
    d = (rng.random((B, 1)) < 0.1).astype(np.float32)
    want = np.array(orc.update_on_batch(step, s, a, r, ns, d))
    got = np.array([float(x) for x in ag.update(step, batch=tuple(torch.from_numpy(x).cuda() for x in (s, a, r, ns, d)))])
    print(step, "rel", np.abs(got-want)/np.abs(want))
    ga = ag.grad_tensor(NET_ACTOR).cpu().numpy()
    # compare actor gradient (post all-layers) with oracle's clipped grads: norm only
    wc = [w for w,_ in ag.critic.layers()]
    print("   max |critic w - oracle|", max(np.abs(w - ow).max() for w,(ow,_) in zip(wc, orc.critic)), " n>1e-6:", sum(int((np.abs(w - ow)>1e-6).sum()) for w,(ow,_) in zip(wc, orc.critic)))
