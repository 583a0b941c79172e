"""Wall-clock ms per SAC / TQC update on an explicit device batch (bench.py shapes)."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "goal-conditioned-rl-framework_b200")
import torch, bench
from gcrl_b200 import SACAgent, TQCAgent
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sys.argv = sys.argv[:1]
args = bench.parse()
args.batch = B
for name, cls in (("sac", SACAgent), ("tqc", TQCAgent)):
    cfg = bench.agent_config(args, 100000)
    cfg.alpha_lr, cfg.alpha_min, cfg.alpha_min_steps, cfg.grad_clip = 3e-4, 0.05, 0, 1.0
    ag = cls(args.obs + args.goal, args.act, cfg, None, 1, 40, index_source="device")
    D, A = args.obs + args.goal, args.act
    batch = (torch.randn(B, D, device="cuda"), torch.rand(B, A, device="cuda"), -torch.ones(B, 1, device="cuda"),
             torch.randn(B, D, device="cuda"), torch.zeros(B, 1, device="cuda"))
    for i in range(5):
        ag.update(i + 1, batch=batch)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(100):
        ag.update(6 + i, batch=batch)
    torch.cuda.synchronize()
    print(name, "B", B, round((time.perf_counter() - t0) * 10, 4), "ms/update")
