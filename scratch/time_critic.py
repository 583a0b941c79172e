import sys, os, ctypes as C
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "goal-conditioned-rl-framework_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from tests.test_ddpg_gpu import make_config
from gcrl_b200 import DDPG
from gcrl_b200._lib import lib, check, vp
B,H,L,D,A = [int(x) for x in sys.argv[1:6]]
cfg = make_config(hidden_dim=H, layer_count=L, batch_size=B)
ag = DDPG(D, A, cfg, None, 1, 40)
rng = np.random.default_rng(0)
batch = tuple(torch.from_numpy(x).cuda() for x in (rng.standard_normal((B,D)).astype(np.float32), rng.uniform(-1,1,(B,A)).astype(np.float32), -np.ones((B,1),np.float32), rng.standard_normal((B,D)).astype(np.float32), np.zeros((B,1),np.float32)))
for s in (1,2,3): ag.update(s, batch=batch)
ms = C.c_float()
check(lib.gcrl_agent_time_critic_kernel(ag._h, B, 200, C.byref(ms), vp(torch.cuda.current_stream().cuda_stream)))
print(f"R={os.environ.get('GCRL_FUSED_R','default')} cluster={os.environ.get('GCRL_B200_CLUSTER','0')} B={B} H={H}: critic kernel {ms.value*1000:.1f} us")
