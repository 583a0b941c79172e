import sys, os, time, cProfile, pstats, random
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "goal-conditioned-rl-framework_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch, bench
from gcrl_b200 import DDPG
sys.argv = sys.argv[:1]
args = bench.parse()
T, k, O, G, A, B = 50, 4, 18, 3, 3, 256
E = 4000
data = bench.synth(np.random.default_rng(0), E, T, O, G, A, k)
ag = DDPG(O + G, A, bench.agent_config(args, E * 246), None, 1, 40)
for e in range(E):
    ag.buffer.push_episode(data["s"][e], data["a"][e], data["ns"][e], data["r"][e], data["d"][e], data["ag"][e], data["fut"][e])
random.seed(1)
def step(i):
    if i % 40 == 0:
        for e in range(16):
            j = (i // 40 * 16 + e) % E
            ag.buffer.push_episode(data["s"][j], data["a"][j], data["ns"][j], data["r"][j], data["d"][j], data["ag"][j], None)
    return ag.update(1 + i)
for i in range(50): step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(400): step(50 + i)
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t0) / 400 * 1e3)
pr = cProfile.Profile(); pr.enable()
for i in range(400): step(450 + i)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
