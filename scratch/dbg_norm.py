import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "goal-conditioned-rl-framework_b200"))
import numpy as np
np.show_config() if False else None
rng=np.random.default_rng(0)
def emu(x):
    n,dim=x.shape
    s=x[0].copy()
    for r in range(1,n): s=(s+x[r]).astype(np.float32)
    mean=(s/np.float32(n)).astype(np.float32)
    d=(x-mean).astype(np.float32); sq=(d*d).astype(np.float32)
    v=sq[0].copy()
    for r in range(1,n): v=(v+sq[r]).astype(np.float32)
    var=(v/np.float32(n)).astype(np.float32)
    return mean,var
for n in (2,5,17,40,4000):
    for dim in (2,3,4,7,19,20):
        x=(rng.standard_normal((n,dim))*3+1).astype(np.float32)
        m,v=emu(x)
        print(n,dim,"mean eq",np.array_equal(m,np.mean(x,axis=0)),"var eq",np.array_equal(v,np.var(x,axis=0)))
from gcrl_b200 import RunningNormalizer
for n,dim in ((5,3),(17,3),(40,3),(40,4),(40,2),(40,20)):
    x=(rng.standard_normal((n,dim))*3+1).astype(np.float32)
    nz=RunningNormalizer(dim); nz.update(x)
    m,v=emu(x)
    # expected state after one update from (0,1,1e-8)
    print(n,dim,"gpu var",nz.var, "np var", np.var(x,axis=0), "emu", v)
