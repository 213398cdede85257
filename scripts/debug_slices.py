import os, sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
import fruits_b200 as fruits
from fruits_b200 import _jit, _backend as be
import specs
name=sys.argv[1] if len(sys.argv)>1 else "C3_general"
n={"C3_general":10000,"C2_reduced":1000,"C4_twi":100000}[name]
X=torch.from_numpy(specs.make_input(name,n)).cuda()
fruit=specs.build_fruit(fruits,specs.SPECS[name])
np.random.seed(0); fruit.fit(X[:64] if name!="C4_twi" else X[:64])
def timed(f,reps=3):
    f(); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter()-t0)/reps*1e3
for si,slc in enumerate(fruit._slices):
    out=torch.zeros((n,slc.nfeatures()),dtype=torch.float64,device='cuda')
    for mode in ("force","0"):
        os.environ["FRUITS_B200_JIT"]=mode
        ms=timed(lambda: slc._transform_device(X,[],None,out,0,True))
        print(name,"slice",si,"jit" if mode=="force" else "generic","%.2f ms"%ms, slc._last_launch[:2])
    os.environ["FRUITS_B200_JIT"]="force"
    from fruits_b200.cache import SharedSeedCache
    cache=SharedSeedCache(X)
    ms=timed(lambda: slc._prepare_device(X,cache,fit=False))
    print("   prepare %.2f ms"%ms)
    iss=slc._iss[0]; iss._cache=cache
    ms=timed(lambda: iss._lookup(X)); print("   lookup %.2f ms"%ms)
    kern=slc._last_launch[2] if slc._last_launch[2] is not None else None
    if kern is not None:
        print("   modules",len(kern.modules),"parts",len(kern.em.p.parts),"gpc",kern.em.gpc,"tt",kern.em.tt,"ppc",kern.em.ppc,"smem",kern.em.smem_bytes())
