import os, sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from sqcc_b200.backend import FockEngine
n=494
eng=FockEngine(device=0); eng.attach_eri_synth(n, 20261019); eng.synchronize()
rng=np.random.default_rng(7); q,_=np.linalg.qr(rng.standard_normal((n,n))); d=2.0*q[:,:47]@q[:,:47].T
dd=torch.from_numpy(d).cuda()
def warm(f):
    t0=time.time()
    while time.time()-t0<1.5:
        for _ in range(5): f()
    torch.cuda.synchronize()
def tm(f,steps=20):
    warm(f); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(steps): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/steps*1e3
print('back-to-back device        ms', tm(lambda: eng.jk_device(dd, True)))
def dev_sync():
    eng.jk_device(dd, True); torch.cuda.synchronize()
print('device + sync per step     ms', tm(dev_sync))
d_pin=eng.pinned(n,n); d_pin[...]=d; out=(eng.pinned(n,n), eng.pinned(1,n,n))
print('host pinned in/out         ms', tm(lambda: eng.jk(d_pin, out=out)))
print('host pageable              ms', tm(lambda: eng.jk(d)))
print('kernel ms (last)', eng.timings())
