import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, numpy as np, torch
import aegolius_b200 as ab
from aegolius_b200 import workloads as w, engine
from aegolius_b200.grid import GridSpec
cloud = w.c4_cloud()
spec = GridSpec((4,4,4),(256,)*3)
for it in range(2):
    torch.cuda.synchronize(); t=time.time()
    out = engine.point_cloud_sdf(cloud, spec, dtype="f32", device_out=True) if 'device_out' in engine.point_cloud_sdf.__code__.co_varnames else engine.point_cloud_sdf(cloud, spec, dtype="f32")
    torch.cuda.synchronize(); dt=time.time()-t
    print("C4 full ms", dt*1e3, "Tpairs/s", 257**3*cloud.shape[-1]/dt/1e12 if cloud.shape[0]==3 else 257**3*cloud.shape[0]/dt/1e12)
