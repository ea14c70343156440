import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aegolius_b200 as ab
from aegolius_b200 import workloads as w, engine
cloud = w.c4_cloud()
pc = ab.PointCloud3D(cloud); pc.onion(0.02); pc.rotate(0.2, (0, 0, 1))
prog = ab.flatten(pc)
spec = ab.GridSpec((2.5, 2.5, 1.5), (256,) * 3)
for dt, gr in (("f32", None), ("f32", "spatial"), ("f64", None)):
    for it in range(2):
        torch.cuda.synchronize(); t = time.time()
        out = engine.create_torch(prog, spec, dtype=dt, grad=gr, device=0)
        torch.cuda.synchronize(); dtm = time.time() - t
    print("leaf-in-tree 1M cloud, 257^3", dt, gr, "ms", round(dtm * 1e3, 2))
