import os, sys, time, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import aegolius_b200 as ab
from aegolius_b200 import engine
obj = ab.workloads.build_c3(); spec = ab.GridSpec((6,6,6),(1024,)*3)
n = spec.n_points
pf = engine.PinnedArray((n,), np.float32); pg = engine.PinnedArray((3, n), np.float32)
ab.create(obj, spec, dtype="f32", grad="spatial", out=pf.array, out_grad=pg.array)
engine.wait_for_compilations()
ts=[]
for _ in range(4):
    t0=time.perf_counter(); ab.create(obj, spec, dtype="f32", grad="spatial", out=pf.array, out_grad=pg.array); ts.append(time.perf_counter()-t0)
print(json.dumps({"AB_JIT_COMPACT": os.environ.get("AB_JIT_COMPACT","1"), "e2e_ms": [round(t*1e3,1) for t in ts]}))
