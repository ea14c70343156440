import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import aegolius_b200 as ab
from aegolius_b200 import workloads as w
from aegolius_b200.grid import GridSpec
rng = np.random.default_rng(1)
def both(co, pts, **kw):
    os.environ["AB_NN_ALGO"] = "brute"; a = ab.point_cloud_sdf(co, pts, **kw)
    os.environ["AB_NN_ALGO"] = "tree"; b = ab.point_cloud_sdf(co, pts, **kw)
    return a, b
cases = {
  "uniform": rng.uniform(-1, 1, (3, 20000)),
  "surface": w.c4_cloud(50000, 3),
  "cluster": np.concatenate([rng.normal(0, 0.01, (3, 5000)), rng.normal(1.5, 0.3, (3, 5000))], axis=1),
  "single": np.array([[0.3], [0.2], [-0.1]]),
  "dups": np.repeat(rng.uniform(-1, 1, (3, 10)), 300, axis=1),
  "line": np.stack([np.linspace(-1, 1, 3000), np.zeros(3000), np.zeros(3000)]),
  "far": rng.uniform(100, 100.5, (3, 4000)),
}
spec = GridSpec((4, 4, 4), (40, 40, 40))
ok = True
for name, pts in cases.items():
    for dt in ("f32", "f64"):
        for dim in (3, 2):
            a, b = both(spec if dim == 3 else GridSpec((4, 4), (150, 150)), pts, dim=dim, dtype=dt)
            same = np.array_equal(a, b)
            ok &= same
            print(name, dt, dim, "identical" if same else f"DIFF max {np.max(np.abs(a-b))} n {np.sum(a!=b)}")
co = rng.uniform(-3, 3, (3, 70001))
a, b = both(co, cases["uniform"], dim=3, dtype="f32"); print("points mode", np.array_equal(a, b)); ok &= np.array_equal(a, b)
print("ALL OK" if ok else "FAIL")
# timing
import torch
cloud = w.c4_cloud()
spec = GridSpec((2.5, 2.5, 1.5), (256,) * 3)
for lv in (6, 7, 8):
    os.environ["AB_NN_LEVELS"] = str(lv)
    for it in range(2):
        t = time.time(); out = ab.point_cloud_sdf(spec, cloud, dtype="f32"); dt = time.time() - t
    print("C4 tree levels", lv, "ms (incl upload+d2h)", dt * 1e3)
del os.environ["AB_NN_LEVELS"]
vol = rng.uniform(-1.2, 1.2, (3, 1000000))
for lv in (6, 7, 8):
    os.environ["AB_NN_LEVELS"] = str(lv)
    for it in range(2):
        t = time.time(); out2 = ab.point_cloud_sdf(spec, vol, dtype="f32"); dt = time.time() - t
    print("volume cloud tree levels", lv, "ms", dt * 1e3)
# 2D tuning
import torch, ctypes as C
from aegolius_b200 import cabi
lib = cabi.lib()
def time_grid(pts, spec, dim, reps=3):
    d_cloud = C.c_void_p()
    cabi.check(lib.ab_cloud_upload(pts.ctypes.data, pts.shape[1], dim, pts.shape[1], cabi.AB_F32, 0, C.byref(d_cloud)))
    out = torch.empty(spec.n_points, dtype=torch.float32, device="cuda")
    g = cabi.make_grid(spec.size, spec.res)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); cabi.check(lib.ab_nn_grid(d_cloud, pts.shape[1], dim, C.byref(g), cabi.AB_F32, out.data_ptr(), 0, st)); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    lib.ab_device_free(d_cloud, 0)
    return best
os.environ["AB_NN_ALGO"] = "tree"
t = np.linspace(0, 40 * np.pi, 1000000)
curve2d = np.stack([0.09 * t / np.pi * np.cos(t) , 0.09 * t / np.pi * np.sin(t), 0 * t])
fill2d = np.concatenate([rng.uniform(-3.5, 3.5, (2, 1000000)), np.zeros((1, 1000000))])
spec2 = GridSpec((8, 8), (4096, 4096))
for name, pts in (("curve2d", curve2d), ("fill2d", fill2d)):
    for lv in (8, 9, 10, 11):
        for lf in (8, 32):
            os.environ["AB_NN_LEVELS"] = str(lv); os.environ["AB_NN_LEAF"] = str(lf)
            print(name, "L", lv, "leaf", lf, "ms", time_grid(pts, spec2, 2))
spec3 = GridSpec((2.5, 2.5, 1.5), (256,) * 3)
for name, pts in (("c4", cloud), ("vol", vol)):
    for lv in (6, 7, 8):
        for lf in (16, 32, 64):
            os.environ["AB_NN_LEVELS"] = str(lv); os.environ["AB_NN_LEAF"] = str(lf)
            print(name, "L", lv, "leaf", lf, "ms", time_grid(pts, spec3, 3))
for m in (1000, 10000, 100000):
    sub = np.ascontiguousarray(vol[:, :m])
    for lv in (3, 4, 5, 6, 7):
        os.environ["AB_NN_LEVELS"] = str(lv); os.environ["AB_NN_LEAF"] = "32"
        print("vol m", m, "L", lv, "ms", time_grid(sub, spec3, 3))
    os.environ["AB_NN_ALGO"] = "brute"; print("vol m", m, "brute ms", time_grid(sub, spec3, 3, reps=1)); os.environ["AB_NN_ALGO"] = "tree"
