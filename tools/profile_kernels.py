#!/usr/bin/env python
"""One launch of each kernel of the path that is not the bench headline, for `ncu --set full` (keep ncu
summaries of the final builds in profiles/): the write-bound compiled sphere kernel and the issue-bound compiled C1 kernel
at 1025^3, from_sdf / box filter / edge filter / vector modifiers at 513^3, the brute-force nearest-neighbour kernel at
65^3 x 1 M and the octree packet walk at C4's size.

    python tools/profile_kernels.py            # plain run (must exit 0 before the same command goes under ncu)
    ncu --set full -k regex:'ab_' ... python tools/profile_kernels.py
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import aegolius_b200 as ab
    from aegolius_b200 import cabi, engine, workloads
    dev = torch.device("cuda", 0)
    lib = cabi.lib()
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    big = ab.GridSpec((4, 4, 4), (1024,) * 3)
    sph = ab.Sphere(1.0)
    sph.move((0.3, 0.1, -0.2))
    buf = torch.empty(big.n_points, dtype=torch.float32, device=dev)
    for obj in (sph, workloads.build_c1()):
        prog = ab.flatten(obj)
        engine.compile_program(prog, dtype="f32")
        engine.create_torch(prog, big, dtype="f32", out=buf)
    if "--lite" in sys.argv:  # the four shallow-tree launches only: value, then field + gradient (pull-back builds)
        gbuf = torch.empty((3, (big.n_points + 7) // 8 * 8), dtype=torch.float32, device=dev)
        for obj in (sph, workloads.build_c1()):
            prog = ab.flatten(obj)
            engine.compile_program(prog, dtype="f32", grad="spatial")
            engine.create_torch(prog, big, dtype="f32", grad="spatial", out=buf, out_grad=gbuf)
        torch.cuda.synchronize()
        print("profile_kernels ok (lite)")
        return
    del buf
    torch.cuda.empty_cache()
    res = (513, 513, 513)
    n = res[0] * res[1] * res[2]
    stride = (n + 7) // 8 * 8
    f = torch.randn(n, dtype=torch.float32, device=dev)
    o = torch.empty(n, dtype=torch.float32, device=dev)
    v = torch.randn(3, stride, dtype=torch.float32, device=dev)
    ang = torch.randn(n, dtype=torch.float32, device=dev)
    r3 = (C.c_uint32 * 3)(*res)
    k3 = (C.c_uint32 * 3)(5, 5, 1)
    g = cabi.make_grid((0.0, 0.0, 0.0), res)
    ops = (cabi.ab_vec_op * 2)()
    ops[0].opcode, ops[0].kind0, ops[0].a0 = cabi.AB_VOP_ROT_Z, cabi.AB_VK_ARRAY, ang.data_ptr()
    ops[1].opcode, ops[1].kind0, ops[1].kind1, ops[1].a1 = cabi.AB_VOP_ROT_AXIS, cabi.AB_VK_VEC3, cabi.AB_VK_ARRAY, ang.data_ptr()
    ops[1].c = (C.c_double * 3)(1.0, 0.0, 0.0)
    cabi.check(lib.ab_fd_gradient(f.data_ptr(), 0, C.byref(g), 3, cabi.AB_F32, 1, v.data_ptr(), stride, 0, st))
    cabi.check(lib.ab_box_filter(f.data_ptr(), r3, k3, 1, cabi.AB_F32, o.data_ptr(), 0, st))
    cabi.check(lib.ab_edge_filter(f.data_ptr(), r3, cabi.AB_F32, o.data_ptr(), 0, st))
    cabi.check(lib.ab_vec_apply(v.data_ptr(), stride, n, ops, 2, cabi.AB_F32, 0, st))
    cabi.check(lib.ab_vec_component(v.data_ptr(), stride, n, cabi.AB_VC_PHI, cabi.AB_F32, o.data_ptr(), 0, st))
    cabi.check(lib.ab_signed_field(f.abs().data_ptr(), r3, 0.01, cabi.AB_F32, o.data_ptr(), 0, st))
    torch.cuda.synchronize()
    del f, o, v, ang
    torch.cuda.empty_cache()
    cloud = workloads.c4_cloud()
    rec = engine.cloud_records(cloud, 3, "f32")
    c4 = ab.GridSpec(workloads.CONFIGS["C4"]["size"], workloads.CONFIGS["C4"]["res"])
    engine.point_cloud_sdf_torch(c4, rec)
    os.environ["AB_NN_ALGO"] = "brute"
    engine.point_cloud_sdf_torch(ab.GridSpec(c4.size, (64, 64, 64)), rec)
    del os.environ["AB_NN_ALGO"]
    torch.cuda.synchronize()
    print("profile_kernels ok")


if __name__ == "__main__":
    main()
