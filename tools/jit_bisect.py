#!/usr/bin/env python
"""Finds the first op at which a program-compiled kernel and the interpreter disagree: evaluates every prefix of a
program that ends in a value with both and reports the first one whose outputs differ.

    python tools/jit_bisect.py --case c2 --prebuild     # here: nvcc the prefix kernels
    python tools/jit_bisect.py --case c2                # on the GPU box
"""
import argparse
import json
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="c2")
    ap.add_argument("--prebuild", action="store_true")
    ap.add_argument("--slots", default="reg")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--res", type=int, default=512)
    args = ap.parse_args()
    import numpy as np
    import aegolius_b200 as ab
    from aegolius_b200 import codegen as cg, engine, cabi, opcodes as oc
    from aegolius_b200.program import Program, OP_DTYPE
    from jit_sweep import cases
    obj, spec, dtype, grad = cases()[args.case]
    spec = ab.GridSpec(spec.size, (args.res,) * (2 if engine._is_2d(spec) else 3))
    prog = ab.flatten(obj)
    ks = [k for k in range(1, prog.n_ops) if int(prog.ops[k - 1]["opcode"]) >= 32]  # prefixes ending in a value-producing op
    prefixes = []
    for k in ks:
        ops = np.concatenate([prog.ops[:k], np.array([(oc.END, 0, 0, 0)], dtype=OP_DTYPE)])
        prefixes.append((k, Program(ops, prog.args, prog.blobs, prog.n_pslots, prog.n_vslots)))
    g = cg._GRAD_NAMES[grad]
    if args.prebuild:
        cg._build_slots = __import__("threading").Semaphore(os.cpu_count())

        def build(item):
            k, p = item
            return k, cg.build_source(cg.generate(cg.signature(p), dtype, g, is2d=engine._is_2d(spec), slots=args.slots, width=args.width))
        with ThreadPoolExecutor(os.cpu_count()) as ex:
            for k, path in ex.map(build, prefixes):
                print(k, os.path.basename(path), flush=True)
        return
    import torch
    lib = cabi.lib()
    os.environ["AB_JIT"] = "off"
    for k, p in prefixes:
        lib.ab_prog_enable(0)
        ref = engine.create_torch(p, spec, dtype=dtype, grad=grad)
        ref = ref[0] if isinstance(ref, tuple) else ref
        lib.ab_prog_enable(1)
        sig = cg.signature(p)
        path = cg.binary_path(cg.generate(sig, dtype, g, is2d=engine._is_2d(spec), slots=args.slots, width=args.width))
        cg._register(path, sig, dtype, g, engine._is_2d(spec))
        out = engine.create_torch(p, spec, dtype=dtype, grad=grad)
        out = out[0] if isinstance(out, tuple) else out
        nbad = int((out != ref).sum())
        last = p.ops[k - 1]
        print(json.dumps({"prefix": k, "last_op": oc.NAMES[int(last["opcode"])], "a": int(last["a"]), "b": int(last["b"]),
                          "mismatches": nbad, "max_abs_diff": float((out - ref).abs().max())}), flush=True)


if __name__ == "__main__":
    main()
