"""Builds libaegolius_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m aegolius_b200.build [--force]

Each .cu is compiled to an object in parallel (the four interpreter instantiations dominate), then linked into one
shared library next to this file so that it travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libaegolius_b200.so")
UNITS = ["ab_capi.cu", "ab_interp_f32.cu", "ab_interp_f32g.cu", "ab_interp_f64.cu", "ab_interp_f64g.cu",
         "ab_interp_f32p.cu", "ab_interp_f64p.cu", "ab_interp_f32_mid.cu", "ab_interp_f32g_mid.cu", "ab_interp_f32_lite.cu", "ab_interp_f32g_lite.cu", "ab_interp_f64_lite.cu", "ab_interp_f64g_lite.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xptxas", "-v"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _digest(unit, extra_flags):
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS + list(extra_flags)).encode())
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cuh", ".h")) or f == unit:
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    return h.hexdigest()


def _compile(unit, extra_flags, force):
    obj = os.path.join(BUILD, unit.replace(".cu", ".o"))
    stamp = obj + ".sha"
    dig = _digest(unit, extra_flags)
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, "cached"
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + ["-c", os.path.join(CSRC, unit), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {unit}:\n{r.stdout}\n{r.stderr}")
    with open(obj.replace(".o", ".ptxas.txt"), "w") as fh:
        fh.write(r.stderr)
    with open(stamp, "w") as fh:
        fh.write(dig)
    return obj, "built"


SPEC_DIR = os.path.join(HERE, "spec")


def build_specialized(enabled_ops, kind, verbose=True):
    """Compiles csrc/ab_interp_spec.cu with only `enabled_ops` (names as in opcodes.NAMES) into
    aegolius_b200/spec/spec_<digest>.so and returns its path (cached: the digest covers the sources, the flags and the op
    set). kind: 0 fp32 values, 1 fp32 + spatial gradient, 2 fp64 values, 3 fp64 + spatial gradient."""
    from . import opcodes as oc
    names = sorted(set(oc.NAMES.values()))
    enabled = sorted(set(enabled_ops) | {"END"})
    unknown = [n for n in enabled if n not in names]
    if unknown:
        raise ValueError(f"unknown ops {unknown}")
    flags = [f"-DAB_SPEC_KIND={int(kind)}", f"-DAB_SPEC_MIN_CTAS={7 if kind < 2 else 6}"]
    flags += [f"-DAB_SPEC_{n}=0" for n in names if n not in enabled]
    dig = hashlib.sha256((_digest("ab_interp_spec.cu", flags) + ",".join(enabled)).encode()).hexdigest()[:20]
    os.makedirs(SPEC_DIR, exist_ok=True)
    out = os.path.join(SPEC_DIR, f"spec_{dig}.so")
    if os.path.exists(out):
        return out
    tmp = out + f".tmp{os.getpid()}"
    cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")] + flags + [
        "-Xcompiler", "-fvisibility=hidden", "-shared", "-o", tmp, os.path.join(CSRC, "ab_interp_spec.cu"), "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for the specialised kernel:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, out)
    if verbose:
        print(f"[aegolius_b200.build] specialised kernel ({len(enabled)} ops, kind {kind}): {out}")
    return out


def build(force=False, extra_flags=(), verbose=True):
    os.makedirs(BUILD, exist_ok=True)
    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        results = list(ex.map(lambda u: _compile(u, extra_flags, force), UNITS))
    objs = [o for o, _ in results]
    if force or not os.path.exists(LIB) or any(s == "built" for _, s in results):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        for (o, s), u in zip(results, UNITS):
            print(f"[aegolius_b200.build] {u}: {s}")
        print(f"[aegolius_b200.build] {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
