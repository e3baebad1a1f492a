"""Build libmaskunet_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python maskunet_b200/build.py [--force] [-v]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmaskunet_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-Xptxas", "-v", "-I", os.path.join(os.path.dirname(HERE), "include")]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "maskunet_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def compile_one(src, force, hm, log):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    srcp = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(srcp), hm):
        return obj, False
    cmd = [NVCC, *ARCH, *FLAGS, "-c", srcp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log[src] = r.stderr
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stderr}\n{r.stdout}")
    return obj, True


def build_variant(out: str, defines, objdir: str) -> str:
    """A/B build: the same sources with extra -D flags into another .so (select it with MASKUNET_B200_LIB)."""
    os.makedirs(objdir, exist_ok=True)
    objs = []
    def one(src):
        obj = os.path.join(objdir, src[:-3] + ".o")
        cmd = [NVCC, *ARCH, *FLAGS, *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stderr}")
        return obj
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(one, sources()))
    r = subprocess.run([NVCC, *ARCH, "-shared", "-o", out, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr}")
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hm = headers_mtime()
    log = {}
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        res = list(ex.map(lambda s: compile_one(s, force, hm, log), sources()))
    objs = [o for o, _ in res]
    if any(ch for _, ch in res) or not os.path.exists(LIB):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stderr}")
    if verbose:
        for s, t in log.items():
            print(f"==== {s}\n{t}")
    with open(os.path.join(OBJ, "ptxas.log"), "a") as fh:
        for s, t in log.items():
            fh.write(f"==== {s}\n{t}\n")
    return LIB


if __name__ == "__main__":
    if "--variant" in sys.argv:       # python build.py --variant out.so DEF1=V1 DEF2=V2 ...
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:], OBJ + "_variant"))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
