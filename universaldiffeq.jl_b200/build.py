"""Builds csrc/libude_cuda.so for sm_100a with nvcc (cross-compiles without a GPU).

One object per translation unit, compiled in parallel; `-lineinfo` so ncu's source page maps to the code.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
SO = os.path.join(CSRC, "libude_cuda.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden"]


def _deps_mtime():
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(CSRC, "*.inc")) + [os.path.join(HERE, "..", "include", "ude_cuda.h")]
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, extra):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    if os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), _deps_mtime()):
        return obj, ""
    cmd = [NVCC, *FLAGS, *extra, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(verbose: bool = False, force: bool = False, ptxas_v: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    if force:
        for f in glob.glob(os.path.join(OBJ, "*.o")):
            os.remove(f)
    extra = ["-Xptxas", "-v"] if ptxas_v else []
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        res = list(ex.map(lambda s: _compile(s, extra), srcs))
    objs = [o for o, _ in res]
    if verbose or ptxas_v:
        for _, log in res:
            if log:
                sys.stderr.write(log)
    if (not os.path.exists(SO)) or any(os.path.getmtime(o) > os.path.getmtime(SO) for o in objs):
        cmd = [NVCC, "-shared", "-o", SO, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl",
               "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return SO


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv, ptxas_v="--ptxas" in sys.argv))
