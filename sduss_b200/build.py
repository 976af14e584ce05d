"""Builds sduss_b200/libsduss_b200.so in-tree with nvcc for sm_100a (no JIT, no env vars).

The reference JIT-compiled its extension at import with no arch flags
(sduss/model_executor/modules/groupnorm.py:17-27); here the build is an explicit step
(`python -m sduss_b200.build` or `__graft_entry__.build()`), and the product path fails
loudly when the library is missing.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsduss_b200.so")
OBJ = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(src, obj):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "sduss_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile(src):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    if not _stale(os.path.join(CSRC, src), obj):
        return src, 0, "(up to date)"
    extra = os.environ.get("SDUSS_B200_NVCC_EXTRA", "").split()
    cmd = ["nvcc", *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    return src, p.returncode, p.stdout + p.stderr


def build(verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(_compile, srcs))
    log = []
    for src, rc, out in results:
        log.append(f"==== {src}\n{out}")
        if rc != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src}")
    with open(os.path.join(OBJ, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in srcs]
    if (not os.path.exists(OUT)) or any(os.path.getmtime(o) > os.path.getmtime(OUT) for o in objs):
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, *objs,
               "-Xcompiler", "-fPIC", "-cudart", "static"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            sys.stderr.write(p.stdout + p.stderr)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
