"""Build a variant of libsduss_b200.so with extra -D flags on selected sources (kernel A/B
experiments): python tools/build_variant.py NAME attn_sm100.cu -DATT_TIMING ...
Writes sduss_b200/variants/libsduss_b200_NAME.so; select it with SDUSS_B200_LIB=<path>."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sduss_b200 import build as b


def build_variant(name, srcs, flags):
    b.build()
    vdir = os.path.join(b.HERE, "variants")
    os.makedirs(vdir, exist_ok=True)
    objs = []
    for s in b.sources():
        if s in srcs:
            obj = os.path.join(vdir, f"{name}_{s[:-3]}.o")
            cmd = ["nvcc", *b.NVCC_FLAGS, *flags, "-c", os.path.join(b.CSRC, s), "-o", obj]
            p = subprocess.run(cmd, capture_output=True, text=True)
            if p.returncode:
                sys.stderr.write(p.stdout + p.stderr)
                raise SystemExit(1)
            open(obj[:-2] + ".ptxas.log", "w").write(p.stdout + p.stderr)
            objs.append(obj)
        else:
            objs.append(os.path.join(b.OBJ, s[:-3] + ".o"))
    out = os.path.join(vdir, f"libsduss_b200_{name}.so")
    subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, *objs,
                    "-Xcompiler", "-fPIC", "-cudart", "static"], check=True)
    return out


if __name__ == "__main__":
    name = sys.argv[1]
    srcs = [a for a in sys.argv[2:] if a.endswith(".cu")]
    flags = [a for a in sys.argv[2:] if not a.endswith(".cu")]
    print(build_variant(name, srcs, flags))
