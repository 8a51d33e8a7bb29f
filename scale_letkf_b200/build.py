"""In-tree nvcc build of libletkf_b200.so for sm_100a (cross-compiles without a GPU)."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "letkf_b200.cu")
OUT = os.path.join(HERE, "libletkf_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-extended-lambda", "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "177",
]


def sources():
    return sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")) + glob.glob(os.path.join(HERE, "csrc", "*.cuh"))
                  + [os.path.join(HERE, "..", "include", "letkf_b200.h")])


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False, out=None, defines=()):
    """out / defines: a variant build (e.g. out="libletkf_b200_x.so", defines=["LETKF_VARIANT_X"]) selected at run
    time with LETKF_B200_LIB, for A/B measurements; the default build is the product."""
    if out is None and not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    out = out or OUT
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, SRC]
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a wrapper without libgomp specs; nvcc only needs a host g++
    subprocess.check_call(cmd, env=env)
    return out


if __name__ == "__main__":
    import sys
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=os.path.join(HERE, outs[0]) if outs else None,
                defines=defs))
