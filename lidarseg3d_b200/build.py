"""Build recipe: every csrc/*.cu -> lidarseg3d_b200/_ls3d.so with plain nvcc for sm_100a (no torch headers)."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--extended-lambda", "-std=c++17",
         "-Xcompiler", "-fPIC"]


def build(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))
    deps = srcs + glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + [os.path.join(HERE, "..", "include", "ls3d.h")]
    out = os.path.join(HERE, "_ls3d.so")
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    objdir = os.path.join(HERE, "csrc", "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if (not force) and os.path.exists(o) and all(os.path.getmtime(o) >= os.path.getmtime(d) for d in
                                                      [s] + deps[len(srcs):]):
            continue
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{log}")
        if verbose:
            print(log)
    subprocess.check_call([NVCC, "-shared", "-o", out] + objs + ["-lcudart"])
    return out


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
