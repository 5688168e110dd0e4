"""Build recipe: every csrc/*.cu -> lidarseg3d_b200/_ls3d.so with plain nvcc for sm_100a (no torch headers)."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--extended-lambda", "-std=c++17",
         "-Xcompiler", "-fPIC"]


def build(force=False, verbose=False, prof=False):
    """prof=True: development build with in-kernel cycle counters (-DLS3D_PROF) -> _ls3d_prof.so (never loaded by the
    package unless LS3D_PROF_SO=1 is set)."""
    if prof:
        return _build_prof()
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
    tmp = out + ".tmp"
    subprocess.check_call([NVCC, "-shared", "-o", tmp] + objs + ["-lcudart"])
    missing = _missing_symbols(tmp)
    if missing:
        os.remove(tmp)
        if not force:
            return build(force=True, verbose=verbose)
        raise RuntimeError(f"_ls3d.so does not export {missing}")
    os.replace(tmp, out)            # atomic: a concurrent snapshot never sees a half-written library
    return out


def _build_prof():
    srcs = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))
    out = os.path.join(HERE, "_ls3d_prof.so")
    subprocess.check_call([NVCC] + FLAGS + ["-DLS3D_PROF", "-shared", "-o", out] + srcs + ["-lcudart"])
    return out


def _missing_symbols(so_path):
    """Every `int ls3d_*(` prototype of include/ls3d.h must be exported."""
    import re
    hdr = open(os.path.join(HERE, "..", "include", "ls3d.h")).read()
    want = set(re.findall(r"^int (ls3d_\w+)\(", hdr, flags=re.M))
    nm = subprocess.run(["nm", "-D", so_path], stdout=subprocess.PIPE, text=True).stdout
    have = {ln.split()[-1] for ln in nm.splitlines() if " T " in ln}
    return sorted(want - have)


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
