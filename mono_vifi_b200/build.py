"""Builds libmonovifi_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libmonovifi_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """`defines` / `out`: build a tuning variant (e.g. defines=["MVF_F1_FWD_MINB=3"], out="/path/lib_v.so")."""
    if out is not None or defines:
        return _build_variant(list(defines), out or SO, verbose)
    if not force and not stale():
        return SO
    objs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for s in sources():
        o = os.path.join(bdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        procs.append((s, subprocess.Popen([NVCC] + FLAGS + ["-c", s, "-o", o], stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % s)
    # cudart (and, for the conv kernels' tensor maps, the driver API) are linked dynamically
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", SO] + objs +
                          ["-lcudart"])
    return SO


def _build_variant(defines, out, verbose):
    tag = "_".join(d.replace("=", "-") for d in defines) or "default"
    bdir = os.path.join(HERE, "build", tag)
    os.makedirs(bdir, exist_ok=True)
    objs, procs = [], []
    for s in sources():
        o = os.path.join(bdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        procs.append((s, subprocess.Popen([NVCC] + FLAGS + ["-D" + d for d in defines] + ["-c", s, "-o", o],
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        outp, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(outp)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % s)
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs +
                          ["-lcudart"])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
