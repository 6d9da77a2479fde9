"""Build libvpk.so (all csrc/*.cu) for sm_100a with nvcc, in-tree.

nvcc cross-compiles without a GPU.  Objects are rebuilt only when a source or
header is newer; `python -m vanishing_points_2017_b200.build [--force]`.
"""
import concurrent.futures
import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "csrc", "_obj")
LIB = os.path.join(PKG, "libvpk.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden,-ffp-contract=off",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
] + ["-D" + d for d in os.environ.get("VPK_DEFINES", "").split() if d]      # diagnostic builds, e.g. VPK_DEFINES=VPK_EM_MARKS


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; libvpk.so cannot be built")
    return p


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _flags_changed():
    """Objects built with other flags / VPK_DEFINES must not be linked silently: the flag string of the
    last build is kept next to the objects."""
    stamp = os.path.join(OBJ, "flags.txt")
    cur = " ".join(NVCC_FLAGS)
    old = open(stamp).read() if os.path.exists(stamp) else None
    if old != cur:
        with open(stamp, "w") as fh:
            fh.write(cur)
        return True
    return False


def build(force=False, verbose=False):
    nvcc = nvcc_path()
    os.makedirs(OBJ, exist_ok=True)
    force = _flags_changed() or force
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    headers = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(PKG, "..", "include", "vpk.h")]
    jobs = []
    objs = []
    for src in sources:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for src, r in ex.map(compile_one, jobs):
            log = os.path.join(OBJ, os.path.basename(src)[:-3] + ".ptxas.log")
            with open(log, "w") as fh:
                fh.write(r.stderr)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed on %s" % src)
            if verbose:
                sys.stderr.write(r.stderr)
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
