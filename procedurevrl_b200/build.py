"""In-tree build of libpvrl_sm100.so (the C-ABI extension declared in include/pvrl.h).

nvcc cross-compiles for sm_100a without a GPU, so this runs in the CPU build container; the resulting
.so is git-ignored but travels to the GPU box with the tree."""
import glob
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libpvrl_sm100.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libpvrl_sm100.so")
    return exe


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a (one object per file, in parallel) and link the shared library."""
    srcs = sources()
    headers = glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(PKG_DIR, "..", "include", "*.h"))
    obj_dir = os.path.join(PKG_DIR, "build")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    procs, objs = [], []
    for s in srcs:
        o = os.path.join(obj_dir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out.decode()}")
    if force or procs or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout.decode())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in os.sys.argv, verbose=True))
