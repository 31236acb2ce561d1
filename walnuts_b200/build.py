"""Build the in-tree CUDA shared library for sm_100a (nvcc cross-compiles without a GPU).

One object per translation unit in csrc/ (the kernel families of wn_dispatch.cuh), compiled in parallel, then
linked into walnuts_b200/_lib/libwalnuts_b200.so."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_DIR = os.path.join(HERE, "_lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB_PATH = os.path.join(LIB_DIR, "libwalnuts_b200.so")
SOURCES = [os.path.join(HERE, "csrc", f) for f in
           ("capi.cu", "wn_stats.cu", "wn_sched.cu", "plans_wpy.cu", "plans_nuts.cu", "plans_pkg.cu", "plans_adapt.cu", "plans_ext.cu")]
HEADERS = [os.path.join(HERE, "csrc", f) for f in
           ("wn_common.cuh", "wn_targets.cuh", "wn_walnutspy.cuh", "wn_package.cuh", "wn_dispatch.cuh", "wn_handle.hpp")] + \
          [os.path.join(ROOT, "include", "walnuts_cuda.h")]

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]


def _obj(src):
    return os.path.join(OBJ_DIR, os.path.splitext(os.path.basename(src))[0] + ".o")


def _stale(out, deps):
    if not os.path.isfile(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(f) > t for f in deps if os.path.isfile(f))


def needs_build():
    return _stale(LIB_PATH, SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile walnuts_b200/_lib/libwalnuts_b200.so.  Returns the library path."""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")

    def compile_one(src):
        obj = _obj(src)
        if not force and not _stale(obj, [src] + HEADERS):
            return obj
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose:
            sys.stderr.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on " + os.path.basename(src) + ":\n" + r.stdout + r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs,
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
