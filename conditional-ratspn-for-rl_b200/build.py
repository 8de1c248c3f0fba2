"""Build libspn_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Every .cu of csrc/ is compiled to its own object (in parallel; objects are cached under build/obj keyed by source, headers
and flags) and the objects are linked into conditional-ratspn-for-rl_b200/lib/libspn_b200.so.

Experiment builds (timing studies only, never the product): ``SPN_TC_PROFILE=1`` (per-phase clock64 hooks),
``SPN_TC_DEBUG_MODE=<n>`` (compile-time switches that skip parts of the tensor-core kernels), ``SPN_MBAR_HINT_NS=<n>``;
``SPN_LIB_OUT=<path>`` writes the library somewhere else (load it with ``SPN_LIB_PATH``).
"""
import concurrent.futures
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "lib", "libspn_b200.so")
OBJ_DIR = os.path.join(ROOT, "build", "obj")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
         "-I" + os.path.join(ROOT, "include")]


def sources():
    return sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))


def headers():
    return sorted(glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h")))


def needs_build(out=OUT):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(p) > t for p in sources() + headers())


def extra_flags():
    extra = ["-DSPN_TC_PROFILE"] if os.environ.get("SPN_TC_PROFILE") else []
    if os.environ.get("SPN_LEAF_TC_PROF"):
        extra.append("-DSPN_LEAF_TC_PROF")
    for name in ("SPN_TC_DEBUG_MODE", "SPN_MBAR_HINT_NS"):
        if os.environ.get(name):
            extra.append(f"-D{name}={os.environ[name]}")
    return extra


def _compile(src, flags, verbose):
    h = hashlib.sha1()
    for p in [src] + headers():
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(flags).encode())
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + "-" + h.hexdigest()[:16] + ".o")
    log = ""
    if not os.path.exists(obj):
        cmd = [NVCC] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj + ".tmp", src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        os.replace(obj + ".tmp", obj)
        log = res.stderr
    return obj, log


def build(force=False, verbose=False, out=None):
    out = out or os.environ.get("SPN_LIB_OUT") or OUT
    if not force and not needs_build(out):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    flags = FLAGS + extra_flags()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 4)) as pool:
        results = list(pool.map(lambda s: _compile(s, flags, verbose), sources()))
    if verbose:
        print("".join(log for _, log in results))
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + [o for o, _ in results]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
