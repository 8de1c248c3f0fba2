"""Build libspn_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "lib", "libspn_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
         "-shared", "-I" + os.path.join(ROOT, "include")]


def sources():
    return sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = sources() + glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    extra = ["-DSPN_TC_PROFILE"] if os.environ.get("SPN_TC_PROFILE") else []  # per-phase clock64 hooks (perturbs timing)
    if os.environ.get("SPN_TC_DEBUGSW"):
        extra.append("-DSPN_TC_DEBUGSW")  # SPN_TC_DEBUG / SPN_TC_N experiment switches inside the tensor-core kernels
    if os.environ.get("SPN_MBAR_HINT_NS"):
        extra.append("-DSPN_MBAR_HINT_NS=" + os.environ["SPN_MBAR_HINT_NS"])  # experiment: suspend hint of the mbarrier waits
    cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
