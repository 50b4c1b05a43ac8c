"""Build libbqa_pointnet2.so (hand-written sm_100a kernels + the C ABI) in-tree.

    python -m bridgeqa_b200.build [--force] [--verbose]

Plain nvcc, no torch / pybind / ATen in the translation units: the product boundary is
the C header include/bqa_pointnet2.h.  The .so is written next to this file
(bridgeqa_b200/lib/) so that it travels with the source tree and is what every Python
entry point loads; there is no JIT cache and no fallback.
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(HERE, "build")
SO_PATH = os.path.join(LIB_DIR, "libbqa_pointnet2.so")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    # keep the arithmetic exactly as written: explicit __fmaf_rn/__fmul_rn everywhere
    # parity matters, and no fast-math anywhere
    "-Xptxas", "-v",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [
        os.path.join(HERE, "..", "include", "bqa_pointnet2.h")]


def up_to_date():
    if not os.path.exists(SO_PATH):
        return False
    t = os.path.getmtime(SO_PATH)
    return all(os.path.getmtime(p) <= t for p in _deps() if os.path.exists(p))


def build(force=False, verbose=False):
    if not force and up_to_date():
        return SO_PATH
    if not os.path.exists(NVCC):
        raise RuntimeError("nvcc not found at %s and no prebuilt %s" % (NVCC, SO_PATH))
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    logs = {}

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "bqa_pointnet2.h")]
        if (not force and os.path.exists(obj)
                and all(os.path.getmtime(p) <= os.path.getmtime(obj) for p in [src] + hdrs)):
            return obj
        cmd = [NVCC] + NVCC_FLAGS + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        logs[src] = r.stdout
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout))
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    link = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO_PATH] + objs
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    if verbose:
        for src, log in logs.items():
            print("==", os.path.basename(src))
            print(log)
    with open(os.path.join(OBJ_DIR, "ptxas.log"), "w") as f:
        for src, log in logs.items():
            f.write("== %s\n%s\n" % (os.path.basename(src), log))
    return SO_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
