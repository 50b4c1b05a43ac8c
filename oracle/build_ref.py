"""Build the UNMODIFIED reference pointnet2 CUDA extension into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under ``bridgeqa_b200/`` imports this.

The reference's own ``lib/pointnet2/setup.py`` cannot be used: it pins
``TORCH_CUDA_ARCH_LIST`` to sm_37..sm_75 (setup.py:17), which nvcc 12.9
rejects.  So the 4 ``.cu`` + 5 ``.cpp`` files are compiled *where they lie*
under /root/reference (no source is copied into this repo) with the same -O3
flags the reference asks for, for sm_100a, and linked into

    oracle/_ref/pointnet2_ref/_ext.so

which imports as ``pointnet2_ref._ext`` once ``oracle/_ref`` is on sys.path
(the loader in oracle/ref_ext.py also aliases it to ``pointnet2._ext``, the
name the reference's pointnet2_utils.py:26 imports).

The extension is CUDA-only (sampling.cpp:83 "CPU not supported"), so it is
built here (no GPU) and *executed* only on the GPU box, where it is the pin
for both the C oracle (oracle/pointnet2_oracle.c) and the product kernels.
oracle/_ref/ is git-ignored but travels with gpurun.
"""
import glob
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/lib/pointnet2/_ext_src"
OUT_DIR = os.path.join(HERE, "_ref", "pointnet2_ref")
OBJ_DIR = os.path.join(HERE, "_ref", "obj")
SO_PATH = os.path.join(OUT_DIR, "_ext.so")


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def available():
    return os.path.exists(SO_PATH)


# The reference's Python layer over the extension, staged UNMODIFIED (byte-for-byte copies, made at
# build time into the git-ignored oracle/_ref/ only) so that the GPU box -- where /root/reference does
# not exist -- can run the stock reference modules as the checker / the reference-on-GPU timing leg.
REF_ROOT = "/root/reference"
TREE_DIR = os.path.join(HERE, "_ref", "ref_tree")
STAGED = ("lib/pointnet2/pointnet2_modules.py", "lib/pointnet2/pointnet2_utils.py",
          "lib/pointnet2/pytorch_utils.py", "models/backbone_module.py", "models/voting_module.py",
          "lib/loss_helper.py", "utils/nn_distance.py")


def stage_python_layer():
    """Copies STAGED into oracle/_ref/ref_tree (same relative paths).  Returns the tree dir or None."""
    import shutil
    if not os.path.isdir(REF_ROOT):
        return TREE_DIR if tree_available() else None
    for rel in STAGED:
        dst = os.path.join(TREE_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_ROOT, rel), dst)
    return TREE_DIR


def tree_available():
    return all(os.path.exists(os.path.join(TREE_DIR, rel)) for rel in STAGED)


def build(force=False, verbose=True):
    """Returns the .so path, or None when /root/reference is absent (GPU box)."""
    if not os.path.isdir(REF_SRC):
        return SO_PATH if available() else None
    srcs = sorted(glob.glob(REF_SRC + "/src/*.cpp") + glob.glob(REF_SRC + "/src/*.cu"))
    stage_python_layer()
    if available() and not force:
        newest = max(os.path.getmtime(s) for s in srcs)
        if os.path.getmtime(SO_PATH) >= newest:
            return SO_PATH
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    incs = ["-I" + REF_SRC + "/include"]
    for p in ce.include_paths("cuda"):
        incs += ["-isystem", p]
    incs += ["-isystem", sysconfig.get_paths()["include"]]
    defs = ["-DTORCH_EXTENSION_NAME=_ext", "-DTORCH_API_INCLUDE_EXTENSION_H",
            "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src) + ".o")
        if src.endswith(".cu"):
            cmd = [cuda_home + "/bin/nvcc", "-O3", "-std=c++17", "-gencode",
                   "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                   "--expt-relaxed-constexpr", "-c", src, "-o", obj] + incs + defs
        else:
            cmd = ["g++", "-O3", "-std=c++17", "-fPIC", "-c", src, "-o", obj] + incs + defs
        _run(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(compile_one, srcs))
    libdirs = ce.library_paths("cuda")
    link = ["g++", "-shared", "-o", SO_PATH] + objs
    for d in libdirs:
        link += ["-L" + d, "-Wl,-rpath," + d]
    link += ["-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lcudart",
             "-lc10_cuda", "-ltorch_cuda"]
    _run(link)
    with open(os.path.join(OUT_DIR, "__init__.py"), "w") as f:
        f.write("# built by oracle/build_ref.py from /root/reference/lib/pointnet2/_ext_src\n")
    if verbose:
        print("built", SO_PATH)
    return SO_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p if p else "reference sources absent and no prebuilt oracle/_ref")
