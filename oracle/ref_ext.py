"""Loader for the reference's own pointnet2 CUDA extension built by oracle/build_ref.py.

TEST INFRASTRUCTURE ONLY.  Returns the `_ext` module (9 ops over CUDA torch tensors,
/root/reference/lib/pointnet2/_ext_src/src/bindings.cpp:6-19) or None when oracle/_ref
is absent.  It only executes on a machine with a GPU.
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
_mod = None
_tried = False


def load():
    global _mod, _tried
    if _tried:
        return _mod
    _tried = True
    root = os.path.join(HERE, "_ref")
    if not os.path.exists(os.path.join(root, "pointnet2_ref", "_ext.so")):
        return None
    import torch  # noqa: F401  (libtorch must be loaded first)
    if root not in sys.path:
        sys.path.insert(0, root)
    try:
        _mod = importlib.import_module("pointnet2_ref._ext")
    except Exception as e:  # ABI drift etc.: report, do not hide
        print("oracle/_ref present but not importable: %r" % (e,))
        _mod = None
    return _mod
