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


def load_reference_modules(ext_module):
    """The UNMODIFIED reference Python layer staged by build_ref.stage_python_layer(), imported on top
    of `ext_module` as `pointnet2._ext` (the reference's own extension from load(), or any object with
    the same 9 functions).  Returns (backbone_module, voting_module, pointnet2_utils) or None when the
    tree is not staged.  The staged files do `sys.path.append(os.path.join(os.getcwd(), "lib"))`
    (backbone_module.py:8) and `import pointnet2_utils` from their own directory, so the tree root is
    put on sys.path and used as cwd for the import."""
    import types
    tree = os.path.join(HERE, "_ref", "ref_tree")
    if not os.path.exists(os.path.join(tree, "models", "backbone_module.py")):
        return None
    pkg = types.ModuleType("pointnet2")
    pkg.__path__ = []
    pkg._ext = ext_module
    sys.modules["pointnet2"] = pkg
    sys.modules["pointnet2._ext"] = ext_module
    for stale in ("pointnet2_utils", "pointnet2_modules", "pytorch_utils", "lib", "lib.pointnet2",
                  "lib.pointnet2.pointnet2_modules", "lib.pointnet2.pointnet2_utils",
                  "lib.pointnet2.pytorch_utils", "models", "models.backbone_module", "models.voting_module"):
        sys.modules.pop(stale, None)
    cwd = os.getcwd()
    sys.path.insert(0, tree)
    os.chdir(tree)
    try:
        backbone_module = importlib.import_module("models.backbone_module")
        voting_module = importlib.import_module("models.voting_module")
        utils = sys.modules["pointnet2_utils"]
    finally:
        os.chdir(cwd)
    return backbone_module, voting_module, utils


def load_reference_loss_helper():
    """The UNMODIFIED lib/loss_helper.py (+ utils/nn_distance.py) of the reference, staged by
    build_ref.stage_python_layer(), or None.  Its other imports (lib.ap_helper, lib.loss, utils.box_util,
    icecream) are only used by the QA / evaluation functions of that file and are satisfied with empty
    stand-ins; the three detection losses run as written."""
    import types
    tree = os.path.join(HERE, "_ref", "ref_tree")
    if not os.path.exists(os.path.join(tree, "lib", "loss_helper.py")):
        return None

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    for stale in ("lib", "lib.loss_helper", "utils", "utils.nn_distance"):
        sys.modules.pop(stale, None)
    stub("icecream", ic=lambda *a, **k: None)
    lib = stub("lib")
    lib.__path__ = [os.path.join(tree, "lib")]
    stub("lib.ap_helper", parse_predictions=None)
    stub("lib.loss", SoftmaxRankingLoss=None)
    utils = stub("utils")
    utils.__path__ = [os.path.join(tree, "utils")]
    stub("utils.box_util", get_3d_box=None, get_3d_box_batch=None, box3d_iou=None, box3d_iou_batch=None)
    cwd = os.getcwd()
    sys.path.insert(0, tree)
    os.chdir(tree)
    try:
        return importlib.import_module("lib.loss_helper")
    finally:
        os.chdir(cwd)
        sys.path.remove(tree)
