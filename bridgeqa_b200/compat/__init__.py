"""Import-path shims that let the UNMODIFIED BridgeQA source tree run on these kernels.

The reference reaches its native code through two import paths
(/root/reference/lib/pointnet2/pointnet2_utils.py:26 `import pointnet2._ext as _ext`, and
models/backbone_module.py:9 / models/proposal_module.py:17-18
`from lib.pointnet2.pointnet2_modules import ...`).  `install()` registers replacements in
sys.modules BEFORE the reference modules are imported:

    level="ext"      only `pointnet2._ext` is replaced (the 9 ops).  The reference's own
                     pointnet2_utils.py / pointnet2_modules.py / pytorch_utils.py run unchanged
                     on top; every op goes through the C ABI, nothing is fused.
    level="modules"  additionally `lib.pointnet2.{pointnet2_utils,pointnet2_modules,
                     pytorch_utils}` (and the bare `pointnet2_utils` / `pytorch_utils` names the
                     reference also uses) resolve to bridgeqa_b200's mirrors, so
                     models/backbone_module.py, voting_module.py and proposal_module.py pick up
                     the fused tcgen05 SA kernels in eval mode.  state_dict keys are identical,
                     so checkpoints load either way.
"""
import sys
import types


def install(level="modules"):
    from .. import ext, pointnet2_modules, pointnet2_utils, pytorch_utils

    pkg = sys.modules.get("pointnet2")
    if pkg is None:
        pkg = types.ModuleType("pointnet2")
        pkg.__path__ = []
        sys.modules["pointnet2"] = pkg
    pkg._ext = ext
    sys.modules["pointnet2._ext"] = ext
    if level == "ext":
        return
    if level != "modules":
        raise ValueError("level must be 'ext' or 'modules'")
    for name, mod in (("pointnet2_utils", pointnet2_utils), ("pytorch_utils", pytorch_utils),
                      ("pointnet2_modules", pointnet2_modules)):
        sys.modules[name] = mod
        sys.modules["lib.pointnet2." + name] = mod
    lib = sys.modules.get("lib")
    if lib is None:
        lib = types.ModuleType("lib")
        lib.__path__ = []
        sys.modules["lib"] = lib
    sub = sys.modules.get("lib.pointnet2")
    if sub is None:
        sub = types.ModuleType("lib.pointnet2")
        sub.__path__ = []
        sys.modules["lib.pointnet2"] = sub
        lib.pointnet2 = sub
    sub.pointnet2_utils, sub.pointnet2_modules, sub.pytorch_utils = (
        pointnet2_utils, pointnet2_modules, pytorch_utils)
