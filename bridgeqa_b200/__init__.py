"""bridgeqa_b200 -- B200 (sm_100a) implementation of BridgeQA's VoteNet point-cloud hot path.

Public surface (mirrors /root/reference/lib/pointnet2 and the three model files on the path):

    bridgeqa_b200.ext                 the 9 `pointnet2._ext` functions over the C ABI
    bridgeqa_b200.pointnet2_utils     furthest_point_sample, gather_operation, ball_query, ...
    bridgeqa_b200.pointnet2_modules   PointnetSAModuleVotes, PointnetFPModule, ...
    bridgeqa_b200.pytorch_utils       SharedMLP, BNMomentumScheduler, ...
    bridgeqa_b200.detector            Pointnet2Backbone, VotingModule, ProposalModule
    bridgeqa_b200.compat              import-path shims for the unmodified reference files

There is no CPU / eager fallback: every operator goes through libbqa_pointnet2.so.
"""
from .fused import set_fused, set_precision  # noqa: F401

__version__ = "0.1.0"
