"""Detection post-processing on the device: the point-in-box counts and the
greedy 3-D NMS that the reference's `parse_predictions` (lib/ap_helper.py:40-178) runs as Python /
NumPy loops on the host, as two sm_100a kernels (csrc/postprocess.cu).

    nms_3d_faster(boxes (K,7), thr, old_type)            -> list of picked indices   (utils/nms.py:74-111)
    nms_3d_faster_samecls(boxes (K,8), thr, old_type)    -> list of picked indices   (utils/nms.py:113-152)
    nms_3d_batch(boxes (B,K,7|8), ...)                   -> pick mask (B,K), pick order (B,K)
    count_points_in_boxes(xyz (B,N,3), lo_hi (B,K,6))    -> (B,K) int32
    prediction_mask(data_dict, ...)                      -> pred_mask (B,K)   (ap_helper.py:86-178, 3-D NMS branches)
"""
import ctypes

import torch

from . import _native as N

_f32 = torch.float32


def nms_3d_batch(boxes, iou_threshold, old_type=False, same_class=False, valid=None):
    """boxes (B,K,7) = x1,y1,z1,x2,y2,z2,score (+ class as 8th column when same_class)."""
    if boxes.dim() != 3 or boxes.size(2) not in (7, 8):
        raise RuntimeError("boxes must be (B, K, 7) or (B, K, 8)")
    if same_class and boxes.size(2) != 8:
        raise RuntimeError("same_class NMS needs the class in column 7")
    b, k, c = boxes.shape
    packed = torch.zeros((b, k, 8), dtype=_f32, device=boxes.device)
    packed[..., :c] = boxes
    N.check_tensor(packed, "boxes", _f32)
    pick = torch.empty((b, k), dtype=torch.int32, device=boxes.device)
    order = torch.empty((b, k), dtype=torch.int32, device=boxes.device)
    vptr = None
    if valid is not None:
        valid = valid.to(torch.int32).contiguous()
        vptr = valid
    with torch.cuda.device(boxes.device):
        N.call("bqa_nms3d", b, k, N.ptr(packed), N.ptr(vptr), ctypes.c_double(float(iou_threshold)),
               1 if old_type else 0, 1 if same_class else 0, N.ptr(pick), N.ptr(order),
               N.stream_ptr(boxes.device))
    return pick.bool(), order


def _single(boxes, iou_threshold, old_type, same_class):
    t = torch.as_tensor(boxes, dtype=_f32)
    if not t.is_cuda:
        raise RuntimeError("boxes must be a CUDA tensor (CPU not supported)")
    _, order = nms_3d_batch(t.unsqueeze(0), iou_threshold, old_type, same_class)
    order = order[0]
    return order[order >= 0].tolist()


def nms_3d_faster(boxes, overlap_threshold, old_type=False):
    """utils/nms.py:74-111 for one scene: boxes (K,7) -> picked indices in pick order."""
    return _single(boxes, overlap_threshold, old_type, False)


def nms_3d_faster_samecls(boxes, overlap_threshold, old_type=False):
    """utils/nms.py:113-152: boxes (K,8) with the class in the last column."""
    return _single(boxes, overlap_threshold, old_type, True)


def count_points_in_boxes(xyz, lo_hi):
    """xyz (B,N,3), lo_hi (B,K,6) = x1,y1,z1,x2,y2,z2 -> (B,K) int32 points inside (bounds inclusive)."""
    N.check_tensor(xyz, "xyz", _f32)
    N.check_tensor(lo_hi, "lo_hi", _f32)
    b, n, _ = xyz.shape
    k = lo_hi.size(1)
    counts = torch.empty((b, k), dtype=torch.int32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        N.call("bqa_count_points_in_boxes", b, n, k, N.ptr(xyz), N.ptr(lo_hi), N.ptr(counts),
               N.stream_ptr(xyz.device))
    return counts


def prediction_mask(data_dict, nms_iou=0.25, remove_empty_box=True, use_old_type_nms=False, cls_nms=False,
                    min_points=5):
    """`pred_mask` (B,K) of parse_predictions (ap_helper.py:86-178) for use_3d_nms=True: drop boxes
    with fewer than `min_points` points inside (:89-100), then greedy 3-D NMS on the objectness
    probability, optionally per semantic class (:123-178).  Needs `bbox_corner` (B,K,8,3) as written
    by ProposalModule's on-device decode, `objectness_scores`, `sem_cls_scores`, `point_clouds`."""
    corners = data_dict["bbox_corner"]
    lo = corners.min(dim=2).values
    hi = corners.max(dim=2).values
    lo_hi = torch.cat([lo, hi], dim=-1).contiguous()
    valid = None
    if remove_empty_box:
        xyz = data_dict["point_clouds"][..., :3].contiguous()
        valid = count_points_in_boxes(xyz, lo_hi) >= min_points
    prob = torch.softmax(data_dict["objectness_scores"], dim=-1)[..., 1:2]
    cols = [lo_hi, prob]
    if cls_nms:
        cols.append(data_dict["sem_cls_scores"].argmax(-1, keepdim=True).to(_f32))
    pick, _ = nms_3d_batch(torch.cat(cols, dim=-1), nms_iou, use_old_type_nms, cls_nms, valid)
    return pick
