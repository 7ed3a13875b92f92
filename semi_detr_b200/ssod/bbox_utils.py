"""Weak-view -> strong-view box warp: ``Transform2D.transform_bboxes``
(detr_ssod/models/utils/bbox_utils.py:18-41, 167-192): corners -> homography -> axis-aligned hull -> clamp."""
import torch


def bbox2points(box):
    x1, y1, x2, y2 = box[:, 0:1], box[:, 1:2], box[:, 2:3], box[:, 3:4]
    return torch.cat([x1, y1, x2, y1, x2, y2, x1, y2], dim=1).reshape(-1, 2)


def points2bbox(point, max_w, max_h):
    point = point.reshape(-1, 4, 2)
    if point.size(0) == 0:
        return point.new_zeros(0, 4)
    lo, hi = point.min(dim=1)[0], point.max(dim=1)[0]
    return torch.stack([lo[:, 0].clamp(0, max_w), lo[:, 1].clamp(0, max_h),
                        hi[:, 0].clamp(0, max_w), hi[:, 1].clamp(0, max_h)], dim=1)


class Transform2D:
    @staticmethod
    def transform_bboxes(bbox, M, out_shape):
        if isinstance(bbox, (list, tuple)):
            assert len(bbox) == len(M)
            return [Transform2D.transform_bboxes(b, m, o) for b, m, o in zip(bbox, M, out_shape)]
        if bbox.shape[0] == 0:
            return bbox
        score = bbox[:, 4:] if bbox.shape[1] > 4 else None
        pts = bbox2points(bbox[:, :4])
        pts = torch.cat([pts, pts.new_ones(pts.shape[0], 1)], dim=1)
        pts = torch.matmul(M, pts.t()).t()
        pts = pts[:, :2] / pts[:, 2:3]
        out = points2bbox(pts, out_shape[1], out_shape[0])
        return out if score is None else torch.cat([out, score], dim=1)
