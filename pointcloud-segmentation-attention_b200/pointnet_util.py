"""The geometry front-ends of the PointNet++ layers, mirroring pointnet2_tensorflow/utils/pointnet_util.py.

Only the parts of the layer library that sit ON the accelerated path are here: ``sample_and_group`` (:16-58),
``sample_and_group_all`` (:61-87) and the interpolation front half of ``pointnet_fp_module`` (:218-226).  The shared
MLPs, batch-norm and pooling that follow them in the reference are stock dense layers and are out of scope
(SURVEY.md 2.1, 8f).
"""
import torch

from .tf_grouping import group_point, knn_point, query_ball_point
from .tf_interpolate import three_interpolate, three_nn, three_weights
from .tf_sampling import farthest_point_sample, farthest_point_sample_and_gather, gather_point


def sample_and_group(npoint, radius, nsample, xyz, points, knn=False, use_xyz=True):
    """Same inputs / outputs as pointnet_util.py:16-58:
    returns new_xyz (B,npoint,3), new_points (B,npoint,nsample,3+C), idx (B,npoint,nsample), grouped_xyz."""
    if xyz.requires_grad:  # keep GatherPoint's registered gradient in the graph (tf_sampling.py:44-48)
        new_xyz = gather_point(xyz, farthest_point_sample(npoint, xyz))
    else:                  # one launch for the pair
        _, new_xyz = farthest_point_sample_and_gather(npoint, xyz)
    if knn:
        _, idx = knn_point(nsample, xyz, new_xyz)
    else:
        idx, _pts_cnt = query_ball_point(radius, nsample, xyz, new_xyz)
    grouped_xyz = group_point(xyz, idx)
    grouped_xyz = grouped_xyz - new_xyz.unsqueeze(2)  # translation normalisation (:40)
    if points is not None and points.numel() == 0 and points.dim() == 1:
        points = None  # the reference's Keras-layer convention (:41-42)
    if points is not None:
        grouped_points = group_point(points, idx)
        new_points = torch.cat([grouped_xyz, grouped_points], dim=-1) if use_xyz else grouped_points
    else:
        new_points = grouped_xyz
    return new_xyz, new_points, idx, grouped_xyz


def sample_and_group_all(xyz, points, use_xyz=True):
    """pointnet_util.py:61-87: one group holding every point, centroid (0,0,0)."""
    b, n, _ = xyz.shape
    new_xyz = torch.zeros((b, 1, 3), dtype=torch.float32, device=xyz.device)
    idx = torch.arange(n, dtype=torch.int32, device=xyz.device).reshape(1, 1, n).repeat(b, 1, 1)
    grouped_xyz = xyz.reshape(b, 1, n, 3)
    if points is not None:
        new_points = torch.cat([xyz, points], dim=2) if use_xyz else points
        new_points = new_points.unsqueeze(1)
    else:
        new_points = grouped_xyz
    return new_xyz, new_points, idx, grouped_xyz


def fp_interpolate(xyz1, xyz2, points1, points2):
    """Front half of pointnet_fp_module (pointnet_util.py:218-226): three_nn -> inverse-distance weights ->
    three_interpolate -> concat with the skip features.  Returns (B, n1, C2 [+ C1])."""
    dist, idx = three_nn(xyz1, xyz2)
    weight = three_weights(dist)
    interpolated = three_interpolate(points2, idx, weight)
    if points1 is not None:
        return torch.cat([interpolated, points1], dim=2)
    return interpolated
