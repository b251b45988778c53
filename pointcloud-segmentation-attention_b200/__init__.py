"""pointcloud-segmentation-attention_b200 -- the PointNet++ geometry-op hot path as sm_100a CUDA kernels.

Import with ``importlib.import_module("pointcloud-segmentation-attention_b200")`` or through the alias module
``pcops_b200`` at the repository root.  Module names follow the reference's op wrappers so call sites read the same:

    from pcops_b200.tf_sampling import farthest_point_sample, gather_point
    from pcops_b200.tf_grouping import query_ball_point, group_point, knn_point, select_top_k
    from pcops_b200.tf_interpolate import three_nn, three_interpolate
    from pcops_b200.attention_layer import AttentionLayer, attention_contract
    from pcops_b200.pointnet_util import sample_and_group, sample_and_group_all, fp_interpolate
    from pcops_b200.sa_modules import PointnetSAModule, PointnetSAModuleAttention, SharedMLP, dense_layer
    from pcops_b200.model_pipeline import ScanNetAttentionModel        # the whole attention model, pre-allocated
    from pcops_b200.experimental_layers import AttentionNetLayer, PoolingAttentionNetLayer

Everything computes in libpcops.so (include/pcops.h); there is no CPU or eager fallback.
"""
from . import _lib  # noqa: F401
from .tf_sampling import (farthest_point_sample, farthest_point_sample_and_gather, gather_point,  # noqa: F401
                          gather_point_grad, prob_sample, cumsum)
from .tf_grouping import (group_point, group_point_grad, knn_point, query_ball_point,  # noqa: F401
                          select_top_k)
from .tf_interpolate import (three_interpolate, three_interpolate_grad, three_nn,  # noqa: F401
                             three_weights)
from .attention_layer import AttentionLayer, attention_contract  # noqa: F401
from .pointnet_util import fp_interpolate, sample_and_group, sample_and_group_all  # noqa: F401

__version__ = "0.1.0"


def library_path():
    return _lib.LIB_PATH
