"""Importable alias of the package directory ``pointcloud-segmentation-attention_b200`` (a hyphen cannot appear in
an ``import`` statement).  ``import pcops_b200`` and ``from pcops_b200.tf_grouping import ...`` both work."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("pointcloud-segmentation-attention_b200")
for _name in ("_lib", "tf_sampling", "tf_grouping", "tf_interpolate", "attention_layer", "pointnet_util", "pipeline", "synth",
              "sharding", "complete_scene_loader"):
    sys.modules[__name__ + "." + _name] = importlib.import_module("pointcloud-segmentation-attention_b200." + _name)
sys.modules[__name__] = _pkg
