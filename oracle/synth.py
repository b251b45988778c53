"""Re-export of the product package's workload generator (pointcloud-segmentation-attention_b200/synth.py) so the
oracle side consumes exactly the bytes the CUDA side does.  Loaded by file path: the oracle never imports the
product package (and the product never imports the oracle)."""
import importlib.util
import os

_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pointcloud-segmentation-attention_b200", "synth.py")
_spec = importlib.util.spec_from_file_location("_pcops_synth", _path)
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
