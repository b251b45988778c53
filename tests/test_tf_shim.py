"""The TensorFlow op shim (integration/tf_shim) must parse as C++ against the C ABI and re-register exactly the op
names / attrs / inputs / outputs of the reference's three custom-op libraries (tf_sampling.cpp:14-63,
tf_grouping.cpp:13-63, tf_interpolate.cpp:12-46), so the reference's Python wrappers load it unchanged."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "integration", "tf_shim")

# op -> (attrs, inputs, outputs) as registered by the reference
REFERENCE_REGISTRY = {
    "ProbSample": ([], ["inp: float32", "inpr: float32"], ["out: int32"]),
    "FarthestPointSample": (["npoint: int"], ["inp: float32"], ["out: int32"]),
    "GatherPoint": ([], ["inp: float32", "idx: int32"], ["out: float32"]),
    "GatherPointGrad": ([], ["inp: float32", "idx: int32", "out_g: float32"], ["inp_g: float32"]),
    "QueryBallPoint": (["radius: float", "nsample: int"], ["xyz1: float32", "xyz2: float32"],
                       ["idx: int32", "pts_cnt: int32"]),
    "SelectionSort": (["k: int"], ["dist: float32"], ["outi: int32", "out: float32"]),
    "GroupPoint": ([], ["points: float32", "idx: int32"], ["out: float32"]),
    "GroupPointGrad": ([], ["points: float32", "idx: int32", "grad_out: float32"], ["grad_points: float32"]),
    "ThreeNN": ([], ["xyz1: float32", "xyz2: float32"], ["dist: float32", "idx: int32"]),
    "ThreeInterpolate": ([], ["points: float32", "idx: int32", "weight: float32"], ["out: float32"]),
    "ThreeInterpolateGrad": ([], ["points: float32", "idx: int32", "weight: float32", "grad_out: float32"],
                             ["grad_points: float32"]),
}


def parse_registry():
    reg = {}
    for f in ("sampling_ops.cc", "grouping_ops.cc", "interpolation_ops.cc"):
        text = open(os.path.join(SHIM, f)).read()
        for m in re.finditer(r'REGISTER_OP\("(\w+)"\)(.*?)\.SetShapeFn', text, re.S):
            body = m.group(2)
            reg[m.group(1)] = (re.findall(r'\.Attr\("([^"]+)"\)', body), re.findall(r'\.Input\("([^"]+)"\)', body),
                               re.findall(r'\.Output\("([^"]+)"\)', body))
        for op in re.findall(r'REGISTER_KERNEL_BUILDER\(Name\("(\w+)"\)\.Device\(DEVICE_GPU\)', text):
            assert op in reg, op
    return reg


REF_TF_OPS = "/root/reference/pointnet2_tensorflow/tf_ops"
REF_SOURCES = ("sampling/tf_sampling.cpp", "grouping/tf_grouping.cpp", "interpolation_3d/tf_interpolate.cpp")


def parse_reference_registry():
    """The op registry as the reference's own sources state it (REGISTER_OP blocks of tf_sampling.cpp:14-63,
    tf_grouping.cpp:13-63, tf_interpolate.cpp:12-46), parsed where /root/reference is mounted."""
    reg = {}
    for f in REF_SOURCES:
        text = open(os.path.join(REF_TF_OPS, f)).read()
        for m in re.finditer(r'REGISTER_OP\("(\w+)"\)(.*?)\.SetShapeFn', text, re.S):
            body = m.group(2)
            reg[m.group(1)] = (re.findall(r'\.Attr\("([^"]+)"\)', body), re.findall(r'\.Input\("([^"]+)"\)', body),
                               re.findall(r'\.Output\("([^"]+)"\)', body))
    return reg


def test_shim_registers_the_reference_ops_exactly():
    assert parse_registry() == REFERENCE_REGISTRY


def test_hand_typed_registry_equals_the_one_parsed_from_the_reference_sources():
    """REFERENCE_REGISTRY above travels to the GPU box (no /root/reference there); here, where the reference is
    mounted, it is checked against the REGISTER_OP blocks of the reference's own .cpp files, and so is the shim."""
    import pytest
    if not os.path.isdir(REF_TF_OPS):
        pytest.skip("/root/reference is not mounted")
    ref = parse_reference_registry()
    assert ref == REFERENCE_REGISTRY
    assert parse_registry() == ref


def test_every_registered_op_has_a_gpu_kernel_calling_the_c_abi():
    text = "".join(open(os.path.join(SHIM, f)).read() for f in ("sampling_ops.cc", "grouping_ops.cc", "interpolation_ops.cc"))
    kernels = set(re.findall(r'REGISTER_KERNEL_BUILDER\(Name\("(\w+)"\)\.Device\(DEVICE_GPU\)', text))
    assert kernels == set(REFERENCE_REGISTRY)
    for fn in ("pc_prob_sample", "pc_fps", "pc_gather_point", "pc_gather_point_grad", "pc_query_ball_grid", "pc_selection_sort",
               "pc_group_point", "pc_group_point_grad", "pc_three_nn_grid", "pc_three_interpolate",
               "pc_three_interpolate_grad"):
        assert re.search(r"\b%s\(" % fn, text), fn


def test_shim_parses_against_the_c_abi_header():
    out = subprocess.run(["sh", os.path.join(SHIM, "check.sh")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert out.stdout.count("ok:") == 4


def test_attention_shim_registers_gpu_kernels_over_the_c_abi():
    """attention_ops.cc adds ops the reference does not have (its AttentionLayer composes stock TF ops): the contraction,
    its gradient, and the whole layer on the tensor cores.  They must be GPU kernels that call the C ABI."""
    text = open(os.path.join(SHIM, "attention_ops.cc")).read()
    ops = set(re.findall(r'REGISTER_OP\("(\w+)"\)', text))
    assert ops == {"PointAttentionContract", "PointAttentionContractGrad", "PointAttentionLayer"}
    assert set(re.findall(r'REGISTER_KERNEL_BUILDER\(Name\("(\w+)"\)\.Device\(DEVICE_GPU\)', text)) == ops
    for fn in ("pc_attention_fwd", "pc_attention_bwd", "pc_attention_layer_fwd", "pc_attention_layer_workspace_bytes"):
        assert re.search(r"\b%s\(" % fn, text), fn
