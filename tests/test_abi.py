"""The C-ABI library loads on a CPU-only box and exports exactly what include/pcops.h declares.
No kernel is launched here: only argument-rejection paths (which return before any CUDA call) are exercised."""
import ctypes
import os
import re
import subprocess

import pytest

import pcops_b200
from pcops_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pcops.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"PC_API[^;(]*?\b(pc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for must in ("pc_fps", "pc_gather_point", "pc_gather_point_grad", "pc_query_ball", "pc_group_point",
                 "pc_group_point_grad", "pc_selection_sort", "pc_knn", "pc_three_nn", "pc_three_interpolate",
                 "pc_three_interpolate_grad", "pc_attention_fwd", "pc_attention_bwd"):
        assert must in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build libpcops.so first (python -c 'import __graft_entry__ as g; g.build()')"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], check=True, capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    declared = declared_symbols()
    assert set(declared) <= exported, sorted(set(declared) - exported)
    # nothing but the C ABI leaks out of the library
    assert {s for s in exported if not s.startswith("pc_")} == set()
    # and the ctypes table binds every one of them
    assert set(_lib.SIGNATURES) == set(declared)


def test_library_loads_and_reports_version():
    L = _lib.lib()
    assert L.pc_version() >= 100
    assert L.pc_error_string(0) == b"PC_OK"
    assert L.pc_error_string(-1) == b"PC_ERR_INVALID_ARGUMENT"
    assert L.pc_error_string(-2) == b"PC_ERR_UNSUPPORTED"
    assert L.pc_error_string(-3) == b"PC_ERR_WORKSPACE"


def test_workspace_queries_are_pure_host_functions():
    L = _lib.lib()
    assert L.pc_fps_workspace_bytes(16, 8192, 1024) == 0            # state stays on chip
    assert L.pc_fps_workspace_bytes(2, 100000, 64) == 0               # one 16-CTA cluster per scene, still on chip
    assert L.pc_fps_workspace_bytes(2, 200000, 64) == 0  # up to 262144 points: one 16-CTA cluster per scene
    # beyond: a cooperative grid of 16384-point CTAs exchanging 24-byte records through global memory (2 parities)
    assert L.pc_fps_workspace_bytes(2, 300000, 64) == 2 * 2 * 19 * 24
    assert L.pc_fps_workspace_bytes(64, 1 << 20, 1024) == 2 * 2 * 64 * 24          # two scenes per launch
    assert L.pc_fps_workspace_bytes(2, 3000000, 64) == 2 * 3000000 * 4              # > 148 slices: streamed minima
    assert L.pc_group_point_grad_workspace_bytes(16, 1024, 64, 256, 32) == 16 * (1024 + 1 + 256 * 32) * 4
    assert L.pc_three_interpolate_grad_workspace_bytes(16, 8192, 128, 1024) == 16 * (1024 + 1 + 8192 * 3) * 4
    assert L.pc_gather_point_grad_workspace_bytes(4, 512, 128) == 4 * (512 + 1 + 128) * 4


def test_attribute_checks_reject_before_touching_the_device():
    L = _lib.lib()
    null = ctypes.c_void_p(0)
    # QueryBallPoint expects positive radius / nsample (tf_grouping.cpp:70-74)
    assert L.pc_query_ball(1, 8, 4, 0.0, 4, null, null, null, null, null) == _lib.PC_ERR_INVALID_ARGUMENT
    assert L.pc_query_ball(1, 8, 4, 0.1, 0, null, null, null, null, null) == _lib.PC_ERR_INVALID_ARGUMENT
    # SelectionSort expects positive k (tf_grouping.cpp:112-113)
    assert L.pc_selection_sort(1, 8, 4, 0, null, null, null, null) == _lib.PC_ERR_INVALID_ARGUMENT
    # FarthestPointSample: m <= 0 is a no-op (tf_sampling_g.cu:106-107); empty batches are no-ops
    assert L.pc_fps(4, 8, 0, null, null, null, null) == _lib.PC_OK
    assert L.pc_fps(0, 8, 4, null, null, null, null) == _lib.PC_OK
    assert L.pc_group_point(0, 8, 3, 4, 2, null, null, null, null) == _lib.PC_OK
    # outside the implemented envelope
    assert L.pc_knn(1, 8, 4, 200, 3, null, null, null, null, null) == _lib.PC_ERR_UNSUPPORTED
    assert L.pc_attention_fwd(4, 32, 16, 3, null, null, null, null, null) == _lib.PC_ERR_UNSUPPORTED


def test_wrappers_raise_reference_messages_and_have_no_cpu_fallback():
    import torch
    xyz = torch.zeros(2, 16, 3)
    with pytest.raises(ValueError, match="FarthestPointSample expects positive npoint"):
        pcops_b200.farthest_point_sample(0, xyz)
    with pytest.raises(ValueError, match=r"FarthestPointSample expects \(batch_size,num_points,3\) inp shape"):
        pcops_b200.farthest_point_sample(4, torch.zeros(2, 16, 4))
    with pytest.raises(ValueError, match="QueryBallPoint expects positive radius"):
        pcops_b200.query_ball_point(-1.0, 4, xyz, xyz)
    with pytest.raises(ValueError, match="QueryBallPoint expects positive nsample"):
        pcops_b200.query_ball_point(0.1, 0, xyz, xyz)
    with pytest.raises(ValueError, match=r"ThreeNN expects \(b,n,3\) xyz1 shape"):
        pcops_b200.three_nn(torch.zeros(2, 16), xyz)
    with pytest.raises(ValueError, match=r"GroupPoint expects \(batch_size, npoints, nsample\) idx shape"):
        pcops_b200.group_point(xyz, torch.zeros(3, 4, 2, dtype=torch.int32))
    with pytest.raises(ValueError, match=r"ThreeInterpolate expects \(b,n,3\) weight shape"):
        pcops_b200.three_interpolate(xyz, torch.zeros(2, 5, 3, dtype=torch.int32), torch.zeros(2, 6, 3))
    # CPU tensors are refused outright: there is no host implementation behind these names
    with pytest.raises(_lib.PcopsError, match="CUDA tensor"):
        pcops_b200.farthest_point_sample(4, xyz)
    with pytest.raises(ValueError, match=r"ProbSample expects \(batch_size,num_choices\) inp shape"):
        pcops_b200.prob_sample(xyz, xyz)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pointcloud-segmentation-attention_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh", ".cc")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("# oracle-free", ""), os.path.join(dirpath, f)


def test_library_reads_no_environment_variables():
    """include/pcops.h promises a re-entrant library without global knobs: no getenv in the kernel sources (the
    statically linked CUDA runtime imports the symbol for its own use, so the check is on the sources), no os.environ
    in the Python package."""
    pkg = os.path.join(ROOT, "pointcloud-segmentation-attention_b200")
    for f in os.listdir(os.path.join(pkg, "csrc")):
        assert "getenv" not in open(os.path.join(pkg, "csrc", f)).read(), f
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            assert "os.environ" not in open(os.path.join(pkg, f)).read(), f


def test_wrappers_reject_tensors_on_different_devices_without_a_gpu():
    """on_tensor_device only inspects CUDA tensors; on this CPU box the wrapper must still reach its own checks."""
    import torch
    with pytest.raises(_lib.PcopsError, match="CUDA tensor"):
        pcops_b200.gather_point(torch.zeros(1, 4, 3), torch.zeros(1, 2, dtype=torch.int32))


@pytest.mark.gpu
def test_ops_run_on_the_device_of_their_tensors_not_the_current_one():
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from oracle import cpu, synth
    xyz, _ = synth.scannet_batch(5, 2, 1024)
    torch.cuda.set_device(0)
    x1 = torch.from_numpy(xyz).to("cuda:1")
    idx = pcops_b200.farthest_point_sample(64, x1)
    assert idx.device == x1.device and np.array_equal(idx.cpu().numpy(), cpu.farthest_point_sample(64, xyz))
    nx = pcops_b200.gather_point(x1, idx)
    bi, _ = pcops_b200.query_ball_point(0.3, 16, x1, nx)
    assert np.array_equal(bi.cpu().numpy(), cpu.query_ball_point(0.3, 16, xyz, nx.cpu().numpy())[0])
    with pytest.raises(_lib.PcopsError, match="one CUDA device"):
        pcops_b200.gather_point(x1, idx.to("cuda:0"))
    assert torch.cuda.current_device() == 0
