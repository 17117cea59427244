"""CPU check of the leaf-id encoding of the degree-2 descent records (csp_encode_leaf_runs, csrc/common.cuh) against the decode
rule of the descent kernels: exhaustive over the child patterns of a node (tests/cpp/leaf_runs.cpp, compiled with the host
compiler against the CUDA headers; no GPU)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_leaf_run_encoding_roundtrip(tmp_path):
    cxx = shutil.which("g++")
    inc = "/usr/local/cuda/include"
    if not cxx or not os.path.exists(os.path.join(inc, "cuda_runtime.h")):
        pytest.skip("host compiler or CUDA headers not available")
    exe = str(tmp_path / "leaf_runs")
    subprocess.run([cxx, "-O1", "-std=c++17", "-I", inc, "-I", os.path.join(ROOT, "coordinatesplittingptrees.jl_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "leaf_runs.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    tag, checked, direct, table = out.stdout.split()
    assert tag == "ok" and int(checked) > 500 and int(direct) > int(table) > 0
