"""examples/band_two_process.cpp: a C++ host that drives the band-sharded path through the C ABI alone (fork, CUDA IPC,
shared-memory barrier) — compiles here, runs on a box with >= 2 GPUs."""
import os
import subprocess

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
BIN = os.path.join(ROOT, "examples", "_build", "band_two_process")


def _compile(built_lib):
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "band_two_process.cpp"),
           built_lib, "-Wl,-rpath,$ORIGIN/../../painty_b200", "-o", BIN]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_band_example_compiles_against_the_c_abi_only(built_lib):
    _compile(built_lib)
    assert os.path.exists(BIN)
    src = open(os.path.join(ROOT, "examples", "band_two_process.cpp")).read()
    assert "torch" not in src.replace("no torch", "") and "nccl.h" not in src and "cuda_runtime" not in src


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_band_example_renders_the_single_gpu_image(built_lib, world):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    if not os.path.exists(BIN):
        _compile(built_lib)
    r = subprocess.run([BIN, str(world)], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "-> OK" in r.stdout, r.stdout + r.stderr
