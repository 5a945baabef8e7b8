"""CPU check of the host-side logic of the tcgen05 trailing update (pysfm_b200/csrc/ba_solve_tc.cuh):
the tile enumeration covers the lower triangle exactly once for every size, the instruction
descriptor encodes kind::i8 / S32 / M = 128 as cute::UMMA::InstrDescriptor lays it out, and the
pipeline depth fits the shared memory.  A small host program is compiled with nvcc and run here
(no GPU involved)."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT


def test_tile_enumeration_descriptor_and_stages(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "tc_tiles_check")
    src = os.path.join(ROOT, "tests", "host", "tc_tiles_check.cu")
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-o", exe, src],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0 and run.stdout.strip() == "ok", run.stdout + run.stderr
