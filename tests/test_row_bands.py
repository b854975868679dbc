"""CPU test of the row-band partition of the multi-GPU solver (csrc/srb_row_bands.h): pure host arithmetic, compiled
with g++ and run here.  The GPU side (tests/test_gpu_multi_solver.py) runs the solver on the bands it plans."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bands_tile_the_range_and_pulls_cover_the_halo(tmp_path):
    out = str(tmp_path / "row_bands_check")
    src = os.path.join(ROOT, "tests", "row_bands_check.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++14", "-Wall", "-Wextra", "-o", out, src])
    res = subprocess.run([out], capture_output=True, text=True, timeout=600)
    sys.stdout.write(res.stdout)
    assert res.returncode == 0 and res.stdout.startswith("OK"), res.stdout + res.stderr[-2000:]
