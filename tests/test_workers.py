"""CPU test of the helper-thread pool behind the multi-GPU solver (csrc/srb_workers.h): compiled with g++ and run
here -- the pool contains no CUDA.  The GPU side (tests/test_gpu_multi_solver.py) exercises it with real work."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def binary(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("workers") / "workers_check")
    src = os.path.join(ROOT, "tests", "workers_check.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++14", "-pthread", "-Wall", "-Wextra", "-o", out, src])
    return out


@pytest.mark.parametrize("helpers,rounds", [(0, 2000), (1, 20000), (3, 20000), (7, 20000)])
def test_fork_join_rounds(binary, helpers, rounds):
    res = subprocess.run([binary, str(helpers), str(rounds)], capture_output=True, text=True, timeout=300)
    sys.stdout.write(res.stdout)
    assert res.returncode == 0 and res.stdout.startswith("OK"), res.stdout + res.stderr


def test_thread_sanitizer_clean(binary, tmp_path):
    """The same check under ThreadSanitizer when the toolchain has it (skipped otherwise)."""
    out = str(tmp_path / "workers_tsan")
    src = os.path.join(ROOT, "tests", "workers_check.cpp")
    build = subprocess.run(["g++", "-O1", "-g", "-std=c++14", "-pthread", "-fsanitize=thread", "-o", out, src],
                           capture_output=True, text=True)
    if build.returncode != 0:
        pytest.skip("g++ -fsanitize=thread not available here")
    res = subprocess.run([out, "3", "3000"], capture_output=True, text=True, timeout=300)
    if "FATAL: ThreadSanitizer" in res.stderr or "unexpected memory mapping" in res.stderr:
        pytest.skip("ThreadSanitizer cannot run in this container")
    assert res.returncode == 0 and "WARNING: ThreadSanitizer" not in res.stderr, res.stdout + res.stderr[-2000:]
