"""The boundary is a plain C ABI: include/bhstep.h compiles as C99 and a C program drives libbhstep.so through dlopen."""
import os
import subprocess

import pytest

from gpu_nbody_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "c_abi_smoke")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", exe, "-ldl", "-lm"])
    return exe


def test_header_is_c99_and_library_loads_from_c(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe, "cpu", _lib.build()], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "c_abi_smoke ok" in out.stdout


@pytest.mark.gpu
def test_c_program_runs_the_step(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe, "gpu", _lib.build()], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "c_abi_smoke ok (gpu" in out.stdout
