"""The C++ facade (include/viterbi_cuda/viterbi_decoder_cuda.h) compiled and run the way the reference's CI runs its example
programs (exit code = verdict): examples/run_simple_cuda.cpp mirrors examples/run_simple.cpp."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_run_simple_cuda(cuda_lib, tmp_path):
    exe = str(tmp_path / "run_simple_cuda")
    libdir = os.path.join(ROOT, "viterbidecodercpp_b200", "csrc")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "run_simple_cuda.cpp"),
                           "-L" + libdir, "-lviterbi_b200", "-Wl,-rpath," + libdir, "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "error_metric=0" in r.stdout and "0/8192 incorrect bits" in r.stdout      # the reference prints the same two lines
    assert "batch: 0 mismatching bytes" in r.stdout
