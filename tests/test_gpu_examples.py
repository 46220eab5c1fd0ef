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


def test_reference_schema_scripts(cuda_lib):
    """examples/run_snr_ber_cuda.py and run_benchmark_cuda.py print the JSON schemas of the reference's run_snr_ber / run_benchmark"""
    import json
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "run_snr_ber_cuda.py"), "-c", "2", "-d", "HARD8", "-n", "2048", "-L", "1024",
                        "--max-points", "3", "--min-errors", "50"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1500:]
    res = json.loads(r.stdout)
    assert set(res[0]) == {"name", "decode_type", "simd_type", "K", "R", "G", "EbNo_dB", "ber"}        # run_snr_ber.cpp:425-440
    assert res[0]["ber"][0] > res[0]["ber"][-1] > 0
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "run_benchmark_cuda.py"), "-c", "2", "-d", "SOFT16", "-n", "4096", "-L", "1024", "-T", "3"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1500:]
    res = json.loads(r.stdout)
    for key in ("name", "decode_type", "simd_type", "K", "R", "G", "total_input_bits", "total_symbols", "update_symbols_ns", "chainback_bits_ns"):
        assert key in res[0]                                                                           # run_benchmark.cpp:308-325
    assert len(res[0]["update_symbols_ns"]) == 3 and min(res[0]["update_symbols_ns"]) > 0
