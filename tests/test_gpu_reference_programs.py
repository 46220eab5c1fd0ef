"""The reference's OWN programs driving the CUDA backend: examples/run_tests.cpp, run_punctured_decoder.cpp
and run_simple.cpp are compiled UNMODIFIED (included from /root/reference by the wrappers in oracle/ref_programs/, which only add a
SIMD_CUDA entry to the decoder dispatch of examples/helpers/simd_type.h:21-112) into oracle/_ref/ by oracle/Makefile, in the build
container; the binaries travel to the GPU box with the repository snapshot and are run here.  The CUDA decoder class they use
(include/viterbi_cuda/viterbi_decoder_cuda_ref.h) works on the reference's own ViterbiDecoder_Core object, so get_error and chainback
are the reference's code running on GPU-made metrics and decision rows.  Skipped where the binaries were not built."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
ANSI = re.compile(r"\x1b\[[0-9;]*m")


def run_program(name, *args):
    path = os.path.join(REF, name)
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (needs /root/reference at build time)")
    p = subprocess.run([path, *args], capture_output=True, text=True, timeout=600)
    return p.returncode, ANSI.sub("", p.stdout + p.stderr)


def test_reference_run_tests_passes_with_the_cuda_decoder(cuda_lib):
    """run_tests.cpp:133-191: 8 codes x 3 decode types; every case is also decoded by SIMD_CUDA with 0 bit errors and error metric 0
    (noise free), except Cassini SOFT8 which is skipped for SIMD_CUDA for the reason the reference skips it for SCALAR"""
    rc, out = run_program("run_tests_cuda")
    assert rc == 0, out[-3000:]
    cuda_lines = [l for l in out.splitlines() if "SIMD_CUDA" in l]
    passed = [l for l in cuda_lines if l.lstrip().startswith("PASS")]
    skipped = [l for l in cuda_lines if l.lstrip().startswith("SKIP")]
    assert len(passed) == 23 and len(skipped) == 1, "\n".join(cuda_lines)
    host = [l for l in out.splitlines() if l.lstrip().startswith("PASS") and "SIMD_CUDA" not in l]
    assert len(host) >= 23            # the host decoders of the same run (SCALAR always; SSE / AVX where the box has them)


def test_reference_run_punctured_decoder_passes_with_the_cuda_decoder(cuda_lib):
    """run_punctured_decoder.cpp:139-190: DAB FIC, 774 update calls per frame through decode_punctured_symbols; SIMD_CUDA reports
    the reference's known traceback errors (100584 / 2376 / 792) like every other decoder"""
    rc, out = run_program("run_punctured_decoder_cuda")
    assert rc == 0, out[-3000:]
    lines = out.splitlines()
    errors, bits = [], []
    for i, l in enumerate(lines):
        if "SIMD_CUDA results" in l:
            block = lines[i + 1:i + 4]
            errors += [x.split("=")[1].strip() for x in block if x.startswith("traceback_error=")]
            bits += [x for x in block if "incorrect bits" in x]
    assert sorted(errors, key=int) == ["792", "2376", "100584"], (errors, out[-2000:])
    assert len(bits) == 3 and all(x.startswith("0/") for x in bits), bits
    assert "PASSED 12/12 TESTS" in out


def test_reference_run_simple_decodes_on_the_gpu(cuda_lib):
    """run_simple.cpp with its decoder class re-pointed at the CUDA decoder: exit code 0 = no decoding errors"""
    rc, out = run_program("run_simple_cuda")
    assert rc == 0, out[-2000:]
    assert "0/" in out and "incorrect bits" in out


def test_facade_accepts_the_reference_objects(cuda_lib):
    """converting constructors from ::ViterbiBranchTable / ::ViterbiDecoder_Config, table operator[] / data(), and the m_metrics /
    m_decisions / m_current_decoded_bit accessors against the reference object's public members"""
    rc, out = run_program("facade_from_reference_cuda")
    assert rc == 0, out[-3000:]
    assert out.count("PASS") == 5


def _json_entries(text):
    import json
    start = text.index("[")
    return json.loads(text[start:text.rindex("]") + 1])


def test_reference_run_snr_ber_with_the_cuda_decoder(cuda_lib):
    """run_snr_ber.cpp:255-275 from its thread pool (4 threads, one GPU handle per thread): with the same seed the SIMD_CUDA curve
    must equal the SCALAR curve point for point (same generated frames, bit-exact decoder)"""
    rc, out = run_program("run_snr_ber_cuda", "-c", "2", "-d", "soft16", "-s", "scalar", "-s", "simd_cuda", "-t", "4", "-n", "300", "-D", "4",
                          "-L", "256", "-S", "7", "-k", "0.2")
    assert rc == 0, out[-3000:]
    entries = {e["simd_type"]: e for e in _json_entries(out)}
    assert set(entries) == {"SCALAR", "SIMD_CUDA"}, list(entries)
    a, b = entries["SCALAR"], entries["SIMD_CUDA"]
    assert len(a["ber"]) == len(b["ber"]) > 0
    assert a["EbNo_dB"] == b["EbNo_dB"]
    assert a["total_bit_errors"] == b["total_bit_errors"] if "total_bit_errors" in a else a["ber"] == b["ber"]


def test_reference_run_benchmark_with_the_cuda_decoder(cuda_lib):
    """run_benchmark.cpp:188-245 timing the streaming calls of the CUDA decoder class (reported, not judged: compatibility path)"""
    rc, out = run_program("run_benchmark_cuda", "-c", "2", "-d", "hard8", "-s", "simd_cuda", "-T", "0.3")      # (-M is listed in its usage but not parsed)
    assert rc == 0, out[-3000:]
    entries = _json_entries(out)
    assert len(entries) == 1 and entries[0]["simd_type"] == "SIMD_CUDA", entries
