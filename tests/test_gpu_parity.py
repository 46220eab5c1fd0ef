"""GPU parity tests: the CUDA path (through the C ABI) against the scalar oracle on identical inputs, bit-exact:
decoded bytes, sum of renormalisation minima, final path metric, and - for the streaming API - every decision row and metric.
Mirrors the reference's own programs: run_simple, run_tests, run_punctured_decoder."""
import numpy as np
import pytest

import viterbidecodercpp_b200 as v
from viterbidecodercpp_b200 import synth
from common import CODE_BY_NAME, GPU_CODES, PAIR_CODES, assert_batch_equal, frames, make_cuda_decoder, make_oracle
from oracle_binding import MODE_SCALAR, MODE_SIMD

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("decode_type", ["SOFT16", "SOFT8", "HARD8"])
@pytest.mark.parametrize("name", GPU_CODES)
def test_noise_free_round_trip(cuda_lib, name, decode_type):
    """examples/run_tests.cpp:153-191: 64-byte frames, noise free, 0 bit errors; run_simple.cpp:81-93: error metric 0"""
    code = CODE_BY_NAME[name]
    dec, dc = make_cuda_decoder(code, decode_type)
    tx, sym = frames(code, dc, 70, 512, None, seed=7)
    ora, _ = make_oracle(code, decode_type)
    want = ora.decode_frames(sym, 70, 512)
    for lanes in dec.variants:
        dec.set_variant(lanes)
        out, acc, fin = dec.decode_batch(sym, 512)
        assert_batch_equal((out, acc, fin), want, f"{name} {decode_type} lanes/pair={lanes}")
        if name == "Cassini" and decode_type == "SOFT8":
            continue   # uint8 metrics wrap for K=15 (the reference skips this case too, run_tests.cpp:63-65); still bit-exact vs the oracle
        assert (out == tx).all(), lanes
        assert ((acc + fin) == 0).all(), lanes


@pytest.mark.parametrize("EbNo_dB", [-3.0, 1.0, 4.0])
@pytest.mark.parametrize("decode_type", ["SOFT16", "SOFT8", "HARD8"])
@pytest.mark.parametrize("name", GPU_CODES)
def test_batch_parity_awgn(cuda_lib, name, decode_type, EbNo_dB):
    """noisy frames: ties and renormalisations are exercised; ragged frame count (not a multiple of 64)"""
    code = CODE_BY_NAME[name]
    dec, dc = make_cuda_decoder(code, decode_type)
    ora, _ = make_oracle(code, decode_type)
    n_frames, L = 149, 1000 if decode_type != "SOFT16" else 2048
    L = (L // 8) * 8
    if code.K >= 9:
        n_frames, L = 75, 1000
    if code.K >= 15:
        n_frames, L = 5, 2048          # long enough for several renormalisations (rollback + replay path of acs_cta.cuh)
    tx, sym = frames(code, dc, n_frames, L, EbNo_dB, seed=1234)
    want = ora.decode_frames(sym, n_frames, L)
    assert dec.variants, "no kernel variants"
    for lanes in dec.variants:          # every compiled lanes-per-pair variant must be bit-exact
        dec.set_variant(lanes)
        got = dec.decode_batch(sym, L)
        tag = "T4" if lanes == 104 else f"T{lanes}"      # 104: the K = 7 history kernel with 4 lanes per FRAME (acs_hist<..,T4,..>)
        assert tag in dec.kernel_name, dec.kernel_name
        assert_batch_equal(got, want, f"{name} {decode_type} {EbNo_dB} dB lanes/pair={lanes}")


@pytest.mark.parametrize("decode_type", ["SOFT16", "SOFT8", "HARD8"])
@pytest.mark.parametrize("name", ["Voyager", "LTE", "DAB Radio", "Basic K=5 R=1/2", "CDMA IS-95A", "CDMA 2000", "Cassini"])
def test_streaming_api_matches_oracle_state(cuda_lib, name, decode_type):
    """reset / update in ragged pieces / get_error / chainback + the public m_decisions and m_metrics fields"""
    code = CODE_BY_NAME[name]
    dec, dc = make_cuda_decoder(code, decode_type)
    ora, _ = make_oracle(code, decode_type)
    L = 1024
    tx, sym = frames(code, dc, 1, L, 0.0 if decode_type == "SOFT8" else 2.0, seed=99)      # SOFT8 at 0 dB: many ties, uint8 wrap for long K
    sym = sym[0]
    for d in (dec, ora):
        d.set_traceback_length(L)
        d.reset()
    assert dec.get_traceback_length() == L
    pieces = [1, 5, 6, 7, 64, 300]
    pos, acc_g, acc_o = 0, 0, 0
    steps_total = L + code.K - 1
    k = 0
    while pos < steps_total:
        n = min(pieces[k % len(pieces)], steps_total - pos)
        k += 1
        chunk = sym[pos * code.R:(pos + n) * code.R]
        acc_g += dec.update(chunk)
        acc_o += ora.update(chunk)
        pos += n
        assert dec.m_current_decoded_bit == pos
    assert acc_g == acc_o
    assert (dec.m_metrics == ora.metrics()).all()
    if code.K >= 15 and decode_type == "SOFT16":
        assert acc_o > 0, "test should cross the renormalisation threshold at least once"
    for end_state in (0, 1, (1 << (code.K - 1)) - 1):
        assert dec.get_error(end_state) == ora.get_error(end_state)
        assert (dec.chainback(L, end_state) == ora.chainback(L, end_state)).all()
    assert (dec.m_decisions(0, steps_total) == ora.decisions(steps_total)).all()
    # ragged total_bits: the last byte carries end-state bits below the decoded bits (core.h:96-113)
    for bits in (1, 5, 13, 1021):
        assert (dec.chainback(bits, 3) == ora.chainback(bits, 3)).all()


def test_streaming_api_error_behaviour(cuda_lib):
    """where the reference asserts (scalar.h:37-40, core.h:216-218) the C ABI returns an error code"""
    code = CODE_BY_NAME["Voyager"]
    dec, dc = make_cuda_decoder(code, "SOFT16")
    dec.set_traceback_length(64)
    dec.reset()
    with pytest.raises(v.ViterbiError):
        dec.update(np.zeros(3, dtype=np.int16))                 # not a multiple of R
    with pytest.raises(v.ViterbiError):
        dec.update(np.zeros(2 * (64 + 6 + 1), dtype=np.int16))  # more steps than traceback_length + K-1
    dec.update(np.zeros(2 * 10, dtype=np.int16))
    with pytest.raises(v.ViterbiError):
        dec.chainback(64)                                       # not enough steps decoded yet
    with pytest.raises(v.ViterbiError):
        dec.get_error(64)                                       # end_state out of range
    bt = v.ViterbiBranchTable(9, 2, [0o561, 0o751], 127, -127)
    assert not v.ViterbiDecoder_CUDA.is_valid(bt, dc.decoder_config)   # uncatalogued polynomials with K > 7: no kernel (yet)


@pytest.mark.parametrize("decode_type,expected", [("SOFT16", 100584), ("SOFT8", 2376), ("HARD8", 792)])
def test_dab_fic_punctured_known_answer(cuda_lib, decode_type, expected):
    """examples/run_punctured_decoder.cpp: traceback_error = 792 erased symbols x |high - 0|, 0 bit errors"""
    code = CODE_BY_NAME["DAB Radio"]
    dec, dc = make_cuda_decoder(code, decode_type)
    keep = v.dab_fic_keep_schedule()
    L = 768
    tx, sym = frames(code, dc, 130, L, None, seed=5)
    rx = synth.puncture(sym, keep)
    assert rx.shape[1] == 2304
    dec.set_puncture_schedule(keep, 0)
    out, acc, fin = dec.decode_batch(rx, L)
    assert (out == tx).all()
    assert ((acc + fin) == expected).all()
    # same through the streaming API, one decoded bit per update call like decode_punctured_symbols (puncture_code_helpers.h:32-52)
    dec.set_puncture_schedule(None)
    dec.set_traceback_length(L)
    dec.reset()
    ora, _ = make_oracle(code, decode_type)
    ora.set_traceback_length(L)
    ora.reset()
    depunct = np.zeros(len(keep), dtype=rx.dtype)
    depunct[np.asarray(keep, dtype=bool)] = rx[0]
    total = 0
    for t in range(L + code.K - 1):
        total += dec.update(depunct[t * code.R:(t + 1) * code.R])
    assert total + dec.get_error() == expected
    assert (dec.chainback(L) == tx[0]).all()


@pytest.mark.parametrize("EbNo_dB", [-2.0, 2.0])
def test_dab_fic_punctured_parity_awgn(cuda_lib, EbNo_dB):
    code = CODE_BY_NAME["DAB Radio"]
    dec, dc = make_cuda_decoder(code, "SOFT16")
    ora, _ = make_oracle(code, "SOFT16")
    keep = np.asarray(v.dab_fic_keep_schedule(), dtype=bool)
    L, F = 768, 200
    tx, sym = frames(code, dc, F, L, EbNo_dB, seed=77)
    rx = synth.puncture(sym, keep)
    dec.set_puncture_schedule(keep.astype(np.uint8), 0)
    dep = np.zeros_like(sym)
    dep[:, keep] = rx
    want = ora.decode_frames(dep, F, L)
    for lanes in dec.variants:
        dec.set_variant(lanes)
        assert_batch_equal(dec.decode_batch(rx, L), want, f"FIC {EbNo_dB} dB lanes/pair={lanes}")


@pytest.mark.parametrize("name,decode_type", [("Voyager", "SOFT16"), ("Voyager", "HARD8"), ("DAB Radio", "SOFT16"), ("CDMA IS-95A", "SOFT16")])
def test_simd_tie_break_matches_simd_semantics(cuda_lib, name, decode_type):
    """VITB_TIE_SIMD reproduces the SSE/AVX decision rule (decision = min == path1); checked against the oracle's SIMD mode"""
    code = CODE_BY_NAME[name]
    dec, dc = make_cuda_decoder(code, decode_type, tie_break=v.VITB_TIE_SIMD)
    ora, _ = make_oracle(code, decode_type, mode=MODE_SIMD)
    tx, sym = frames(code, dc, 100, 1024, 0.0, seed=3)
    want = ora.decode_frames(sym, 100, 1024)
    for lanes in dec.variants:
        dec.set_variant(lanes)
        assert_batch_equal(dec.decode_batch(sym, 1024), want, f"{name} {decode_type} simd tie lanes/pair={lanes}")


def test_inconsistent_max_error_config(cuda_lib):
    """soft_decision_max_error != R*(high-low): inverted error is max_error - total, not the complementary table entry"""
    code = CODE_BY_NAME["Voyager"]
    cfg = v.ViterbiDecoder_Config(600, 0, 3000, 60000)
    dec, dc = make_cuda_decoder(code, "SOFT16", config_override=cfg)
    ora, _ = make_oracle(code, "SOFT16", config_override=cfg)
    tx, sym = frames(code, dc, 80, 1024, 1.0, seed=11)
    want = ora.decode_frames(sym, 80, 1024)
    for lanes in dec.variants:
        dec.set_variant(lanes)
        assert_batch_equal(dec.decode_batch(sym, 1024), want, f"cinv lanes/pair={lanes}")
        assert "cinv" in dec.kernel_name


def test_start_and_end_state_options(cuda_lib):
    code = CODE_BY_NAME["Voyager"]
    dec, dc = make_cuda_decoder(code, "SOFT16")
    ora, _ = make_oracle(code, "SOFT16")
    L = 512
    tx, sym = frames(code, dc, 66, L, 3.0, seed=21)
    for lanes in dec.variants:
        dec.set_variant(lanes)
        out, acc, fin = dec.decode_batch(sym, L, starting_state=5, end_state=9)
        ora.set_traceback_length(L)
        for f in range(66):
            ora.reset(5)
            a = ora.update(sym[f])
            assert a == acc[f] and ora.get_error(9) == fin[f]
            assert (ora.chainback(L, 9) == out[f]).all()


def test_workspace_chunking_gives_identical_results(cuda_lib):
    code = CODE_BY_NAME["Voyager"]
    dec, dc = make_cuda_decoder(code, "HARD8")
    tx, sym = frames(code, dc, 500, 512, 2.0, seed=8)
    ref = dec.decode_batch(sym, 512)
    dec.set_workspace_limit(dec.workspace_bytes(64, 512) * 2)      # two 64-frame blocks per chunk
    assert_batch_equal(dec.decode_batch(sym, 512), ref, "chunked")


def test_decode_batch_multi_matches_single(cuda_lib):
    """vitb_decode_batch_multi: contiguous frame ranges, one handle per range (here both handles on device 0; on a multi-GPU box
    each handle is created on its own device), no collective - results identical to one call"""
    code = CODE_BY_NAME["Voyager"]
    dec0, dc = make_cuda_decoder(code, "HARD8")
    dec1, _ = make_cuda_decoder(code, "HARD8")
    tx, sym = frames(code, dc, 333, 512, 3.0, seed=31)
    want = dec0.decode_batch(sym, 512)
    got = v.decode_batch_multi([dec0, dec1], sym, 512)
    assert_batch_equal(got, want, "multi")


def test_pipelined_host_path_matches_device_path(cuda_lib):
    """large enough that the host-pointer call is cut into several copy/compute chunks (and picks a different kernel variant for
    the smaller chunks); must agree with the oracle on a sample and with itself across variants"""
    code = CODE_BY_NAME["Voyager"]
    dec, dc = make_cuda_decoder(code, "HARD8")
    ora, _ = make_oracle(code, "HARD8")
    F, L = 20000, 2048
    tx, sym = frames(code, dc, F, L, 4.0, seed=77)
    got = dec.decode_batch(sym, L)
    dec.set_variant(1)
    ref = dec.decode_batch(sym, L)
    assert_batch_equal(got, ref, "pipelined auto-variant vs T1")
    idx = np.arange(0, F, 97)
    want = ora.decode_frames(sym[idx], idx.size, L)
    assert_batch_equal((got[0][idx], got[1][idx], got[2][idx]), want, "pipelined vs oracle sample")
