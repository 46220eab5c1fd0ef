"""VITB_TIE_SIMD_SAT: the reference's SSE / AVX / NEON decoders reproduced EXACTLY - saturating metric adds, inverted error as
max_error -sat total, saturating sum of the branch errors, ties (saturated candidates included) to path 1
(x86/viterbi_decoder_avx_u16.h:95-115, avx_u8.h alike) - in the decision-row kernels (csrc/acs_pair.cuh "saturating flavour").
Checked against the oracle's SIMD mode (itself pinned to the reference's AVX2 / SSE decoders on the CPU) AND against the
reference's AVX2 decoder directly, on inputs that really saturate: uint8_t metrics of the long codes (Cassini SOFT8 is the case
SURVEY.md section 8 f-1 names: the scalar decoder wraps there and the reference skips it, run_tests.cpp:63-65), maximum-error
symbols, inconsistent max_error."""
import numpy as np
import pytest

import viterbidecodercpp_b200 as v
import oracle_binding as ob
from common import CODE_BY_NAME, GPU_CODES, assert_batch_equal, frames, make_cuda_decoder, make_oracle, oracle_batch
from oracle_binding import MODE_SIMD
from test_gpu_tag_stress import adversarial_batches

pytestmark = pytest.mark.gpu
SAT = 2      # VITB_TIE_SIMD_SAT


def decode_oracle(ora, code, sym, n, L):
    ora.max_metric_seen_all(clear=True)
    want = oracle_batch(ora, code, sym, L) if code.K >= 15 else ora.decode_frames(sym, n, L)
    return want, ora.max_metric_seen_all()


@pytest.mark.parametrize("decode_type", ["SOFT16", "SOFT8", "HARD8"])
@pytest.mark.parametrize("name", GPU_CODES)
def test_saturating_flavour_equals_the_simd_oracle(cuda_lib, name, decode_type):
    """every catalogue code x decode type x compiled variant, noisy frames at a low SNR"""
    code = CODE_BY_NAME[name]
    dec, dc = make_cuda_decoder(code, decode_type, tie_break=SAT)
    ora, _ = make_oracle(code, decode_type, mode=MODE_SIMD)
    n, L = (4, 512) if code.K >= 15 else (67, 600)
    tx, sym = frames(code, dc, n, L, 0.0, seed=11)
    want, top = decode_oracle(ora, code, sym, n, L)
    for lanes in dec.variants:
        dec.set_variant(lanes)
        got = dec.decode_batch(sym, L)
        assert "simd-sat" in dec.kernel_name, dec.kernel_name
        assert_batch_equal(got, want, f"{name} {decode_type} variant {lanes} ({dec.kernel_name})")
    dec.close()


@pytest.mark.parametrize("name,decode_type", [("Voyager", "HARD8"), ("Voyager", "SOFT8"), ("DAB Radio", "SOFT8"), ("LTE", "HARD8"),
                                               ("CDMA IS-95A", "SOFT8"), ("CDMA 2000", "HARD8"), ("Voyager", "SOFT16"), ("Basic K=5 R=1/2", "SOFT8")])
def test_saturating_flavour_on_saturating_inputs(cuda_lib, name, decode_type):
    """a configuration that MUST saturate: the non-start states begin 5 below the top of the metric type and the threshold is the
    top itself, so the first steps pile saturated candidates on each other (ties between saturated values select path 1), and the
    metrics of the whole trellis live at the top of the range until state 0 itself saturates and triggers the renormalisation"""
    code = CODE_BY_NAME[name]
    dc = v.DECODE_TYPES[decode_type](code.R)
    c = dc.decoder_config
    top = 255 if dc.soft_bytes == 1 else 65535
    cfg = v.ViterbiDecoder_Config(c.soft_decision_max_error, c.initial_start_error, top - 5, top)
    dec, _ = make_cuda_decoder(code, decode_type, tie_break=SAT, config_override=cfg)
    ora, _ = make_oracle(code, decode_type, mode=MODE_SIMD, config_override=cfg)
    n, L = 66, 180
    tops = []
    for what, sym in adversarial_batches(dc, n, (L + code.K - 1) * code.R, seed=5).items():
        want, seen = decode_oracle(ora, code, sym, n, L)
        tops.append(seen)
        for lanes in dec.variants:
            dec.set_variant(lanes)
            got = dec.decode_batch(sym, L)
            assert_batch_equal(got, want, f"{name} {decode_type} {what} variant {lanes}")
    assert max(tops) == top, f"{name} {decode_type}: nothing saturated (max {max(tops)})"
    dec.close()


@pytest.mark.parametrize("name,decode_type", [("Voyager", "SOFT8"), ("DAB Radio", "HARD8"), ("CDMA IS-95A", "SOFT16"), ("Cassini", "SOFT8")])
def test_saturating_flavour_with_inconsistent_max_error(cuda_lib, name, decode_type):
    """max_error smaller than a possible total error: the inverted error saturates at 0 (avx_u16.h:106 subs_epu16)"""
    code = CODE_BY_NAME[name]
    dc = v.DECODE_TYPES[decode_type](code.R)
    c = dc.decoder_config
    cfg = v.ViterbiDecoder_Config(max(1, c.soft_decision_max_error // 2), c.initial_start_error, c.initial_non_start_error, c.renormalisation_threshold)
    dec, _ = make_cuda_decoder(code, decode_type, tie_break=SAT, config_override=cfg)
    ora, _ = make_oracle(code, decode_type, mode=MODE_SIMD, config_override=cfg)
    n, L = (3, 256) if code.K >= 15 else (40, 304)
    tx, sym = frames(code, dc, n, L, 1.0, seed=3)
    want, _ = decode_oracle(ora, code, sym, n, L)
    got = dec.decode_batch(sym, L)
    assert_batch_equal(got, want, f"{name} {decode_type} max_error / 2")
    dec.close()


@pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref/libvitref.so not built")
@pytest.mark.parametrize("name,decode_type", [("Cassini", "SOFT8"), ("Cassini", "HARD8"), ("CDMA 2000", "SOFT8"), ("Voyager", "HARD8"),
                                               ("DAB Radio", "SOFT16")])
def test_saturating_flavour_equals_the_reference_avx_decoder(cuda_lib, name, decode_type):
    """the reference's ViterbiDecoder_AVX_u16 / _u8 itself (unmodified headers), including where its metrics saturate"""
    code = CODE_BY_NAME[name]
    dec, dc = make_cuda_decoder(code, decode_type, tie_break=SAT)
    c = dc.decoder_config
    n, L = (4, 512) if code.K >= 15 else (67, 1000)
    tx, sym = frames(code, dc, n, L, 0.0, seed=12)
    r = ob.ref_decode(code.K, code.R, code.G, dc.soft_bytes, dc.soft_decision_high, dc.soft_decision_low,
                      [c.soft_decision_max_error, c.initial_start_error, c.initial_non_start_error, c.renormalisation_threshold],
                      ob.IMPL_AVX, sym, n, L)
    got = dec.decode_batch(sym, L)
    assert_batch_equal(got, (r["bytes"], r["acc"], r["final"]), f"{name} {decode_type} vs reference AVX2")
    dec.close()


def test_saturating_flavour_streaming_rows(cuda_lib):
    """streaming calls in the saturating flavour: metrics and every decision row against the oracle's SIMD mode"""
    for name, decode_type in (("Voyager", "SOFT8"), ("CDMA IS-95A", "HARD8"), ("Cassini", "SOFT8")):
        code = CODE_BY_NAME[name]
        dec, dc = make_cuda_decoder(code, decode_type, tie_break=SAT)
        ora, _ = make_oracle(code, decode_type, mode=MODE_SIMD)
        L = 200
        tx, sym = frames(code, dc, 1, L, -1.0, seed=8)
        f = sym[0]
        dec.set_traceback_length(L); dec.reset()
        ora.set_traceback_length(L); ora.reset()
        pos, ag, ao = 0, 0, 0
        for piece in (3, 7, 64, 10 ** 6):
            n = min(piece * code.R, f.size - pos)
            if n <= 0:
                break
            ag += dec.update(f[pos:pos + n]); ao += ora.update(f[pos:pos + n]); pos += n
        rows = L + code.K - 1
        assert ag == ao and (dec.m_metrics == ora.metrics()).all(), name
        assert (dec.m_decisions(0, rows) == ora.decisions(rows)).all(), name
        assert (dec.chainback(L) == ora.chainback(L)).all() and dec.get_error() == ora.get_error()
        dec.close()
