"""GPU parity tests for codes OUTSIDE the compiled catalogue: arbitrary generator polynomials go through acs_generic_kernel
(csrc/acs_generic.cuh: branch patterns from a run-time table, branch metrics in shared memory).  Batch and streaming calls, all three
decode types, both tie-break flavours, against the scalar oracle (which takes any polynomials) - bit-exact."""
import numpy as np
import pytest

import viterbidecodercpp_b200 as v
from viterbidecodercpp_b200 import synth
from common import assert_batch_equal
from oracle_binding import OracleDecoder, MODE_SCALAR, MODE_SIMD

pytestmark = pytest.mark.gpu

# (K, R, G): none of these is in examples/helpers/common_codes.h
CODES = [
    (3, 2, [0b101, 0b111]),          # the catalogue has {7, 5}: another order is another code
    (3, 3, [0b111, 0b101, 0b111]),   # every polynomial needs its first and last tap (the butterfly symmetry the reference relies on)
    (4, 2, [0b1111, 0b1101]),
    (5, 3, [0b10011, 0b11101, 0b10111]),
    (6, 2, [0o53, 0o75]),
    (6, 4, [0o53, 0o75, 0o67, 0o71]),
    (7, 2, [0o133, 0o165]),
    (7, 3, [0o133, 0o145, 0o175]),
    (7, 5, [0o133, 0o171, 0o145, 0o165, 0o117]),
    (7, 6, [0o133, 0o171, 0o145, 0o165, 0o117, 0o135]),
]


def make(K, R, G, decode_type, tie=0, mode=MODE_SCALAR, cfg=None):
    dc = v.DECODE_TYPES[decode_type](R)
    c = cfg or dc.decoder_config
    bt = v.ViterbiBranchTable(K, R, G, dc.soft_decision_high, dc.soft_decision_low, dc.soft_bytes)
    dec = v.ViterbiDecoder_CUDA(bt, c, tie_break=tie)
    ora = OracleDecoder(K, R, G, dc.soft_bytes, dc.soft_decision_high, dc.soft_decision_low,
                        [c.soft_decision_max_error, c.initial_start_error, c.initial_non_start_error, c.renormalisation_threshold], mode)
    return dec, ora, dc


@pytest.mark.parametrize("decode_type", ["SOFT16", "SOFT8", "HARD8"])
@pytest.mark.parametrize("K,R,G", CODES)
def test_generic_batch_parity(cuda_lib, K, R, G, decode_type):
    dec, ora, dc = make(K, R, G, decode_type)
    L = 600
    tx, sym = synth.make_frames(K, R, G, 131, L, dc.soft_decision_high, dc.soft_decision_low, dc.soft_bytes, 1.0, 21)
    want = ora.decode_frames(sym, 131, L)
    got = dec.decode_batch(sym, L)
    assert dec.kernel_name.startswith("acs_generic<"), dec.kernel_name
    assert_batch_equal(got, want, f"K={K} R={R} {decode_type}")
    # noise free: the transmitted bytes come back with error 0 (examples/run_tests.cpp:153-191)
    tx, sym = synth.make_frames(K, R, G, 40, 256, dc.soft_decision_high, dc.soft_decision_low, dc.soft_bytes, None, 3)
    out, acc, fin = dec.decode_batch(sym, 256)
    assert (out == tx).all() and ((acc + fin) == 0).all()


@pytest.mark.parametrize("decode_type", ["SOFT16", "HARD8"])
@pytest.mark.parametrize("K,R,G", [CODES[1], CODES[5], CODES[8]])
def test_generic_random_symbols_and_ragged_lengths(cuda_lib, K, R, G, decode_type):
    """random symbols (constant ties), ragged bit counts, non-zero start / end states"""
    dec, ora, dc = make(K, R, G, decode_type)
    rng = np.random.default_rng(8)
    dt = np.int8 if dc.soft_bytes == 1 else np.int16
    ns = 1 << (K - 1)
    for L, start, end in [(1, 0, 0), (13, 1, ns - 1), (64, 0, 0), (333, ns - 1, 2 % ns)]:
        sym = rng.integers(dc.soft_decision_low, dc.soft_decision_high + 1, size=(70, (L + K - 1) * R)).astype(dt)
        out = np.zeros((70, (L + 7) // 8), dtype=np.uint8); acc = np.zeros(70, dtype=np.uint64); fin = np.zeros(70, dtype=np.uint32)
        ora.set_traceback_length(L)
        for f in range(70):
            ora.reset(start); acc[f] = ora.update(sym[f]); fin[f] = ora.get_error(end); out[f] = ora.chainback(L, end)
        got = dec.decode_batch(sym, L, starting_state=start, end_state=end)
        assert_batch_equal(got, (out, acc, fin), f"K={K} R={R} {decode_type} L={L}")


@pytest.mark.parametrize("K,R,G", [CODES[3], CODES[6]])
def test_generic_simd_tie_and_inconsistent_config(cuda_lib, K, R, G):
    dc = v.DECODE_TYPES["SOFT16"](R)
    c = dc.decoder_config
    cfg = v.ViterbiDecoder_Config(c.soft_decision_max_error + 5, c.initial_start_error, c.initial_non_start_error, c.renormalisation_threshold)
    for tie, mode, conf in [(v.VITB_TIE_SIMD, MODE_SIMD, None), (0, MODE_SCALAR, cfg)]:
        dec, ora, _ = make(K, R, G, "SOFT16", tie=tie, mode=mode, cfg=conf)
        tx, sym = synth.make_frames(K, R, G, 90, 512, dc.soft_decision_high, dc.soft_decision_low, dc.soft_bytes, 0.0, 5)
        assert_batch_equal(dec.decode_batch(sym, 512), ora.decode_frames(sym, 90, 512), f"K={K} R={R} tie={tie} cfg={conf is not None}")


@pytest.mark.parametrize("decode_type", ["SOFT16", "HARD8"])
@pytest.mark.parametrize("K,R,G", [CODES[0], CODES[4], CODES[9]])
def test_generic_streaming_api(cuda_lib, K, R, G, decode_type):
    """reset / update in ragged pieces / get_error / chainback, plus the decision rows and metrics the reference exposes"""
    dec, ora, dc = make(K, R, G, decode_type)
    L = 504
    tx, sym = synth.make_frames(K, R, G, 1, L, dc.soft_decision_high, dc.soft_decision_low, dc.soft_bytes, 2.0, 17)
    sym = sym[0]
    for d in (dec, ora):
        d.set_traceback_length(L)
        d.reset()
    pos, steps_total, k, acc_g, acc_o = 0, L + K - 1, 0, 0, 0
    pieces = [1, 3, 8, 50, 200]
    while pos < steps_total:
        n = min(pieces[k % len(pieces)], steps_total - pos)
        k += 1
        chunk = sym[pos * R:(pos + n) * R]
        acc_g += dec.update(chunk)
        acc_o += ora.update(chunk)
        pos += n
    assert acc_g == acc_o
    assert dec.get_error() == ora.get_error()
    assert (dec.chainback(L) == ora.chainback(L)).all()
    assert (dec.m_metrics == ora.metrics()).all()
    assert (dec.m_decisions(0, steps_total) == ora.decisions(steps_total)).all()
