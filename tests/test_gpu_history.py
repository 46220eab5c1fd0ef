"""GPU parity tests aimed at the survivor-history kernel (csrc/acs_hist.cuh) and its traceback (traceback_hist_kernel):
random soft symbols inside [low, high] (arbitrary survivor paths, many metric ties - the tie-break lives in the history tags),
frame lengths around every record boundary (8 steps for uint8_t metrics, 16 for uint16_t), ragged bit counts, non-zero start and
end states, both tie-break flavours, an inconsistent max_error, packed-stream input (unaligned rows), and agreement with the
decision-row kernels.  Everything against the scalar oracle through the C ABI, bit-exact."""
import numpy as np
import pytest

import viterbidecodercpp_b200 as v
from common import CODE_BY_NAME, PAIR_CODES, assert_batch_equal, make_cuda_decoder, make_oracle, oracle_batch, random_symbols
from oracle_binding import MODE_SCALAR, MODE_SIMD

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("decode_type", ["SOFT16", "SOFT8", "HARD8"])
@pytest.mark.parametrize("name", PAIR_CODES)
def test_history_kernel_every_record_residue(cuda_lib, name, decode_type):
    """total_bits from 1 to 40 (every residue of the step count modulo the 8 / 16-step record, ragged last bytes) and a long frame"""
    code = CODE_BY_NAME[name]
    dec, dc = make_cuda_decoder(code, decode_type)
    ora, _ = make_oracle(code, decode_type)
    dec.set_variant(1)
    for L in list(range(1, 41)) + [257, 1003]:
        n_sym = (L + code.K - 1) * code.R
        n_frames = 70 if L < 100 else 67
        sym = random_symbols(dc, n_frames, n_sym, seed=1000 + L)
        want = ora.decode_frames(sym, n_frames, L)
        got = dec.decode_batch(sym, L)
        assert dec.kernel_name.startswith("acs_hist<"), dec.kernel_name
        assert_batch_equal(got, want, f"{name} {decode_type} L={L}")


@pytest.mark.parametrize("decode_type", ["SOFT16", "SOFT8", "HARD8"])
@pytest.mark.parametrize("name", ["Voyager", "LTE", "DAB Radio", "Basic K=3 R=1/2"])
def test_history_kernel_renormalises_every_step(cuda_lib, name, decode_type):
    """renormalisation_threshold 0 / 1 / the non-start error: the trigger fires after every step or on the very first steps, and
    two frames of a register trigger independently (the uint16_t format defers the subtraction into the next step's branch metric
    table, the uint8_t format subtracts in place); also a renormalisation pending after the LAST step must reach the final metrics"""
    code = CODE_BY_NAME[name]
    dc = v.DECODE_TYPES[decode_type](code.R)
    c = dc.decoder_config
    for thr in (0, 1, c.initial_non_start_error):
        cfg = v.ViterbiDecoder_Config(c.soft_decision_max_error, c.initial_start_error, c.initial_non_start_error, thr)
        dec, _ = make_cuda_decoder(code, decode_type, config_override=cfg)
        ora, _ = make_oracle(code, decode_type, config_override=cfg)
        dec.set_variant(1)
        for L in (3, 16, 47, 300):
            sym = random_symbols(dc, 70, (L + code.K - 1) * code.R, seed=thr * 7 + L)
            want = ora.decode_frames(sym, 70, L)
            got = dec.decode_batch(sym, L)
            assert dec.kernel_name.startswith("acs_hist<"), dec.kernel_name
            assert_batch_equal(got, want, f"{name} {decode_type} thr={thr} L={L}")


@pytest.mark.parametrize("decode_type", ["SOFT16", "HARD8"])
@pytest.mark.parametrize("name", ["Voyager", "DAB Radio", "Basic K=5 R=1/2"])
def test_history_kernel_matches_decision_row_kernel(cuda_lib, name, decode_type):
    code = CODE_BY_NAME[name]
    dec, dc = make_cuda_decoder(code, decode_type)
    dec.set_variant(1)
    L = 777
    sym = random_symbols(dc, 130, (L + code.K - 1) * code.R, seed=5)
    a = dec.decode_batch(sym, L)
    assert dec.kernel_name.startswith("acs_hist<")
    dec.set_history_kernel(False)
    b = dec.decode_batch(sym, L)
    assert dec.kernel_name.startswith("acs<"), dec.kernel_name
    dec.set_history_kernel(True)
    assert_batch_equal(a, b, f"{name} {decode_type} history vs decision rows")


@pytest.mark.parametrize("decode_type", ["SOFT16", "SOFT8", "HARD8"])
@pytest.mark.parametrize("name", ["Voyager", "LTE", "DAB Radio"])
def test_history_kernel_start_and_end_states(cuda_lib, name, decode_type):
    code = CODE_BY_NAME[name]
    dec, dc = make_cuda_decoder(code, decode_type)
    ora, _ = make_oracle(code, decode_type)
    dec.set_variant(1)
    ns = 1 << (code.K - 1)
    for L, start, end in [(100, 5, 0), (100, 0, ns - 1), (61, 9, 37 % ns), (203, ns - 1, 21 % ns)]:
        sym = random_symbols(dc, 33, (L + code.K - 1) * code.R, seed=L + start + end)
        want = oracle_batch(ora, code, sym, L, start, end)
        got = dec.decode_batch(sym, L, starting_state=start, end_state=end)
        assert dec.kernel_name.startswith("acs_hist<")
        assert_batch_equal(got, want, f"{name} {decode_type} L={L} start={start} end={end}")


@pytest.mark.parametrize("decode_type", ["SOFT16", "HARD8"])
@pytest.mark.parametrize("name", ["Voyager", "LTE", "DAB Radio", "Basic K=3 R=1/2"])
def test_history_kernel_simd_tie_break(cuda_lib, name, decode_type):
    """VITB_TIE_SIMD: decision = (min == path 1), x86/viterbi_decoder_avx_u16.h:112-115; random symbols tie all the time"""
    code = CODE_BY_NAME[name]
    dc = v.DECODE_TYPES[decode_type](code.R)
    c = dc.decoder_config
    # the SIMD flavour saturates where the CUDA path wraps (only its tie-break is reproduced): renormalise early enough that no
    # uint8_t metric gets near 255 with random symbols
    thr = c.renormalisation_threshold if decode_type == "SOFT16" else 100
    cfg = v.ViterbiDecoder_Config(c.soft_decision_max_error, c.initial_start_error, c.initial_non_start_error, thr)
    dec, _ = make_cuda_decoder(code, decode_type, tie_break=v.VITB_TIE_SIMD, config_override=cfg)
    ora, _ = make_oracle(code, decode_type, mode=MODE_SIMD, config_override=cfg)
    dec.set_variant(1)
    for L in (13, 64, 500):
        sym = random_symbols(dc, 70, (L + code.K - 1) * code.R, seed=77 + L)
        want = ora.decode_frames(sym, 70, L)
        got = dec.decode_batch(sym, L)
        assert dec.kernel_name.startswith("acs_hist<")
        assert_batch_equal(got, want, f"{name} {decode_type} simd tie L={L}")


@pytest.mark.parametrize("decode_type", ["SOFT16", "HARD8"])
def test_history_kernel_inconsistent_max_error(cuda_lib, decode_type):
    """soft_decision_max_error != R * (high - low): the inverted error is no longer the complementary table entry (scalar.h:107)"""
    code = CODE_BY_NAME["Voyager"]
    dc = v.DECODE_TYPES[decode_type](code.R)
    c = dc.decoder_config
    cfg = v.ViterbiDecoder_Config(c.soft_decision_max_error + 3, c.initial_start_error, c.initial_non_start_error, c.renormalisation_threshold)
    dec, _ = make_cuda_decoder(code, decode_type, config_override=cfg)
    ora, _ = make_oracle(code, decode_type, config_override=cfg)
    dec.set_variant(1)
    L = 300
    sym = random_symbols(dc, 70, (L + code.K - 1) * code.R, seed=3)
    want = ora.decode_frames(sym, 70, L)
    got = dec.decode_batch(sym, L)
    assert dec.kernel_name.startswith("acs_hist<")
    assert_batch_equal(got, want, f"Voyager {decode_type} inconsistent max_error")


@pytest.mark.parametrize("decode_type", ["SOFT16", "HARD8"])
@pytest.mark.parametrize("name", ["Voyager", "LTE", "DAB Radio"])
def test_history_kernel_packed_stream_input(cuda_lib, name, decode_type):
    """rows whose stride is not a multiple of 4 bytes cannot be fetched directly: ingest -> packed stream -> history kernel"""
    code = CODE_BY_NAME[name]
    dec, dc = make_cuda_decoder(code, decode_type)
    ora, _ = make_oracle(code, decode_type)
    dec.set_variant(1)
    L = 333
    n_sym = (L + code.K - 1) * code.R
    pad = 1 if (n_sym * dc.soft_bytes) % 4 != 3 else 2
    if ((n_sym + pad) * dc.soft_bytes) % 4 == 0:
        pad += 1
    full = random_symbols(dc, 70, n_sym, seed=11, pad=pad)
    want = ora.decode_frames(np.ascontiguousarray(full[:, :n_sym]), 70, L)
    # decode_batch takes the row stride from the array's second dimension: the padded array gives unaligned rows
    import ctypes as C
    out = np.zeros((70, (L + 7) // 8), dtype=np.uint8)
    acc = np.zeros(70, dtype=np.uint64)
    fin = np.zeros(70, dtype=np.uint32)
    o = dec._opts(full.shape[1], 0, 0)
    rc = dec._L.vitb_decode_batch(dec._h, full.ctypes.data, 70, L, C.byref(o), out.ctypes.data, acc.ctypes.data, fin.ctypes.data)
    assert rc == 0
    assert dec.kernel_name.startswith("acs_hist<")
    assert_batch_equal((out, acc, fin), want, f"{name} {decode_type} packed stream")


@pytest.mark.parametrize("decode_type", ["SOFT16", "HARD8"])
@pytest.mark.parametrize("name", ["Voyager", "DAB Radio", "Basic K=5 R=1/2"])
def test_segmented_traceback_and_its_repair_path(cuda_lib, name, decode_type):
    """the history-record traceback walks a frame in concurrent segments after a warm-up from a guessed state; with the warm-up
    switched off (overlap 0) nearly every segment starts from the wrong state and must be repaired: results stay bit-exact"""
    code = CODE_BY_NAME[name]
    dec, dc = make_cuda_decoder(code, decode_type)
    ora, _ = make_oracle(code, decode_type)
    dec.set_variant(1)
    for L, end in [(1000, 0), (2049, 3), (517, 1)]:
        n_sym = (L + code.K - 1) * code.R
        sym = random_symbols(dc, 40, n_sym, seed=L)
        want = oracle_batch(ora, code, sym, L, 0, end)
        for seg, ov in [(0, -1), (7, 0), (5, 1), (16, 3), (1, 0), (1000, 0)]:
            dec.set_traceback_segments(seg, ov)
            got = dec.decode_batch(sym, L, end_state=end)
            assert dec.kernel_name.startswith("acs_hist<")
            assert_batch_equal(got, want, f"{name} {decode_type} L={L} seg={seg} overlap={ov}")
    dec.set_traceback_segments(0, -1)


@pytest.mark.parametrize("decode_type", ["SOFT16", "HARD8"])
def test_segmented_traceback_k15(cuda_lib, decode_type):
    """K = 15 decision rows: the same segmented walk (segments in units of 8 decoded bits, warm-up in units of 8 rows), automatic
    setting and settings that force the repair path; ragged bit count and a non-zero end state"""
    from common import frames
    code = CODE_BY_NAME["Cassini"]
    dec, dc = make_cuda_decoder(code, decode_type)
    ora, _ = make_oracle(code, decode_type)
    for L, end in [(2048, 0), (1003, 5)]:
        n_sym = (L + code.K - 1) * code.R
        if L % 8 == 0:
            tx, sym = frames(code, dc, 6, L, 3.0, seed=L)
        else:
            sym = random_symbols(dc, 6, n_sym, seed=L)
        want = oracle_batch(ora, code, sym, L, 0, end)
        for seg, ov in [(0, -1), (8, 0), (16, 4), (3, 1)]:
            dec.set_traceback_segments(seg, ov)
            got = dec.decode_batch(sym, L, end_state=end)
            assert_batch_equal(got, want, f"Cassini {decode_type} L={L} seg={seg} overlap={ov}")
    dec.set_traceback_segments(0, -1)


@pytest.mark.parametrize("name,decode_type,lanes", [("Voyager", "SOFT16", 1), ("Voyager", "HARD8", 1), ("DAB Radio", "SOFT16", 4),
                                                     ("Voyager", "HARD8", 2), ("CDMA IS-95A", "SOFT16", 16), ("CDMA IS-95A", "HARD8", 8),
                                                     ("Cassini", "SOFT16", 0)])
def test_best_end_state(cuda_lib, name, decode_type, lanes):
    """end_state = VITB_END_STATE_BEST: every frame is traced back from its own best final state (smallest metric, lowest index on a
    tie) - what a caller of the reference computes with get_error(s) over all s before chainback(..., s).  Random symbols, so the best
    state differs from frame to frame; covers the history, lane-group and K = 15 traceback kernels, ragged bit counts included."""
    code = CODE_BY_NAME[name]
    dec, dc = make_cuda_decoder(code, decode_type)
    ora, _ = make_oracle(code, decode_type)
    if lanes:
        dec.set_variant(lanes)
    n_frames = 6 if code.K >= 15 else 37
    for L in ([520, 1003] if code.K < 15 else [1003]):
        sym = random_symbols(dc, n_frames, (L + code.K - 1) * code.R, seed=L + lanes)
        out = np.zeros((n_frames, (L + 7) // 8), dtype=np.uint8); acc = np.zeros(n_frames, dtype=np.uint64); fin = np.zeros(n_frames, dtype=np.uint32)
        ora.set_traceback_length(L)
        bests = []
        for f in range(n_frames):
            ora.reset(0)
            acc[f] = ora.update(sym[f])
            best = int(np.argmin(np.asarray(ora.metrics())))
            bests.append(best)
            fin[f] = ora.get_error(best)
            out[f] = ora.chainback(L, best)
        assert len(set(bests)) > 1
        got = dec.decode_batch(sym, L, end_state=v.VITB_END_STATE_BEST)
        assert_batch_equal(got, (out, acc, fin), f"{name} {decode_type} lanes={lanes} L={L} best end state")


@pytest.mark.parametrize("name,decode_type", [("Voyager", "SOFT16"), ("CDMA IS-95A", "HARD8"), ("Cassini", "SOFT16")])
def test_best_end_state_streaming_calls(cuda_lib, name, decode_type):
    """get_error / chainback of the streaming API with VITB_END_STATE_BEST (pair, lane-group and K = 15 decision rows)"""
    code = CODE_BY_NAME[name]
    dec, dc = make_cuda_decoder(code, decode_type)
    ora, _ = make_oracle(code, decode_type)
    L = 203
    sym = random_symbols(dc, 1, (L + code.K - 1) * code.R, seed=12)[0]
    for d in (dec, ora):
        d.set_traceback_length(L)
        d.reset()
    assert dec.update(sym) == ora.update(sym)
    best = int(np.argmin(np.asarray(ora.metrics())))
    assert best != 0
    assert dec.get_error(v.VITB_END_STATE_BEST) == ora.get_error(best)
    assert (dec.chainback(L, v.VITB_END_STATE_BEST) == ora.chainback(L, best)).all()
    assert (dec.chainback(L, 0) == ora.chainback(L, 0)).all()
