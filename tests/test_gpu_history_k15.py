"""GPU parity tests for the K = 15 survivor-history kernel (csrc/acs_hist_cta.cuh: one frame per 512-thread CTA, 16-step records in
position order, exchange every 5 steps, renormalisation by speculation / rollback / replay) and the record walk that reads it
(traceback_hist_kernel with logt = 9): every residue of the step count modulo 16 (record) x modulo 5 (exchange period), random
symbols (constant metric ties), thresholds that renormalise on every step / early / at the stock rate, start / end / best end states,
both tie-break flavours, an inconsistent max_error, unaligned and strided rows, forced traceback segments, agreement with the
decision-row kernels.  Everything through the C ABI against the scalar oracle, bit-exact."""
import ctypes as C

import numpy as np
import pytest

import viterbidecodercpp_b200 as v
from common import CODE_BY_NAME, assert_batch_equal, frames, make_cuda_decoder, make_oracle, oracle_batch, random_symbols
from oracle_binding import MODE_SIMD

pytestmark = pytest.mark.gpu

CODE = CODE_BY_NAME["Cassini"]


def hist_cta_decoder(**kw):
    dec, dc = make_cuda_decoder(CODE, "SOFT16", **kw)
    dec.set_variant(256)
    return dec, dc


def check_kernel(dec):
    assert dec.kernel_name.startswith("acs_hist<K15") and ",T256," in dec.kernel_name, dec.kernel_name


def test_k15_history_is_the_default_for_soft16_batches(cuda_lib):
    dec, dc = make_cuda_decoder(CODE, "SOFT16")
    assert 256 in dec.variants
    sym = random_symbols(dc, 3, (64 + CODE.K - 1) * CODE.R, seed=1)
    dec.decode_batch(sym, 64)
    check_kernel(dec)


def test_k15_history_every_record_and_exchange_residue(cuda_lib):
    """total_bits 1 .. 82: S = L + 14 steps covers every residue modulo 16 (record) x modulo 5 (exchange period) = lcm 80, ragged last
    bytes included"""
    dec, dc = hist_cta_decoder()
    ora, _ = make_oracle(CODE, "SOFT16")
    for L in range(1, 83):
        sym = random_symbols(dc, 3, (L + CODE.K - 1) * CODE.R, seed=7000 + L)
        want = ora.decode_frames(sym, 3, L)
        got = dec.decode_batch(sym, L)
        check_kernel(dec)
        assert_batch_equal(got, want, f"Cassini L={L}")


@pytest.mark.parametrize("thr_kind", ["zero", "one", "non_start", "low"])
def test_k15_history_renormalisation_rollback(cuda_lib, thr_kind):
    """thresholds that make the trigger fire on every step (every group is rolled back and replayed), on the first steps, or every few
    groups (random symbols grow the metrics by ~R * 64 per step: threshold 20000 renormalises about every 50 steps)"""
    dc = v.DECODE_TYPES["SOFT16"](CODE.R)
    c = dc.decoder_config
    thr = {"zero": 0, "one": 1, "non_start": c.initial_non_start_error, "low": 20000}[thr_kind]
    cfg = v.ViterbiDecoder_Config(c.soft_decision_max_error, c.initial_start_error, c.initial_non_start_error, thr)
    dec, _ = hist_cta_decoder(config_override=cfg)
    ora, _ = make_oracle(CODE, "SOFT16", config_override=cfg)
    for L in (5, 47, 333):
        sym = random_symbols(dc, 3, (L + CODE.K - 1) * CODE.R, seed=L)
        want = ora.decode_frames(sym, 3, L)
        if thr_kind == "low" and L == 333:
            assert int(want[1].min()) > 0
        got = dec.decode_batch(sym, L)
        check_kernel(dec)
        assert_batch_equal(got, want, f"Cassini thr={thr} L={L}")


@pytest.mark.parametrize("ebno", [-2.0, 3.0])
def test_k15_history_noisy_frames(cuda_lib, ebno):
    dec, dc = hist_cta_decoder()
    ora, _ = make_oracle(CODE, "SOFT16")
    L = 3000
    tx, sym = frames(CODE, dc, 5, L, ebno, seed=int(ebno) + 50)
    want = ora.decode_frames(sym, 5, L)
    assert int(want[1].min()) > 0, "every frame must renormalise"
    got = dec.decode_batch(sym, L)
    check_kernel(dec)
    assert_batch_equal(got, want, f"Cassini Eb/N0={ebno}")


def test_k15_history_start_end_and_best_states(cuda_lib):
    dec, dc = hist_cta_decoder()
    ora, _ = make_oracle(CODE, "SOFT16")
    ns = 1 << (CODE.K - 1)
    for L, start, end in [(100, 5, 0), (61, 0, ns - 1), (203, ns - 1, 12345), (48, 9000, 77)]:
        sym = random_symbols(dc, 3, (L + CODE.K - 1) * CODE.R, seed=L + start + end)
        want = oracle_batch(ora, CODE, sym, L, start, end)
        got = dec.decode_batch(sym, L, starting_state=start, end_state=end)
        check_kernel(dec)
        assert_batch_equal(got, want, f"Cassini L={L} start={start} end={end}")
    L, n = 520, 4
    sym = random_symbols(dc, n, (L + CODE.K - 1) * CODE.R, seed=L)
    out = np.zeros((n, (L + 7) // 8), dtype=np.uint8); acc = np.zeros(n, dtype=np.uint64); fin = np.zeros(n, dtype=np.uint32)
    ora.set_traceback_length(L)
    bests = set()
    for f in range(n):
        ora.reset(0)
        acc[f] = ora.update(sym[f])
        best = int(np.argmin(np.asarray(ora.metrics())))
        bests.add(best)
        fin[f] = ora.get_error(best)
        out[f] = ora.chainback(L, best)
    assert len(bests) > 1
    got = dec.decode_batch(sym, L, end_state=v.VITB_END_STATE_BEST)
    check_kernel(dec)
    assert_batch_equal(got, (out, acc, fin), "Cassini best end state")


def test_k15_history_simd_tie_break_and_inconsistent_max_error(cuda_lib):
    dc = v.DECODE_TYPES["SOFT16"](CODE.R)
    c = dc.decoder_config
    cfg2 = v.ViterbiDecoder_Config(c.soft_decision_max_error + 3, c.initial_start_error, c.initial_non_start_error, c.renormalisation_threshold)
    for tie, mode, cfg in ((v.VITB_TIE_SIMD, MODE_SIMD, None), (v.VITB_TIE_SCALAR, 0, cfg2), (v.VITB_TIE_SIMD, MODE_SIMD, cfg2)):
        dec, _ = hist_cta_decoder(tie_break=tie, config_override=cfg)
        ora, _ = make_oracle(CODE, "SOFT16", mode=mode, config_override=cfg)
        for L in (13, 200):
            sym = random_symbols(dc, 3, (L + CODE.K - 1) * CODE.R, seed=77 + L)
            want = ora.decode_frames(sym, 3, L)
            got = dec.decode_batch(sym, L)
            check_kernel(dec)
            assert_batch_equal(got, want, f"Cassini tie={tie} cfg={'inconsistent' if cfg else 'stock'} L={L}")


def test_k15_history_strided_rows(cuda_lib):
    """row stride larger than the symbols used (odd number of int16_t: rows only 2-byte aligned)"""
    dec, dc = hist_cta_decoder()
    ora, _ = make_oracle(CODE, "SOFT16")
    L = 150
    n_sym = (L + CODE.K - 1) * CODE.R
    full = random_symbols(dc, 4, n_sym, seed=11, pad=3)
    want = ora.decode_frames(np.ascontiguousarray(full[:, :n_sym]), 4, L)
    out = np.zeros((4, (L + 7) // 8), dtype=np.uint8); acc = np.zeros(4, dtype=np.uint64); fin = np.zeros(4, dtype=np.uint32)
    o = dec._opts(full.shape[1], 0, 0)
    rc = dec._L.vitb_decode_batch(dec._h, full.ctypes.data, 4, L, C.byref(o), out.ctypes.data, acc.ctypes.data, fin.ctypes.data)
    assert rc == 0
    check_kernel(dec)
    assert_batch_equal((out, acc, fin), want, "Cassini strided rows")


def test_k15_history_segmented_traceback_and_decision_row_kernels_agree(cuda_lib):
    dec, dc = hist_cta_decoder()
    ora, _ = make_oracle(CODE, "SOFT16")
    L, end = 1003, 5
    sym = random_symbols(dc, 4, (L + CODE.K - 1) * CODE.R, seed=L)
    want = oracle_batch(ora, CODE, sym, L, 0, end)
    for seg, ov in [(0, -1), (7, 0), (5, 1), (16, 3), (1, 0), (1000, 0)]:
        dec.set_traceback_segments(seg, ov)
        got = dec.decode_batch(sym, L, end_state=end)
        check_kernel(dec)
        assert_batch_equal(got, want, f"Cassini seg={seg} overlap={ov}")
    dec.set_traceback_segments(0, -1)
    for lanes in (512, 1024):
        dec.set_variant(lanes)
        got = dec.decode_batch(sym, L, end_state=end)
        assert dec.kernel_name.startswith("acs_cta<"), dec.kernel_name
        assert_batch_equal(got, want, f"Cassini decision rows T{lanes}")
