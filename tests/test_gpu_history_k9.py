"""GPU parity tests for the frame-over-4-lanes survivor-history kernel (csrc/acs_hist_group.cuh: 16-step records in POSITION
order, exchange every 6 steps for K = 9 - variant 4 - and every 4 steps for K = 7 - variant 104, the small-batch kernel -, lazy
renormalisation) and the record walk that reads it (traceback_hist_kernel with logt = 2).  Same net as tests/test_gpu_history.py throws over the K <= 7 kernels: every residue of the step count modulo the
16-step record AND the 6-step exchange period, random symbols (constant metric ties), start / end / best end states, both tie-break
flavours, an inconsistent max_error, forced traceback segments (repair path), agreement with the decision-row kernels.
Everything through the C ABI against the scalar oracle, bit-exact."""
import numpy as np
import pytest

import viterbidecodercpp_b200 as v
from common import CODE_BY_NAME, assert_batch_equal, frames, make_cuda_decoder, make_oracle, oracle_batch, random_symbols
from oracle_binding import MODE_SIMD

pytestmark = pytest.mark.gpu

K9_CODES = ["CDMA IS-95A", "CDMA 2000", "Voyager", "DAB Radio"]     # (the name is historical: K = 7 codes ride along; LTE below)


def group_variant(name):
    """vitb_set_variant number of the frame-over-4-lanes history kernel: 4 for K = 9; 104 for K = 7, where 4 is the decision-row
    kernel with 4 lanes per frame PAIR"""
    return 4 if CODE_BY_NAME[name].K == 9 else 104


def hist_group_decoder(name, **kw):
    code = CODE_BY_NAME[name]
    dec, dc = make_cuda_decoder(code, "SOFT16", **kw)
    dec.set_variant(group_variant(name))      # K = 9: default only for batches that fill the GPU; K = 7: only for small batches
    return code, dec, dc


def check_kernel(dec):
    assert (dec.kernel_name.startswith("acs_hist<K9") or dec.kernel_name.startswith("acs_hist<K7")) and ",T4," in dec.kernel_name, dec.kernel_name


@pytest.mark.parametrize("name", K9_CODES)
def test_k9_history_every_record_and_exchange_residue(cuda_lib, name):
    """total_bits 1 .. 56: S = L + K - 1 steps covers every residue modulo 16 (record) x modulo 6 / 4 (exchange period) = lcm 48 / 16,
    ragged last bytes included; plus long frames"""
    code, dec, dc = hist_group_decoder(name)
    ora, _ = make_oracle(code, "SOFT16")
    for L in list(range(1, 57)) + [257, 1003, 2050]:
        n_frames = 19 if L < 100 else 11
        sym = random_symbols(dc, n_frames, (L + code.K - 1) * code.R, seed=3000 + L)
        want = ora.decode_frames(sym, n_frames, L)
        got = dec.decode_batch(sym, L)
        check_kernel(dec)
        assert_batch_equal(got, want, f"{name} L={L}")


@pytest.mark.parametrize("name", K9_CODES)
@pytest.mark.parametrize("ebno", [-3.0, 1.0, 4.0])
def test_k9_history_noisy_frames_renormalise(cuda_lib, name, ebno):
    """AWGN frames long enough for ~10 renormalisations each (the lazy renormalisation path), ragged frame count"""
    code, dec, dc = hist_group_decoder(name)
    ora, _ = make_oracle(code, "SOFT16")
    L = 4096
    tx, sym = frames(code, dc, 77, L, ebno, seed=int(ebno * 10) + 99)
    want = ora.decode_frames(sym, 77, L)
    assert int(want[1].min()) > 0, "every frame must renormalise at least once for this test to mean anything"
    got = dec.decode_batch(sym, L)
    check_kernel(dec)
    assert_batch_equal(got, want, f"{name} Eb/N0={ebno}")


@pytest.mark.parametrize("name", K9_CODES)
def test_k9_history_renormalise_every_step(cuda_lib, name):
    """renormalisation_threshold = 0: the trigger fires after every step (the pending minimum is folded into every table), and a
    threshold just above the start metric: it fires on the very first steps"""
    code = CODE_BY_NAME[name]
    dc = v.DECODE_TYPES["SOFT16"](code.R)
    c = dc.decoder_config
    for thr in (0, 1, c.initial_non_start_error):
        cfg = v.ViterbiDecoder_Config(c.soft_decision_max_error, c.initial_start_error, c.initial_non_start_error, thr)
        dec, _ = make_cuda_decoder(code, "SOFT16", config_override=cfg)
        dec.set_variant(group_variant(name))
        ora, _ = make_oracle(code, "SOFT16", config_override=cfg)
        for L in (5, 47, 300):
            sym = random_symbols(dc, 13, (L + code.K - 1) * code.R, seed=thr + L)
            want = ora.decode_frames(sym, 13, L)
            got = dec.decode_batch(sym, L)
            check_kernel(dec)
            assert_batch_equal(got, want, f"{name} thr={thr} L={L}")


@pytest.mark.parametrize("name", K9_CODES)
def test_k9_history_start_end_and_best_states(cuda_lib, name):
    code, dec, dc = hist_group_decoder(name)
    ora, _ = make_oracle(code, "SOFT16")
    ns = 1 << (code.K - 1)
    for L, start, end in [(100, 5, 0), (100, 0, ns - 1), (61, 9, 137 % ns), (203, ns - 1, 21), (48, 200 % ns, 77 % ns)]:
        sym = random_symbols(dc, 21, (L + code.K - 1) * code.R, seed=L + start + end)
        want = oracle_batch(ora, code, sym, L, start, end)
        got = dec.decode_batch(sym, L, starting_state=start, end_state=end)
        check_kernel(dec)
        assert_batch_equal(got, want, f"{name} L={L} start={start} end={end}")
    # best end state per frame
    for L in (520, 1003):
        n = 23
        sym = random_symbols(dc, n, (L + code.K - 1) * code.R, seed=L)
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
        assert_batch_equal(got, (out, acc, fin), f"{name} L={L} best end state")


@pytest.mark.parametrize("name", K9_CODES)
def test_k9_history_simd_tie_break_and_inconsistent_max_error(cuda_lib, name):
    code = CODE_BY_NAME[name]
    dc = v.DECODE_TYPES["SOFT16"](code.R)
    c = dc.decoder_config
    dec, _ = make_cuda_decoder(code, "SOFT16", tie_break=v.VITB_TIE_SIMD)
    dec.set_variant(group_variant(name))
    ora, _ = make_oracle(code, "SOFT16", mode=MODE_SIMD)
    for L in (13, 64, 500):
        sym = random_symbols(dc, 19, (L + code.K - 1) * code.R, seed=77 + L)
        want = ora.decode_frames(sym, 19, L)
        got = dec.decode_batch(sym, L)
        check_kernel(dec)
        assert_batch_equal(got, want, f"{name} simd tie L={L}")
    cfg = v.ViterbiDecoder_Config(c.soft_decision_max_error + 3, c.initial_start_error, c.initial_non_start_error, c.renormalisation_threshold)
    for tie, mode in ((v.VITB_TIE_SCALAR, 0), (v.VITB_TIE_SIMD, MODE_SIMD)):
        dec, _ = make_cuda_decoder(code, "SOFT16", tie_break=tie, config_override=cfg)
        dec.set_variant(group_variant(name))
        ora, _ = make_oracle(code, "SOFT16", mode=mode, config_override=cfg)
        sym = random_symbols(dc, 19, (300 + code.K - 1) * code.R, seed=3)
        want = ora.decode_frames(sym, 19, 300)
        got = dec.decode_batch(sym, 300)
        check_kernel(dec)
        assert_batch_equal(got, want, f"{name} inconsistent max_error tie={tie}")


@pytest.mark.parametrize("name", K9_CODES)
def test_k9_history_segmented_traceback_and_repair(cuda_lib, name):
    """position-ordered records walked in forced segments: with the warm-up switched off nearly every segment starts from the wrong
    state and is repaired; results stay bit-exact"""
    code, dec, dc = hist_group_decoder(name)
    ora, _ = make_oracle(code, "SOFT16")
    for L, end in [(1000, 0), (2049, 3), (517, 201 % (1 << (code.K - 1)))]:
        sym = random_symbols(dc, 12, (L + code.K - 1) * code.R, seed=L)
        want = oracle_batch(ora, code, sym, L, 0, end)
        for seg, ov in [(0, -1), (7, 0), (5, 1), (16, 3), (1, 0), (1000, 0)]:
            dec.set_traceback_segments(seg, ov)
            got = dec.decode_batch(sym, L, end_state=end)
            check_kernel(dec)
            assert_batch_equal(got, want, f"{name} L={L} seg={seg} overlap={ov}")
    dec.set_traceback_segments(0, -1)


@pytest.mark.parametrize("name", K9_CODES)
def test_k9_history_matches_decision_row_kernels(cuda_lib, name):
    code, dec, dc = hist_group_decoder(name)
    L = 777
    sym = random_symbols(dc, 50, (L + code.K - 1) * code.R, seed=5)
    a = dec.decode_batch(sym, L)
    check_kernel(dec)
    for lanes in ((8, 16) if code.K == 9 else (1, 4)):
        dec.set_variant(lanes)
        b = dec.decode_batch(sym, L)
        assert dec.kernel_name.startswith("acs<") or (code.K == 7 and lanes == 1 and dec.kernel_name.startswith("acs_hist<")), dec.kernel_name
        assert_batch_equal(a, b, f"{name} history vs the T{lanes} kernel")


def test_k7_small_batches_take_the_lane_group_history_kernel(cuda_lib):
    """automatic selection: a K = 7 soft16 batch that leaves most schedulers without a warp of the one-lane kernel (32 frames per
    warp) runs the frame-over-4-lanes kernel, a batch that fills them keeps the one-lane kernel; both bit-exact"""
    code = CODE_BY_NAME["Voyager"]
    dec, dc = make_cuda_decoder(code, "SOFT16")
    assert 104 in dec.variants
    ora, _ = make_oracle(code, "SOFT16")
    L = 200
    for n_frames, want_t4 in ((600, True), (20000, False)):
        tx, sym = frames(code, dc, n_frames, L, 3.0, seed=n_frames)
        got = dec.decode_batch(sym, L)
        assert (",T4," in dec.kernel_name) == want_t4, dec.kernel_name
        assert dec.kernel_name.startswith("acs_hist<K7"), dec.kernel_name
        want = ora.decode_frames(sym, n_frames, L)
        assert_batch_equal(got, want, f"Voyager {n_frames} frames")
    dec.close()


def test_k7_rate_one_third_lane_group_history_kernel(cuda_lib):
    """LTE (R = 3): the kernel reads whole 32-bit words of symbols, so it serves rows of an even number of int16 symbols (even
    L + K - 1); exchange periods of 4 steps are 6 words.  Random symbols, AWGN frames with renormalisations, both tie-breaks."""
    code = CODE_BY_NAME["LTE"]
    for tie, mode in ((v.VITB_TIE_SCALAR, 0), (v.VITB_TIE_SIMD, MODE_SIMD)):
        dec, dc = make_cuda_decoder(code, "SOFT16", tie_break=tie)
        dec.set_variant(104)
        ora, _ = make_oracle(code, "SOFT16", mode=mode)
        for L in list(range(2, 40, 2)) + [1000, 2050]:
            sym = random_symbols(dc, 21, (L + code.K - 1) * code.R, seed=L)
            want = ora.decode_frames(sym, 21, L)
            got = dec.decode_batch(sym, L)
            check_kernel(dec)
            assert_batch_equal(got, want, f"LTE tie={tie} L={L}")
        tx, sym = frames(code, dc, 77, 4096, 1.0, seed=5)
        want = ora.decode_frames(sym, 77, 4096)
        assert int(want[1].min()) > 0
        got = dec.decode_batch(sym, 4096)
        check_kernel(dec)
        assert_batch_equal(got, want, f"LTE tie={tie} AWGN")
        # rows the kernel cannot fetch (odd number of symbols): the pinned variant does not apply, the result is still exact
        sym = random_symbols(dc, 9, (33 + code.K - 1) * code.R, seed=1)
        assert_batch_equal(dec.decode_batch(sym, 33), ora.decode_frames(sym, 9, 33), "LTE odd row length")
        assert ",T4," not in dec.kernel_name or not dec.kernel_name.startswith("acs_hist"), dec.kernel_name
        dec.close()
