"""Adversarial inputs for EVERY kernel whose decisions ride in spare low bits of the metric registers ("tags"):

    acs_hist_kernel        K <= 7, uint8 (two frames per register, 8-step records) and uint16 (16-step records)   variant 1, batch
    acs_hist_group_kernel  K = 9, uint16, frame over 4 lanes                                                       variant 4, batch
    acs_hist_cta_kernel    K = 15, uint16, frame per CTA                                                           variant 256, batch
    acs_pair_kernel        K <= 7, uint8: tagged butterfly, decision rows (streaming calls, batch with history off)

Round 1 tried the tagged butterfly in the lane-group decision-row kernels too and saw rare wrong decision bits for two codes
(profiles/r01_summary.md); that experiment was reverted without being committed, so its failing input cannot be replayed.  What a
tag scheme can get wrong is known, though: a tie broken the wrong way, a tag bit surviving into the next step or record, a tag
carrying into the metric field when the metric wraps or is renormalised, the two frames of a register interfering.  The inputs
below drive each of those on purpose, for every tag-based kernel that ships, against the scalar oracle:

    all-equal symbols        every branch metric of a step is the same: EVERY compare is a tie, the whole decision row is tie-breaks
    two-level symbols        only +high / -high (or low / high): metrics move in lock step, ties stay frequent, metrics grow fastest
    saturating growth        stock thresholds with maximum-error symbols: uint8 metrics wrap modulo 256 (the reference's scalar wraps
                             too) and renormalise at the same time
    threshold 0 / 1          a renormalisation after every step (and one pending after the last step)
    mixed pairs              frame A all ties, frame B random (the halves of a packed register must not see each other)
"""
import numpy as np
import pytest

import viterbidecodercpp_b200 as v
from common import CODE_BY_NAME, assert_batch_equal, make_cuda_decoder, make_oracle, oracle_batch

pytestmark = pytest.mark.gpu

# (code, decode type, variant to pin, expected kernel-name prefix, frames, bits)
TAG_KERNELS = [
    ("Voyager", "HARD8", 1, "acs_hist<K7", 130, 203),
    ("Voyager", "SOFT8", 1, "acs_hist<K7", 130, 203),
    ("Voyager", "SOFT16", 1, "acs_hist<K7", 70, 203),
    ("LTE", "HARD8", 1, "acs_hist<K7", 130, 131),
    ("LTE", "SOFT16", 1, "acs_hist<K7", 70, 131),
    ("DAB Radio", "HARD8", 1, "acs_hist<K7", 130, 131),
    ("DAB Radio", "SOFT8", 1, "acs_hist<K7", 130, 131),
    ("DAB Radio", "SOFT16", 1, "acs_hist<K7", 70, 131),
    ("Basic K=5 R=1/2", "HARD8", 1, "acs_hist<K5", 130, 99),
    ("Basic K=3 R=1/2", "SOFT8", 1, "acs_hist<K3", 130, 99),
    ("CDMA IS-95A", "SOFT16", 4, "acs_hist<K9", 19, 150),
    ("Cassini", "SOFT16", 256, "acs_hist<K15", 3, 90),
]


def adversarial_batches(dc, n_frames, n_sym, seed):
    hi, lo = dc.soft_decision_high, dc.soft_decision_low
    dt = np.int8 if dc.soft_bytes == 1 else np.int16
    rng = np.random.default_rng(seed)
    out = {}
    for name, val in (("all zero", 0), ("all high", hi), ("all low", lo)):
        out[name] = np.full((n_frames, n_sym), val, dtype=dt)
    out["two-level"] = rng.choice(np.array([lo, hi], dtype=dt), size=(n_frames, n_sym))
    out["three-level"] = rng.choice(np.array([lo, 0, hi], dtype=dt), size=(n_frames, n_sym))
    mixed = rng.integers(lo, hi + 1, size=(n_frames, n_sym)).astype(dt)
    mixed[0::2] = 0                                    # frame A of every register pair: all ties; frame B: random
    out["mixed pairs"] = mixed
    alt = np.zeros((n_frames, n_sym), dtype=dt)
    alt[:, 0::2] = hi
    alt[:, 1::2] = lo
    out["alternating"] = alt
    return out


@pytest.mark.parametrize("tie", [0, 1])
@pytest.mark.parametrize("name,decode_type,variant,prefix,n_frames,L", TAG_KERNELS)
def test_tag_kernels_on_tie_storms(cuda_lib, name, decode_type, variant, prefix, n_frames, L, tie):
    """stock configuration, both tie-break flavours (the SIMD flavour tags the other path and inverts the record).  VITB_TIE_SIMD
    reproduces the SIMD decoders' tie-break, not their saturating adds (include/viterbi_b200.h), and the oracle's SIMD mode
    saturates: a batch in which any addition of the oracle saturated is outside what that flavour promises (the saturating flavour,
    VITB_TIE_SIMD_SAT, is covered by tests/test_gpu_simd_sat.py) and is only checked in the scalar flavour."""
    code = CODE_BY_NAME[name]
    dec, dc = make_cuda_decoder(code, decode_type, tie_break=tie)
    ora, _ = make_oracle(code, decode_type, mode=tie)
    dec.set_variant(variant)
    n_sym = (L + code.K - 1) * code.R
    checked = 0
    for what, sym in adversarial_batches(dc, n_frames, n_sym, seed=L).items():
        ora.clipped(clear=True)
        want = oracle_batch(ora, code, sym, L) if code.K >= 15 else ora.decode_frames(sym, n_frames, L)
        if tie == 1 and ora.clipped() > 0:
            continue
        got = dec.decode_batch(sym, L)
        assert dec.kernel_name.startswith(prefix), dec.kernel_name
        assert_batch_equal(got, want, f"{name} {decode_type} tie={tie} {what}")
        checked += 1
    assert checked >= (3 if tie == 0 else 1)
    dec.close()


@pytest.mark.parametrize("thr_kind", ["zero", "one", "tight"])
@pytest.mark.parametrize("name,decode_type,variant,prefix,n_frames,L", [k for k in TAG_KERNELS if k[0] != "Cassini"])
def test_tag_kernels_renormalising_constantly(cuda_lib, name, decode_type, variant, prefix, n_frames, L, thr_kind):
    """threshold 0 / 1 (every step renormalises) and a threshold just above the non-start error (metrics hover around it, so the
    trigger fires at irregular steps and the two frames of a register fire at different steps)"""
    code = CODE_BY_NAME[name]
    dc = v.DECODE_TYPES[decode_type](code.R)
    c = dc.decoder_config
    thr = {"zero": 0, "one": 1, "tight": c.initial_non_start_error + 1}[thr_kind]
    cfg = v.ViterbiDecoder_Config(c.soft_decision_max_error, c.initial_start_error, c.initial_non_start_error, thr)
    dec, _ = make_cuda_decoder(code, decode_type, config_override=cfg)
    ora, _ = make_oracle(code, decode_type, config_override=cfg)
    dec.set_variant(variant)
    n_sym = (L + code.K - 1) * code.R
    for what, sym in adversarial_batches(dc, n_frames, n_sym, seed=7 * L).items():
        if what in ("all high", "all low"):
            continue
        want = ora.decode_frames(sym, n_frames, L)
        got = dec.decode_batch(sym, L)
        assert dec.kernel_name.startswith(prefix), dec.kernel_name
        assert_batch_equal(got, want, f"{name} {decode_type} thr={thr} {what}")
    dec.close()


@pytest.mark.parametrize("name,decode_type", [("Voyager", "HARD8"), ("LTE", "SOFT8"), ("DAB Radio", "HARD8"), ("DAB Radio", "SOFT8"),
                                               ("Basic K=5 R=1/2", "SOFT8")])
def test_tagged_decision_row_kernel_on_tie_storms(cuda_lib, name, decode_type):
    """acs_pair_kernel's tagged butterfly (uint8 metrics): batch calls with the history kernel switched off AND the streaming calls,
    whose decision rows are compared with the oracle's row for row"""
    code = CODE_BY_NAME[name]
    dec, dc = make_cuda_decoder(code, decode_type)
    ora, _ = make_oracle(code, decode_type)
    dec.set_variant(1)
    dec.set_history_kernel(False)
    L = 160
    n_sym = (L + code.K - 1) * code.R
    for what, sym in adversarial_batches(dc, 130, n_sym, seed=3).items():
        want = ora.decode_frames(sym, 130, L)
        got = dec.decode_batch(sym, L)
        assert dec.kernel_name.startswith("acs<"), dec.kernel_name
        assert_batch_equal(got, want, f"{name} {decode_type} rows kernel {what}")
        # streaming: one frame in ragged pieces, every decision row and metric against the oracle
        f = sym[1]
        dec.set_traceback_length(L)
        dec.reset()
        ora.set_traceback_length(L)
        ora.reset()
        pos, acc_g, acc_o = 0, 0, 0
        for piece in (1, 5, 8, 33, 10 ** 6):
            n = min(piece * code.R, f.size - pos)
            if n <= 0:
                break
            acc_g += dec.update(f[pos:pos + n])
            acc_o += ora.update(f[pos:pos + n])
            pos += n
        assert acc_g == acc_o
        assert (dec.m_metrics == ora.metrics()).all(), f"{name} {decode_type} {what}: streaming metrics"
        rows = L + code.K - 1
        assert (dec.m_decisions(0, rows) == ora.decisions(rows)).all(), f"{name} {decode_type} {what}: streaming decision rows"
    dec.close()
