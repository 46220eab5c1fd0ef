"""shared helpers for the parity tests"""
import numpy as np

import viterbidecodercpp_b200 as v
from viterbidecodercpp_b200 import synth
from oracle_binding import OracleDecoder, MODE_SCALAR, MODE_SIMD

CODE_BY_NAME = {c.name: c for c in v.COMMON_CODES}
PAIR_CODES = ["Basic K=3 R=1/2", "Basic K=5 R=1/2", "Voyager", "LTE", "DAB Radio"]
GPU_CODES = PAIR_CODES + ["CDMA IS-95A", "CDMA 2000", "Cassini"]      # the whole catalogue (examples/helpers/common_codes.h)


def make_cuda_decoder(code, decode_type, tie_break=0, device=0, config_override=None):
    dc = v.DECODE_TYPES[decode_type](code.R)
    cfg = config_override or dc.decoder_config
    bt = v.ViterbiBranchTable(code.K, code.R, code.G, dc.soft_decision_high, dc.soft_decision_low, dc.soft_bytes)
    return v.ViterbiDecoder_CUDA(bt, cfg, device=device, tie_break=tie_break), dc


def make_oracle(code, decode_type, mode=MODE_SCALAR, config_override=None):
    dc = v.DECODE_TYPES[decode_type](code.R)
    c = config_override or dc.decoder_config
    cfg = [c.soft_decision_max_error, c.initial_start_error, c.initial_non_start_error, c.renormalisation_threshold]
    return OracleDecoder(code.K, code.R, code.G, dc.soft_bytes, dc.soft_decision_high, dc.soft_decision_low, cfg, mode), dc


def frames(code, dc, n_frames, total_bits, EbNo_dB, seed):
    return synth.make_frames(code.K, code.R, code.G, n_frames, total_bits, dc.soft_decision_high, dc.soft_decision_low,
                             dc.soft_bytes, EbNo_dB, seed)


def assert_batch_equal(got, want, what=""):
    gb, ga, gf = got
    wb, wa, wf = want
    bad = np.nonzero((gb != wb).any(axis=1))[0]
    assert bad.size == 0, f"{what}: decoded bytes differ in {bad.size}/{gb.shape[0]} frames (first {bad[:5]})"
    assert (ga == wa).all(), f"{what}: accumulated error differs in {int((ga != wa).sum())} frames"
    assert (gf == wf).all(), f"{what}: final error differs in {int((gf != wf).sum())} frames"


def random_symbols(dc, n_frames, n_sym, seed, pad=0):
    rng = np.random.default_rng(seed)
    dt = np.int8 if dc.soft_bytes == 1 else np.int16
    s = rng.integers(dc.soft_decision_low, dc.soft_decision_high + 1, size=(n_frames, n_sym + pad)).astype(dt)
    return s


def oracle_batch(ora, code, sym, L, start=0, end=0):
    """the call protocol of run_simple.cpp:76-80 per frame, with explicit start / end states (core.h:195-236)"""
    n = sym.shape[0]
    out = np.zeros((n, (L + 7) // 8), dtype=np.uint8)
    acc = np.zeros(n, dtype=np.uint64)
    fin = np.zeros(n, dtype=np.uint32)
    ora.set_traceback_length(L)
    for f in range(n):
        ora.reset(start)
        acc[f] = ora.update(sym[f])
        fin[f] = ora.get_error(end)
        out[f] = ora.chainback(L, end)
    return out, acc, fin
