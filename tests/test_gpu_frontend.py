"""Device-side front end (csrc/frontend.cuh): the encoder and the quantiser are exact against the host restatements (and so
against the reference: tests/test_oracle_vs_reference.py pins those), the AWGN statistics and the resulting bit error rates
agree with the host generator, and a whole generate -> decode -> count trial runs on the device."""
import numpy as np
import pytest

import viterbidecodercpp_b200 as v
from viterbidecodercpp_b200 import synth
from common import CODE_BY_NAME, GPU_CODES, make_cuda_decoder, make_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("decode_type", ["SOFT16", "HARD8"])
@pytest.mark.parametrize("name", GPU_CODES)
def test_device_encoder_is_exact(cuda_lib, name, decode_type):
    """noise free: symbols == high/low image of the shift-register encoder's output incl. the K-1 tail steps"""
    code = CODE_BY_NAME[name]
    dec, dc = make_cuda_decoder(code, decode_type)
    L = 256 if code.K < 15 else 64
    tx, sym = dec.synth_frames(37, L, None, seed=5)
    bits = synth.conv_encode(code.K, code.R, code.G, tx).reshape(37, -1)
    want = np.where(bits > 0, dc.soft_decision_high, dc.soft_decision_low)
    assert (sym == want).all()
    assert len({bytes(r) for r in tx}) == 37 and tx.std() > 50          # frames differ, bytes look uniform
    out, acc, fin = dec.decode_batch(sym, L)
    assert (out == tx).all() and ((acc + fin) == 0).all()


def test_device_quantiser_is_exact(cuda_lib):
    """soft = clamp(round(x * scale + mean), low, high) with std::round semantics (run_snr_ber.cpp:352-359)"""
    code = CODE_BY_NAME["Voyager"]
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(0, 1.2, 200000).astype(np.float32), np.array([0.5, -0.5, 1.5, -1.5, 2.5, 1e9, -1e9, 0.0], dtype=np.float32) / np.float32(97.3)])
    for decode_type, scale, mean in (("SOFT16", 97.3, 0.0), ("SOFT8", 2.2, 0.0), ("HARD8", 0.8, 0.0)):
        dec, dc = make_cuda_decoder(code, decode_type)
        got = dec.quantise(x, scale, mean)
        y = x * np.float32(scale) + np.float32(mean)
        r = np.where(y >= 0, np.floor(y + np.float32(0.5)), np.ceil(y - np.float32(0.5)))       # halves away from zero
        want = np.clip(r, dc.soft_decision_low, dc.soft_decision_high)
        assert (got == want).all(), decode_type


def test_device_puncturing_matches_host(cuda_lib):
    code = CODE_BY_NAME["DAB Radio"]
    dec, dc = make_cuda_decoder(code, "SOFT16")
    keep = np.asarray(v.dab_fic_keep_schedule(), dtype=bool)
    dec.set_puncture_schedule(keep.astype(np.uint8), 0)
    tx, rx = dec.synth_frames(50, 768, None, seed=9, row=2304)
    bits = synth.conv_encode(code.K, code.R, code.G, tx).reshape(50, -1)
    want = np.where(bits > 0, 127, -127)[:, keep]
    assert (rx == want).all()
    out, acc, fin = dec.decode_batch(rx, 768)
    assert (out == tx).all() and ((acc + fin) == 792 * 127).all()


def test_device_awgn_statistics_and_ber(cuda_lib):
    """same Eb/N0 -> same symbol statistics and (within sampling noise) the same bit error rate as the host generator"""
    code = CODE_BY_NAME["Voyager"]
    dec, dc = make_cuda_decoder(code, "SOFT16")
    F, L, ebno = 4000, 1024, 2.0
    tx_d, sym_d = dec.synth_frames(F, L, ebno, seed=123)
    tx_h, sym_h = synth.make_frames(code.K, code.R, code.G, F, L, dc.soft_decision_high, dc.soft_decision_low, dc.soft_bytes, ebno, 321)
    bits_d = synth.conv_encode(code.K, code.R, code.G, tx_d).reshape(F, -1)
    bits_h = synth.conv_encode(code.K, code.R, code.G, tx_h).reshape(F, -1)
    sd = np.where(bits_d > 0, 1, -1) * sym_d.astype(np.float64)       # fold the sign: distribution of the received amplitude
    sh = np.where(bits_h > 0, 1, -1) * sym_h.astype(np.float64)
    assert abs(sd.mean() - sh.mean()) < 0.5 and abs(sd.std() - sh.std()) < 0.5
    ber_d = np.unpackbits(dec.decode_batch(sym_d, L)[0] ^ tx_d).mean()
    ber_h = np.unpackbits(dec.decode_batch(sym_h, L)[0] ^ tx_h).mean()
    assert 1e-4 < ber_h < 1e-2
    assert 0.7 < ber_d / ber_h < 1.4
    # different seeds give different noise, the same seed reproduces
    assert (dec.synth_frames(10, L, ebno, seed=7)[1] == dec.synth_frames(10, L, ebno, seed=7)[1]).all()
    assert (dec.synth_frames(10, L, ebno, seed=7)[1] != dec.synth_frames(10, L, ebno, seed=8)[1]).any()


def test_ber_trial_on_device(cuda_lib):
    """generate + decode + count without leaving the GPU; BER falls with Eb/N0 and is 0 without noise"""
    code = CODE_BY_NAME["Voyager"]
    dec, dc = make_cuda_decoder(code, "HARD8")
    F, L = 8192, 2048
    assert dec.ber_trial(F, L, None, seed=1) == 0
    e = [dec.ber_trial(F, L, x, seed=2) / (F * L) for x in (2.0, 4.0, 6.0)]
    assert e[0] > e[1] > e[2] >= 0
    assert 1e-4 < e[1] < 3e-3          # hard decision K=7 R=1/2 at 4 dB: ~8e-4 (bench.py reports 8.4e-4 for host-generated frames)
