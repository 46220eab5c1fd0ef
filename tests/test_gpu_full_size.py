"""BASELINE.json's configurations at FULL size, checked through size-independent properties (the scalar oracle would need minutes to
hours at these sizes): noise-free encode -> decode round trips must return the transmitted bytes with total path error 0 (the
reference's own pass criterion, examples/run_tests.cpp:184 and run_simple.cpp:81-93; for the punctured DAB frames the known
answer 792 x 127, run_punctured_decoder.cpp), every kernel variant must produce identical results, and a strided sample of noisy
frames is compared with the oracle bit for bit."""
import zlib

import numpy as np
import pytest

import viterbidecodercpp_b200 as v
from viterbidecodercpp_b200 import synth
from common import CODE_BY_NAME, assert_batch_equal, make_cuda_decoder, make_oracle

pytestmark = pytest.mark.gpu

FULL = [  # (name, code, decode type, frames, bits, punctured)
    ("cfg2", "Voyager", "HARD8", 65536, 2048, False),
    ("cfg3", "CDMA IS-95A", "SOFT16", 16384, 8192, False),
    ("cfg4", "DAB Radio", "SOFT16", 65536, 768, True),
    ("cfg5", "Cassini", "SOFT16", 1024, 16384, False),
]


def _gen(code, dc, F, L, ebno, seed, keep):
    tx_all, sym_all = [], []
    for f0 in range(0, F, 4096):
        n = min(4096, F - f0)
        tx, sym = synth.make_frames(code.K, code.R, code.G, n, L, dc.soft_decision_high, dc.soft_decision_low, dc.soft_bytes, ebno, seed + f0)
        if keep is not None:
            sym = synth.puncture(sym, keep)
        tx_all.append(tx)
        sym_all.append(sym)
    return np.concatenate(tx_all), np.concatenate(sym_all)


@pytest.mark.parametrize("name,code_name,decode_type,F,L,punctured", FULL)
def test_full_size_noise_free_round_trip(cuda_lib, name, code_name, decode_type, F, L, punctured):
    code = CODE_BY_NAME[code_name]
    dec, dc = make_cuda_decoder(code, decode_type)
    keep = np.asarray(v.dab_fic_keep_schedule(), dtype=bool) if punctured else None
    if keep is not None:
        dec.set_puncture_schedule(keep.astype(np.uint8), 0)
    tx, sym = _gen(code, dc, F, L, None, 2024, keep)
    out, acc, fin = dec.decode_batch(sym, L)
    assert (out == tx).all(), f"{name}: {int((out != tx).any(axis=1).sum())} frames decoded wrongly"
    expected_error = 792 * 127 if punctured else 0
    assert ((acc + fin) == expected_error).all()


@pytest.mark.parametrize("name,code_name,decode_type,F,L,punctured", FULL[:3])
def test_full_size_noisy_variants_agree_and_sample_matches_oracle(cuda_lib, name, code_name, decode_type, F, L, punctured):
    code = CODE_BY_NAME[code_name]
    dec, dc = make_cuda_decoder(code, decode_type)
    ora, _ = make_oracle(code, decode_type)
    keep = np.asarray(v.dab_fic_keep_schedule(), dtype=bool) if punctured else None
    if keep is not None:
        dec.set_puncture_schedule(keep.astype(np.uint8), 0)
    F = F // 4                       # a quarter of the batch keeps host-side generation short; still far beyond oracle reach
    tx, sym = _gen(code, dc, F, L, 3.0, 7, keep)
    sums = set()
    ref = None
    for lanes in dec.variants:
        dec.set_variant(lanes)
        out, acc, fin = dec.decode_batch(sym, L)
        sums.add((zlib.crc32(out.tobytes()), zlib.crc32(acc.tobytes()), zlib.crc32(fin.tobytes())))   # checksum of checksums
        ref = (out, acc, fin)
    assert len(sums) == 1, f"{name}: kernel variants disagree"
    idx = np.arange(0, F, max(1, F // 48))
    dep = sym[idx]
    if keep is not None:
        full = np.zeros((idx.size, keep.size), dtype=sym.dtype)
        full[:, keep] = dep
        dep = full
    want = ora.decode_frames(dep, idx.size, L)
    assert_batch_equal((ref[0][idx], ref[1][idx], ref[2][idx]), want, f"{name} sample vs oracle")


@pytest.mark.parametrize("ebno", [1.0, 4.0])
def test_cfg5_full_length_noisy_frames_match_oracle(cuda_lib, ebno):
    """Cassini K=15 frames at BASELINE.json's FULL length (16384 bits, ~130 renormalisations per frame: the speculation / rollback /
    replay path of csrc/acs_cta.cuh and the 18-segment traceback) against the scalar oracle, bit for bit; 9 frames so that the last
    CTA holds a frame pair with one padding frame."""
    code = CODE_BY_NAME["Cassini"]
    dec, dc = make_cuda_decoder(code, "SOFT16")
    ora, _ = make_oracle(code, "SOFT16")
    F, L = 9, 16384
    tx, sym = synth.make_frames(code.K, code.R, code.G, F, L, dc.soft_decision_high, dc.soft_decision_low, dc.soft_bytes, ebno, 4242)
    want = ora.decode_frames(sym, F, L)
    assert int(want[1].min()) > 0          # every frame renormalised
    for lanes in dec.variants:
        dec.set_variant(lanes)
        got = dec.decode_batch(sym, L)
        assert_batch_equal(got, want, f"cfg5 full length Eb/N0={ebno} dB variant {lanes}")
