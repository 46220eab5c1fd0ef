"""GPU output against the UNMODIFIED reference decoder itself (oracle/_ref/libvitref.so = the reference's headers compiled in place by
oracle/Makefile), without the C restatement in between: ViterbiDecoder_Scalar for the scalar tie-break, ViterbiDecoder_AVX_u16/u8
for VITB_TIE_SIMD where no metric saturates.  Skipped where the reference build is absent (it travels to the GPU box with the repo
snapshot; /root/reference itself is not needed at run time)."""
import numpy as np
import pytest

import viterbidecodercpp_b200 as v
import oracle_binding as ob
from common import CODE_BY_NAME, GPU_CODES, assert_batch_equal, frames, make_cuda_decoder

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref/libvitref.so not built (needs /root/reference at build time)")]


def cfg_list(c):
    return [c.soft_decision_max_error, c.initial_start_error, c.initial_non_start_error, c.renormalisation_threshold]


@pytest.mark.parametrize("decode_type", ["SOFT16", "SOFT8", "HARD8"])
@pytest.mark.parametrize("name", GPU_CODES)
def test_gpu_equals_reference_scalar_decoder(cuda_lib, name, decode_type):
    """every catalogue code x decode type, noisy frames, every compiled kernel variant: decoded bytes, sum of renormalisation minima
    and final metric equal ViterbiDecoder_Scalar's (viterbi_decoder_scalar.h:28-153 + core.h:195-236)"""
    code = CODE_BY_NAME[name]
    dec, dc = make_cuda_decoder(code, decode_type)
    n_frames, L = (5, 1024) if code.K >= 15 else (67, 1000)
    tx, sym = frames(code, dc, n_frames, L, 2.0, seed=31)
    r = ob.ref_decode(code.K, code.R, code.G, dc.soft_bytes, dc.soft_decision_high, dc.soft_decision_low, cfg_list(dc.decoder_config),
                      ob.IMPL_SCALAR, sym, n_frames, L)
    want = (r["bytes"], r["acc"], r["final"])
    for lanes in dec.variants:
        dec.set_variant(lanes)
        got = dec.decode_batch(sym, L)
        assert_batch_equal(got, want, f"{name} {decode_type} variant {lanes} vs reference scalar ({dec.kernel_name})")


@pytest.mark.parametrize("name,decode_type", [("Voyager", "SOFT16"), ("DAB Radio", "SOFT16"), ("CDMA IS-95A", "SOFT16"), ("Voyager", "HARD8")])
def test_gpu_simd_flavour_equals_reference_avx_decoder(cuda_lib, name, decode_type):
    """VITB_TIE_SIMD against the reference's AVX2 decoder (x86/viterbi_decoder_avx_u16.h / avx_u8.h) on the stock presets at a
    moderate SNR, where the saturating adds never saturate"""
    code = CODE_BY_NAME[name]
    dec, dc = make_cuda_decoder(code, decode_type, tie_break=v.VITB_TIE_SIMD)
    n_frames, L = 67, 1000
    tx, sym = frames(code, dc, n_frames, L, 3.0, seed=32)
    r = ob.ref_decode(code.K, code.R, code.G, dc.soft_bytes, dc.soft_decision_high, dc.soft_decision_low, cfg_list(dc.decoder_config),
                      ob.IMPL_AVX, sym, n_frames, L)
    want = (r["bytes"], r["acc"], r["final"])
    for lanes in dec.variants:
        dec.set_variant(lanes)
        got = dec.decode_batch(sym, L)
        assert_batch_equal(got, want, f"{name} {decode_type} variant {lanes} vs reference AVX2 ({dec.kernel_name})")
