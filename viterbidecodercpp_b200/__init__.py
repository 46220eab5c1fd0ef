"""viterbidecodercpp_b200 -- B200 (sm_100a) batched Viterbi decoder behind the reference's decoder interface.

The compute lives in csrc/libviterbi_b200.so (hand-written CUDA + the C ABI of include/viterbi_b200.h); this package is the
thin host-side mirror of the reference interface used by tests and bench.py.  No CPU fallback exists.
"""
from ._lib import LIB_PATH, VITB_END_STATE_BEST, VITB_TIE_SCALAR, VITB_TIE_SIMD, load as load_library
from .presets import (COMMON_CODES, DECODE_TYPES, Code, Decoder_Config, ViterbiDecoder_Config, dab_fic_keep_schedule, dab_pi,
                      get_hard8_decoding_config, get_soft16_decoding_config, get_soft8_decoding_config)
from .decoder import ViterbiBranchTable, ViterbiDecoder_CUDA, ViterbiError, decode_batch_multi, decode_batch_multi_raw

__all__ = [
    "LIB_PATH", "VITB_END_STATE_BEST", "VITB_TIE_SCALAR", "VITB_TIE_SIMD", "load_library", "COMMON_CODES", "DECODE_TYPES", "Code", "Decoder_Config",
    "ViterbiDecoder_Config", "dab_fic_keep_schedule", "dab_pi", "get_hard8_decoding_config", "get_soft16_decoding_config",
    "get_soft8_decoding_config", "ViterbiBranchTable", "ViterbiDecoder_CUDA", "ViterbiError", "decode_batch_multi", "decode_batch_multi_raw",
]
