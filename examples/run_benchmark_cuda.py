#!/usr/bin/env python
"""Timing harness printing the JSON schema of the reference's examples/run_benchmark.cpp:297-326, so examples/parse_benchmark.py
of the reference (symbols/s = total_symbols / update_symbols_ns, bits/s = total_input_bits / chainback_bits_ns) reads GPU results
unchanged.  One "sample" is one batch of frames; the per-frame nanoseconds are batch time / frames, i.e. the aggregate rate.

    python examples/run_benchmark_cuda.py [-c code_index ...] [-d SOFT16|HARD8 ...] [-n frames] [-L bits] [-T samples]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import viterbidecodercpp_b200 as v  # noqa: E402


def run(code, decode_type, frames, bits, samples):
    dc = v.DECODE_TYPES[decode_type](code.R)
    bt = v.ViterbiBranchTable(code.K, code.R, code.G, dc.soft_decision_high, dc.soft_decision_low, dc.soft_bytes)
    dec = v.ViterbiDecoder_CUDA(bt, dc.decoder_config)
    tx, sym = dec.synth_frames(frames, bits, None, seed=1)        # noise-free frame like run_benchmark.cpp:199-211
    dec.set_profiling(True)
    upd, chain = [], []
    for k in range(samples + 2):
        out, _, _ = dec.decode_batch(sym, bits, want=("bytes",))
        st = dec.stage_ms()
        if k >= 2:                                               # two warm-up batches
            upd.append(int((st["ingest"] + st["acs"]) * 1e6 / frames))
            chain.append(int(max(st["traceback"], 1e-6) * 1e6 / frames))
    assert (out == tx).all()
    return {"name": code.name, "decode_type": decode_type, "simd_type": "SIMD_CUDA", "K": code.K, "R": code.R, "G": code.G,
            "total_input_bits": bits, "total_symbols": (bits + code.K - 1) * code.R, "frames_per_batch": frames,
            "update_symbols_ns": upd, "chainback_bits_ns": chain}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-c", type=int, nargs="*", default=[2, 4, 5])
    ap.add_argument("-d", nargs="*", default=["SOFT16", "HARD8"])
    ap.add_argument("-n", type=int, default=16384)
    ap.add_argument("-L", type=int, default=2048)
    ap.add_argument("-T", type=int, default=10)
    a = ap.parse_args()
    print(json.dumps([run(v.COMMON_CODES[c], d, a.n, a.L, a.T) for c in a.c for d in a.d], indent=1))
