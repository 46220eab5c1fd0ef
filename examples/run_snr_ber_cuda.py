#!/usr/bin/env python
"""BER-vs-Eb/N0 sweep on the GPU, printing the JSON schema of the reference's examples/run_snr_ber.cpp:413-441 (so
examples/plot_snr_ber.py of the reference reads it unchanged; "simd_type" is "SIMD_CUDA").  Frames are generated, decoded and
counted on the device (vitb_ber_trial): nothing but the error counts crosses PCIe.

    python examples/run_snr_ber_cuda.py [-c code_index ...] [-d SOFT16|SOFT8|HARD8 ...] [-n frames_per_trial] [-L bits] [-S seed]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import viterbidecodercpp_b200 as v  # noqa: E402


def sweep(code, decode_type, frames, bits, seed, max_points, min_errors):
    dc = v.DECODE_TYPES[decode_type](code.R)
    bt = v.ViterbiBranchTable(code.K, code.R, code.G, dc.soft_decision_high, dc.soft_decision_low, dc.soft_bytes)
    dec = v.ViterbiDecoder_CUDA(bt, dc.decoder_config)
    ebno, ber = [], []
    for point in range(max_points):                       # get_test_range: start 0 dB, step 0.5 dB (run_snr_ber.cpp:218-231)
        e = 0.5 * point
        errors, total, trial = 0, 0, 0
        while errors < min_errors and trial < 64:         # more frames until enough errors were seen (run_snr_ber.cpp:367-369)
            errors += dec.ber_trial(frames, bits, e, seed=seed + 1000 * point + trial)
            total += frames * bits
            trial += 1
        ebno.append(e)
        ber.append(errors / total)
        print(f"name='{code.name}',K={code.K},R={code.R},decode={decode_type},simd=SIMD_CUDA,EbNo_dB={e:.1f},BER={ber[-1]:.3e},bits={total}", file=sys.stderr)
        if errors == 0:
            break
    return {"name": code.name, "decode_type": decode_type, "simd_type": "SIMD_CUDA", "K": code.K, "R": code.R, "G": code.G,
            "EbNo_dB": [round(x, 1) for x in ebno], "ber": [float(f"{x:.3e}") for x in ber]}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-c", type=int, nargs="*", default=[2], help="code indices in the catalogue (default: 2 = Voyager)")
    ap.add_argument("-d", nargs="*", default=["SOFT16", "HARD8"])
    ap.add_argument("-n", type=int, default=16384, help="frames per trial")
    ap.add_argument("-L", type=int, default=2048, help="data bits per frame")
    ap.add_argument("-S", type=int, default=1234)
    ap.add_argument("--max-points", type=int, default=20)
    ap.add_argument("--min-errors", type=int, default=1000)
    a = ap.parse_args()
    out = [sweep(v.COMMON_CODES[c], d, a.n, a.L, a.S, a.max_points, a.min_errors) for c in a.c for d in a.d]
    print(json.dumps(out, indent=1))
