"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libvitref.so, built from /root/reference in place).
Run in the build container:  python tests/golden/make_golden.py
Each file: code, config, noisy input symbols and what the reference decoder produced (decoded bytes, accumulated error, final
error for every frame; all decision rows and final metrics of frame 0).  mode 0 = ViterbiDecoder_Scalar, 1 = ViterbiDecoder_AVX_*."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_binding as ob  # noqa: E402
from viterbidecodercpp_b200 import synth  # noqa: E402

CASES = [  # (code, decode type, frames, bits, Eb/N0 dB)
    ("voyager", "soft16", 8, 1024, 3.0), ("voyager", "hard8", 8, 2048, 4.0), ("voyager", "hard8", 8, 512, 0.0), ("voyager", "soft8", 6, 512, 2.0),
    ("k3r2", "soft16", 4, 256, 2.0), ("k5r2", "hard8", 4, 256, 3.0), ("lte", "soft16", 6, 512, 0.0), ("lte", "hard8", 6, 512, 3.0),
    ("dab", "soft16", 6, 768, 0.0), ("dab", "hard8", 6, 768, 2.0), ("is95a", "soft16", 3, 1024, 4.0), ("is95a", "hard8", 3, 512, 3.0),
    ("cdma2000", "soft16", 3, 512, 0.0), ("cassini", "soft16", 1, 128, 2.0), ("cassini", "hard8", 1, 128, 3.0),
]

if __name__ == "__main__":
    assert ob.have_ref(), "build oracle/_ref first (needs /root/reference)"
    for i, (name, dt, F, L, ebno) in enumerate(CASES):
        K, R, G = ob.CODES[name]
        sb, high, low, cfg = ob.preset(dt, R)
        tx, sym = synth.make_frames(K, R, G, F, L, high, low, sb, ebno, 1000 + i)
        for mode, impl in ((0, ob.IMPL_SCALAR), (1, ob.IMPL_AVX)):
            try:
                r = ob.ref_decode(K, R, G, sb, high, low, cfg, impl, sym, F, L, want_decisions=True, want_metrics=True)
            except RuntimeError:
                continue
            path = os.path.join(HERE, f"{name}_{dt}_{'scalar' if mode == 0 else 'avx'}_{L}b.npz")
            np.savez_compressed(path, K=K, R=R, G=np.array(G, dtype=np.uint32), soft_bytes=sb, high=high, low=low, cfg=np.array(cfg, dtype=np.uint64),
                                mode=mode, L=L, EbNo_dB=ebno, symbols=sym, tx=tx, bytes=r["bytes"], acc=r["acc"], final=r["final"],
                                decisions0=r["decisions"][0], metrics0=r["metrics"][0])
            print("wrote", os.path.basename(path), os.path.getsize(path))
