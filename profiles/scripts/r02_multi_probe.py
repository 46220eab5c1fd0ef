"""probe: does vitb_decode_batch_multi overlap the devices?  config 2 frames per device, pinned host memory"""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import viterbidecodercpp_b200 as v
from viterbidecodercpp_b200 import synth

code = {c.name: c for c in v.COMMON_CODES}["Voyager"]
dc = v.DECODE_TYPES["HARD8"](code.R)
F, L = 65536, 2048
tx, sym = synth.make_frames(code.K, code.R, code.G, 4096, L, 1, -1, 1, 4.0, 1)
sym = np.tile(sym, (F // 4096, 1)).astype(np.int8)
n_dev = v.load_library().vitb_device_count()
bt = v.ViterbiBranchTable(code.K, code.R, code.G, 1, -1, 1)
decs = [v.ViterbiDecoder_CUDA(bt, dc.decoder_config, device=d) for d in range(n_dev)]
bufs = []
for d in range(n_dev):
    bufs.append((torch.from_numpy(sym).pin_memory(), torch.zeros((F, L // 8), dtype=torch.uint8).pin_memory(),
                 torch.zeros(F, dtype=torch.int64).pin_memory(), torch.zeros(F, dtype=torch.int32).pin_memory()))
big = torch.from_numpy(np.concatenate([sym] * n_dev)).pin_memory()
bo = torch.zeros((F * n_dev, L // 8), dtype=torch.uint8).pin_memory()
ba = torch.zeros(F * n_dev, dtype=torch.int64).pin_memory()
bf = torch.zeros(F * n_dev, dtype=torch.int32).pin_memory()

def t(fn, n=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e3

for d in range(n_dev):
    s, o, a, f = bufs[d]
    print(f"device {d} alone, multi call with one handle: {t(lambda: v.decode_batch_multi_raw([decs[d]], s.data_ptr(), F, L, o.data_ptr(), a.data_ptr(), f.data_ptr(), row_stride=sym.shape[1])):.2f} ms")
print(f"all {n_dev} devices, one multi call: {t(lambda: v.decode_batch_multi_raw(decs, big.data_ptr(), F * n_dev, L, bo.data_ptr(), ba.data_ptr(), bf.data_ptr(), row_stride=sym.shape[1])):.2f} ms")
# enqueue on every device by hand, then synchronise: what the multi call should cost
streams = [torch.cuda.Stream(device=d) for d in range(n_dev)]
def by_hand():
    for d in range(n_dev):
        s, o, a, f = bufs[d]
        decs[d].decode_batch_async(s.data_ptr(), F, L, o.data_ptr(), a.data_ptr(), f.data_ptr(), stream=streams[d].cuda_stream, row_stride=sym.shape[1])
    for d in range(n_dev):
        streams[d].synchronize()
print(f"all {n_dev} devices, decode_batch_async per device then synchronise: {t(by_hand):.2f} ms")
def enqueue_only():
    t0 = time.perf_counter()
    for d in range(n_dev):
        s, o, a, f = bufs[d]
        decs[d].decode_batch_async(s.data_ptr(), F, L, o.data_ptr(), a.data_ptr(), f.data_ptr(), stream=streams[d].cuda_stream, row_stride=sym.shape[1])
    t1 = time.perf_counter()
    for d in range(n_dev):
        streams[d].synchronize()
    return (t1 - t0) * 1e3
print("host time of the enqueue alone:", [round(enqueue_only(), 2) for _ in range(3)], "ms")
